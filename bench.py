#!/usr/bin/env python
"""bench.py -- env-steps/s of the PULSE HumanoidIm hot path on B200 (BASELINE.json metric).

One bench "step" = one PPO iteration of BASELINE config C4 (HumanoidIm PPO, 16384 envs total,
AMASS-shaped synthetic MotionLib, horizon 32): 32 post-physics env steps (fused reward/reset/obs
kernel + AMP-obs kernel each) followed by the rollout post-processing (GAE / returns / advantage
normalisation).  Isaac Gym physics is excluded on every arm (not installable here).
`config.phases` lists exactly what runs inside the timed region and `config.not_yet` what the
reference iteration additionally does that this build does not run yet.

Contract: `python bench.py --gpus N --steps K --warmup W` (torchrun for N > 1), one JSON line on
rank 0.  `--impl reference` times the CPU port of the reference path (oracle/, kind "port") on the
host cores.  Envs shard across ranks (16384 / N each, "strong" scaling); no data-path collective in
the rollout; NCCL is only used for the timing barrier / max-over-ranks here.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TOTAL_ENVS = 16384
HORIZON = 32
ALGO_BYTES_PER_ENV_STEP = 9396  # fused step kernel, core total incl. power term
METRIC = "env-steps/sec at 16384 humanoid envs, 1/2/4/8 B200; obs-kernel HBM GB/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=TOTAL_ENVS)
    ap.add_argument("--median-frames", type=int, default=150)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons with nvidia-smi during the timed region."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        sm, reasons, mx = [], set(), None
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = float(s[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
def pick_cpu_threads(max_threads):
    """PyTorch CPU ops on [N,24,4]-sized tensors stop scaling (and can collapse) long before 128
    threads; give the CPU arm the thread count at which it runs fastest on this host."""
    best, best_rate = 1, 0.0
    for t in sorted({1, 4, 8, 16, 32, 64, max_threads}):
        if t > max_threads:
            continue
        rate, _ = cpu_port_rate(1024, 1, t, gae=False)
        if rate > best_rate:
            best, best_rate = t, rate
    return best


def cpu_port_rate(n_envs, n_iters, threads, gae=True):
    """Reference path on the host CPU: the oracle port of get_motion_state x2 + reward + reset + self/task
    obs + AMP obs + GAE, timed per env-step on a bounded sample."""
    import torch
    from oracle import pulse_oracle as po
    from tests.helpers import synthetic_step_inputs, synthetic_tables
    torch.set_num_threads(threads)
    tb = synthetic_tables(min(n_envs, 2048), seed=0)
    z = synthetic_step_inputs(tb, n_envs, seed=1)
    cfg = po.ImStepConfig()
    amp = torch.zeros(n_envs, 10, 196)
    T = HORIZON
    r, v, nv = (torch.randn(T, n_envs, 1) for _ in range(3))
    d = (torch.rand(T, n_envs) < 0.05).float()

    def one_env_step():
        po.humanoid_im_step(tb, cfg, z["body_state"], z["dof_vel"], z["dof_force"], z["progress_buf"], z["motion_ids"],
                            z["start_times"], z["start_offset"], z["global_offset"], z["cycle_counter"], z["reset_buf_in"])
        return po.amp_obs_step(amp, z["body_state"], z["dof_pos"], z["dof_vel"])

    def gae_pass():
        adv = po.discount_values(d, v, r, nv)
        return po.normalized_advantages(po.swap_and_flatten01(adv + v), po.swap_and_flatten01(v))

    one_env_step()
    t0 = time.perf_counter()
    for _ in range(n_iters):
        one_env_step()
    t_step = (time.perf_counter() - t0) / n_iters
    t_gae = 0.0
    if gae:
        gae_pass()
        t0 = time.perf_counter()
        gae_pass()
        t_gae = time.perf_counter() - t0
    per_iter = HORIZON * t_step + t_gae  # one PPO iteration over n_envs
    return HORIZON * n_envs / per_iter, per_iter


def run_reference(a):
    """--impl reference: the reference's CPU path (oracle port), all host threads, bounded sample."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = pick_cpu_threads(os.cpu_count() or 1)
    n = min(a.envs, 4096)
    vals, per = [], []
    for i in range(a.warmup + a.steps):
        rate, per_iter = cpu_port_rate(n, 2, threads)
        if i >= a.warmup:
            vals.append(rate)
            per.append(per_iter)
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * sum(per) / len(per) * (a.envs / n), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(a, 1, a.envs),
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": threads, "host_cores": os.cpu_count(), "kind": "port",
                         "sample": f"{n} envs x 2 env-steps + one GAE pass per step, scaled to a 32-step iteration; torch {torch.__version__} CPU"},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(a, world, envs_total):
    return {
        "workload": "HumanoidIm PPO iteration, 16384 envs total, AMASS-shaped synthetic MotionLib (one clip per env, "
                    f"lognormal lengths, median {a.median_frames} frames @30fps), horizon 32 (BASELINE configs[3])",
        "envs_total": envs_total, "envs_per_gpu": envs_total // world, "horizon": HORIZON, "parallelism": f"env-shard x{world}",
        "phases": ["32x fused reward+reset+obs kernel (K1-K5)", "32x AMP obs + history kernel (K6)", "GAE + returns + adv-norm (K11,K12)"],
        "not_yet": ["policy/value/disc MLP forward (K7-K10)", "PPO/disc update: fwd+bwd+Adam+grad allreduce (K13-K16)"],
        "physics": "excluded (Isaac Gym not installable; state tensors are synthetic, resident in HBM)",
        "l2": "256 MiB L2 flush write before every timed iteration; per-iteration inputs (tables 3.9 GB/rank at N=1) exceed L2",
    }


def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)
    import torch
    import torch.distributed as dist
    from pulse_b200 import _lib
    from pulse_b200.humanoid_im import HumanoidImCompute
    from pulse_b200.motion_lib import MotionLibB200
    from pulse_b200.rollout import discount_values
    from tools.synth import device_step_inputs, device_tables

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == a.gpus, f"--gpus {a.gpus} but WORLD_SIZE={world} (launch with torch.distributed.run)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = a.envs // world
    lib = _lib.load()

    # ---- synthetic inputs, resident in HBM (shard: envs [rank*n, (rank+1)*n), one clip per env) ----
    tabs = device_tables(n, dev, seed=100 + rank, median_frames=a.median_frames)
    ml = MotionLibB200.from_tables(tabs)
    del tabs
    z = device_step_inputs(ml, n, seed=200 + rank)
    comp = HumanoidImCompute(ml)
    T = HORIZON
    obses = torch.zeros(T, n, 934, device=dev)          # experience buffer slice the obs kernel writes into
    rewards = torch.zeros(T, n, device=dev)
    reward_raw = torch.zeros(n, 5, device=dev)
    reset_buf = torch.zeros(n, dtype=torch.long, device=dev)
    term_buf = torch.zeros(n, dtype=torch.long, device=dev)
    dones = torch.zeros(T, n, device=dev)
    amp_buf = torch.zeros(n, 10, 196, device=dev)
    values = torch.randn(T, n, 1, device=dev)            # stand-ins until the critic MLP is in the loop
    next_values = torch.randn(T, n, 1, device=dev)
    progress0 = z["progress_buf"].clone()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    # pinned host mirrors for the end-to-end arm
    h_body = z["body_state"].cpu().pin_memory()
    h_dof = z["dof_state"].cpu().pin_memory()
    h_force = z["dof_force"].cpu().pin_memory()
    h_rew = torch.empty(n, dtype=torch.float32).pin_memory()
    h_reset = torch.empty(n, dtype=torch.long).pin_memory()
    h_term = torch.empty(n, dtype=torch.long).pin_memory()
    h2d = h_body.numel() * 4 + h_dof.numel() * 4 + h_force.numel() * 4
    d2h = n * (4 + 8 + 8)
    step_events = []

    def iteration(e2e, record):
        z["progress_buf"].copy_(progress0)
        for t in range(T):
            if e2e:
                z["body_state"].copy_(h_body, non_blocking=True)
                z["dof_state"].copy_(h_dof, non_blocking=True)
                z["dof_force"].copy_(h_force, non_blocking=True)
            z["progress_buf"] += 1
            if record:
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
            comp.step(body_state=z["body_state"], dof_vel=z["dof_vel"], dof_force=z["dof_force"], progress_buf=z["progress_buf"],
                      motion_ids=z["motion_ids"], motion_start_times=z["motion_start_times"], motion_start_offset=z["motion_start_offset"],
                      global_offset=z["global_offset"], cycle_counter=z["cycle_counter"], obs_buf=obses[t], rew_buf=rewards[t],
                      reward_raw=reward_raw, reset_buf=reset_buf, terminate_buf=term_buf)
            if record:
                e.record()
                step_events.append((s, e))
            comp.amp_obs(body_state=z["body_state"], dof_pos=z["dof_pos"], dof_vel=z["dof_vel"], amp_obs_buf=amp_buf)
            dones[t].copy_(reset_buf)
            if e2e:
                h_rew.copy_(rewards[t], non_blocking=True)
                h_reset.copy_(reset_buf, non_blocking=True)
                h_term.copy_(term_buf, non_blocking=True)
        adv, ret = discount_values(dones, values, rewards, next_values, normalize_advantage=True)
        return adv

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(e2e, steps, record):
        for _ in range(a.warmup):
            flush.fill_(1)
            iteration(e2e, False)
        barrier()
        launches0 = lib.pulse_launch_count()
        total_ms = 0.0
        for _ in range(steps):
            flush.fill_(1)  # L2 flush, outside the timed span
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            s.record()
            iteration(e2e, record)
            e.record()
            barrier()
            ms = torch.tensor([s.elapsed_time(e)], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            total_ms += float(ms.item())
        return total_ms / steps, (lib.pulse_launch_count() - launches0) // steps

    sampler = ClockSampler(local)
    sampler.start()
    ms_dev, launches = timed(False, a.steps, True)
    clocks = sampler.stop()
    ms_e2e, _ = timed(True, max(2, a.steps // 2), False)

    # dominant kernel: fused step kernel, live CUDA-event duration inside the timed region
    torch.cuda.synchronize()
    k_ms = sorted(s.elapsed_time(e) for s, e in step_events)
    k_avg = sum(k_ms) / len(k_ms)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = ALGO_BYTES_PER_ENV_STEP * n / (k_avg * 1e-3) / 1e9
    env_steps = T * a.envs
    if rank == 0:
        line = {
            "metric": METRIC, "value": env_steps / (ms_dev * 1e-3), "unit": "env-steps/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(a, world, a.envs),
            "e2e": {"value": env_steps / (ms_e2e * 1e-3), "unit": "env-steps/s", "h2d_bytes_per_step": T * h2d * world,
                    "d2h_bytes_per_step": T * d2h * world, "ms_per_step": ms_e2e},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"kernel": "im_step_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                         "algorithmic_bytes_per_env_step": ALGO_BYTES_PER_ENV_STEP, "avg_launch_ms": k_avg, "launches_timed": len(k_ms),
                         "note": "event pairs include launch gaps of back-to-back stream work; see profiles/ for ncu per-launch times"},
        }
        if not a.no_cpu_baseline:
            threads = pick_cpu_threads(os.cpu_count() or 1)
            rate, per_iter = cpu_port_rate(2048, 2, threads)
            line["cpu_baseline"] = {"value": rate, "unit": "env-steps/s", "cores": threads, "host_cores": os.cpu_count(), "kind": "port",
                                    "sample": "2048 envs x 2 env-steps + one GAE pass (oracle port of the reference PyTorch path), scaled to a 32-step iteration"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
