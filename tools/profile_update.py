"""One PPO + discriminator minibatch update (bench.py's update_mb, im.yaml sizes) for profiling.

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
      python tools/profile_update.py            # launch list of exactly ONE minibatch (eager, serialised)
  python tools/profile_update.py --time         # wall time per minibatch: eager / CUDA graph (what bench.py replays)
"""
import argparse
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pulse_b200 import _lib  # noqa: E402
if os.environ.get("PULSE_ALT_LIB"):
    _lib.LIB_PATH = os.environ["PULSE_ALT_LIB"]
from pulse_b200.ppo import PPOPolicy  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=16384)
    ap.add_argument("--amp-rows", type=int, default=4096)
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--prefetch", default="none", choices=("none", "start", "loss", "reduce"),
                    help="prepare the next minibatch's normalised operands on a side stream, forked at this point of the step")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    M, B = a.rows, a.amp_rows
    pol = PPOPolicy(device=dev, seed=0, with_disc=True)
    obs = torch.randn(M, 934, device=dev, generator=g)
    act = torch.randn(M, 69, device=dev, generator=g) * 0.1
    mu = torch.randn(M, 69, device=dev, generator=g) * 0.1
    nlp = torch.randn(M, device=dev, generator=g) + 60
    adv, ret = torch.randn(M, device=dev, generator=g), torch.randn(M, device=dev, generator=g)
    amp = tuple(torch.randn(B, 1960, device=dev, generator=g) for _ in range(3))

    if a.prefetch != "none":
        import os
        os.environ["PULSE_PREFETCH_AT"] = a.prefetch
        pol.prepare_inputs(obs, amp, slot=0)

    def step():
        if a.prefetch == "none":
            pol.train_minibatch(obs, act, nlp, adv, ret, old_mu=mu, amp=amp)
        else:       # timing only: slot 0 keeps its prepared operands, the prefetch refills slot 1 every step
            pol.train_minibatch(obs, act, nlp, adv, ret, old_mu=mu, amp=amp, slot=0, prepared=True, prefetch=(obs, amp))

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if not a.time:
        rt = ctypes.CDLL("libcudart.so")
        rt.cudaProfilerStart()
        step()
        torch.cuda.synchronize()
        rt.cudaProfilerStop()
        return

    def timed(fn, iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s.record()
        for _ in range(iters):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / iters * 1e3

    eager = timed(step, a.iters)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        step()
    graph = timed(gr.replay, a.iters)
    print(f"minibatch update M={M} amp={B}: eager {eager:.1f} us, graph {graph:.1f} us")


if __name__ == "__main__":
    main()
