"""Multi-GPU check + timing of the peer-memory optimizer step (csrc/peer_adam.cu) against the NCCL path it replaces.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/probe_peer.py [--numel 5529600]

Per rank: a FlatParams in peer mode (symmetric memory) and a plain one; identical parameters / moments, rank-dependent gradients.
Checks   peer step == all-reduce(AVG) + sum_squares + Adam  (parameters, bf16 operands, the rank's moment slice, cleared gradients),
         over several steps, eagerly and as a replayed CUDA graph, with peer loads/stores and (if the box has NVLS) the multicast path.
Times    both sequences with CUDA events (max over ranks) and prints one JSON line on rank 0."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def build_flat(numel, dev, peer):
    from pulse_b200.nets import FlatParams
    f = FlatParams(dev)
    f.reserve(numel)
    f.finalize(peer=peer)
    return f


def fill(f, rank, seed, it=0):
    g = torch.Generator(device=f.device).manual_seed(seed)
    f.params.copy_(torch.randn(f.numel, device=f.device, generator=g) * 0.05)
    f.exp_avg.copy_(torch.randn(f.numel, device=f.device, generator=g) * 1e-3)
    f.exp_avg_sq.copy_(torch.rand(f.numel, device=f.device, generator=g) * 1e-5)
    f.sync_bf16()
    f.step.zero_()


def set_grads(f, rank, it):
    g = torch.Generator(device=f.device).manual_seed(1000 * (rank + 1) + it)
    f.grads.copy_(torch.randn(f.numel, device=f.device, generator=g) * 0.02 * (rank + 1))
    f.clean = False


def nccl_step(f, world, lr, max_norm):
    from pulse_b200.dist_utils import average_gradients
    average_gradients(f.grads, world)
    f.adam_step(lr, max_norm=max_norm)


def timed(fn, iters, dev):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(dev)
    dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--numel", type=int, default=5529600)      # the PPO + discriminator flat buffer of im.yaml (22 MB fp32)
    ap.add_argument("--iters", type=int, default=50)
    args = ap.parse_args()
    os.environ.setdefault("PULSE_PEER_TIMEOUT_MS", "20000")
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    out = {"world": world, "numel": args.numel}
    lr, max_norm = 3e-3, 50.0
    ref = build_flat(args.numel, dev, peer=False)
    variants = [("p2p", "0")]
    if os.environ.get("PULSE_PROBE_MC", "1") == "1":
        variants.append(("multicast", "1"))
    for name, mc in variants:
        os.environ["PULSE_PEER_MC"] = mc
        f = build_flat(args.numel, dev, peer=True)
        if f.peer is None:
            out[name] = "peer mode unavailable"
            continue
        if mc == "1" and not f.peer["multicast"]:
            out[name] = "no multicast pointer on this box"
            continue
        out.setdefault("has_multicast", bool(getattr(f.peer["handles"][0], "multicast_ptr", 0)))
        fill(f, rank, 7)
        fill(ref, rank, 7)
        worst = 0.0
        for it in range(4):
            set_grads(f, rank, it)
            set_grads(ref, rank, it)
            nccl_step(ref, world, lr, max_norm)
            f.peer_adam_step(lr, max_norm=max_norm)
            torch.cuda.synchronize(dev)
            s0, s1 = f.shard_span()
            d = float((f.params - ref.params).abs().max())
            worst = max(worst, d)
            ok = (d <= 2e-6 and torch.equal(f.params_bf16, f.params.bfloat16()) and float(f.grads.abs().max()) == 0.0
                  and float((f.exp_avg[s0:s1] - ref.exp_avg[s0:s1]).abs().max()) <= 1e-8 and int(f.step.item()) == it + 1)
            # every rank must hold the same parameters bit for bit
            chk = f.params.double().sum().reshape(1).clone()
            lo, hi = chk.clone(), chk.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            ok = ok and bool(lo.item() == hi.item())
            flag = torch.tensor([1.0 if ok else 0.0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if flag.item() != 1.0:
                out[name] = f"MISMATCH at step {it}: max |dp| {d:.3e} on rank {rank}"
                break
        else:
            # CUDA-graph replay of the peer step
            g = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream(dev)
            s.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.graph(g, stream=s):
                f.peer_adam_step(lr, max_norm=max_norm)
            set_grads(f, rank, 99)
            set_grads(ref, rank, 99)
            nccl_step(ref, world, lr, max_norm)
            torch.cuda.synchronize(dev)
            g.replay()
            torch.cuda.synchronize(dev)
            dg = float((f.params - ref.params).abs().max())
            t_peer = timed(lambda: f.peer_adam_step(lr, max_norm=max_norm), args.iters, dev)
            t_graph = timed(lambda: g.replay(), args.iters, dev)
            out[name] = {"max_abs_dp": worst, "graph_replay_max_abs_dp": dg, "us_per_step": round(t_peer, 2), "us_per_step_graph": round(t_graph, 2)}
        del f
    t_nccl = timed(lambda: nccl_step(ref, world, lr, max_norm), args.iters, dev)
    from pulse_b200.dist_utils import average_gradients
    t_ar = timed(lambda: average_gradients(ref.grads, world), args.iters, dev)
    out["nccl_allreduce_plus_adam_us"] = round(t_nccl, 2)
    out["nccl_allreduce_only_us"] = round(t_ar, 2)
    if rank == 0:
        print("PROBE " + json.dumps(out), flush=True)
    dist.barrier()
    torch.cuda.synchronize(dev)
    os._exit(0)


if __name__ == "__main__":
    main()
