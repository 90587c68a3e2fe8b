"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: env sharding and gradient / statistics averaging."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pulse_b200.dist_utils import average_gradients, env_shard, minibatches_per_rank, rank_device_seed, sync_running_stats


def test_env_shard_partitions_exactly():
    for total, world in ((16384, 8), (16384, 1), (10, 4), (7, 8)):
        spans = [env_shard(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    assert minibatches_per_rank(2048, 32, 16384) == 4 and minibatches_per_rank(16384, 32, 16384) == 32
    with pytest.raises(ValueError):
        minibatches_per_rank(1000, 32, 16384)
    assert rank_device_seed(5, 3) == ("cuda:3", 8)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # a linear model's gradient over the full batch == average of the per-shard gradients (equal shard sizes)
        g = torch.Generator().manual_seed(0)
        x, y = torch.randn(64, 5, generator=g), torch.randn(64, generator=g)
        w = torch.zeros(5, requires_grad=True)
        a, b = env_shard(64, rank, world)
        loss = ((x[a:b] @ w - y[a:b]) ** 2).mean()
        loss.backward()
        flat = w.grad.clone()
        average_gradients(flat, world)
        w2 = torch.zeros(5, requires_grad=True)
        ((x @ w2 - y) ** 2).mean().backward()
        ok = torch.allclose(flat, w2.grad, atol=1e-6)
        mean = torch.full((3,), float(rank), dtype=torch.float64)
        var = torch.full((3,), 1.0 + rank, dtype=torch.float64)
        cnt = torch.tensor(10.0 * (rank + 1), dtype=torch.float64)
        sync_running_stats(mean, var, cnt, world)
        ok = ok and torch.allclose(mean, torch.full((3,), 0.5, dtype=torch.float64)) and abs(cnt.item() - 15.0) < 1e-12
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_gradient_and_stats_averaging_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def _chain_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pulse_b200.dist_utils import ChainReducer
        red = ChainReducer(world, 3)
        flat = torch.arange(30, dtype=torch.float32) * (rank + 1)          # three back-to-back "network" slices
        spans = ((0, 8), (8, 20), (20, 30))
        for chain in (2, 1, 0):                                            # any issue order, as the chains finish
            a, b = spans[chain]
            red.reduce(flat[a:b], chain)
        ok = torch.allclose(flat, torch.arange(30, dtype=torch.float32) * (sum(range(1, world + 1)) / world))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_chain_reducer_averages_every_slice_world2():
    """Per-chain gradient averaging (one communicator per backward chain) == one all-reduce over the whole flat buffer."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30100 + (os.getpid() % 500)
    procs = [ctx.Process(target=_chain_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def _flat_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # no NCCL / no CUDA here: the flat buffers must fall back to plain memory (peer = None), the moments stay whole, and the
        # documented fallback sequence -- average_gradients on the flat gradient buffer -- is what every rank then runs
        from pulse_b200.nets import FlatParams
        f = FlatParams("cpu")
        f.reserve(100, 7)
        f.reserve(33)
        f.finalize()                                       # default: peer mode wanted, unavailable -> silent plain allocation on CPU
        ok = f.peer is None and f.shard_span() == (0, f.numel) and f.numel % 64 == 0
        f.grads.fill_(float(rank + 1))
        average_gradients(f.grads, world)
        ok = ok and torch.allclose(f.grads, torch.full_like(f.grads, (world + 1) / 2))
        f.gather_moments()                                 # no-op outside peer mode (must not issue a collective on one rank only)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_flat_params_fall_back_to_plain_buffers_without_nccl_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30700 + (os.getpid() % 500)
    procs = [ctx.Process(target=_flat_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_peer_shard_spans_cover_the_flat_buffer():
    """The slice arithmetic of csrc/peer_adam.cu (per = ceil(n4 / world) float4s per rank) as FlatParams.shard_span states it."""
    from pulse_b200.nets import FlatParams
    for numel, world in ((5529600, 8), (5529600, 2), (64 * 7, 8), (64, 8), (31700032, 4)):
        f = FlatParams("cpu")
        f._numel = numel
        spans = []
        for r in range(world):
            f.peer = {"world": world, "rank": r}
            spans.append(f.shard_span())
        assert spans[0][0] == 0 and spans[-1][1] == numel
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1)) and all(a % 4 == 0 and b % 4 == 0 for a, b in spans)
