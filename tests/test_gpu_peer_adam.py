"""pulse_peer_reduce_adam (csrc/peer_adam.cu) on ONE GPU: `world` ranks are emulated by `world` buffer sets on the same device and
`world` concurrent launches on separate streams -- the flag protocol, the slice arithmetic, the norm exchange and the push of the new
parameters run exactly as over NVLink (only the peer mapping is local).  Reference: the single-GPU sequence it replaces on every rank,
all-reduce(AVG) -> pulse_sum_squares -> pulse_adam_step (amp_agent.py:725-750).  The real 2- and 8-GPU runs: tools/probe_peer.py."""
import ctypes as C
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _isolated(case: str):
    """The emulated ranks WAIT for each other on the device: should the launches ever fail to be co-resident, the bounded wait traps and
    the CUDA context is lost -- so every case runs in its own process (the pytest process keeps its context)."""
    env = dict(os.environ, CUDA_DEVICE_MAX_CONNECTIONS="32", PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, os.path.abspath(__file__), case], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, f"case {case} failed:\n{r.stdout[-2000:]}\n{r.stderr[-4000:]}"


def _rank_state(world, n, dev, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    params = torch.randn(n, device=dev, generator=g) * 0.05
    m = torch.randn(n, device=dev, generator=g) * 1e-3
    v = torch.rand(n, device=dev, generator=g) * 1e-5
    grads = [torch.randn(n, device=dev, generator=g) * (0.02 * (r + 1)) for r in range(world)]
    return params, m, v, grads


def _reference(params, m, v, grads, step0, max_norm, lr):
    """The path the kernel replaces, through the library's own single-GPU entry points."""
    from pulse_b200 import _lib
    lib = _lib.load()
    dev = params.device
    avg = grads[0].clone()
    for g in grads[1:]:
        avg += g
    avg *= 1.0 / len(grads)
    p, m, v = params.clone(), m.clone(), v.clone()
    pb = torch.zeros_like(p, dtype=torch.bfloat16)
    sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
    step = torch.full((1,), step0, dtype=torch.int32, device=dev)
    sync = torch.zeros(1, dtype=torch.int32, device=dev)
    st = _lib.current_stream(dev)
    _lib.check(lib.pulse_sum_squares(avg.data_ptr(), avg.numel(), sumsq.data_ptr(), st), "sum_squares")
    _lib.check(lib.pulse_adam_step(p.data_ptr(), avg.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), sumsq.data_ptr(), max_norm, lr, 0.9, 0.999, 1e-8,
                                   step.data_ptr(), pb.data_ptr(), 3, sync.data_ptr(), st), "adam_step")
    torch.cuda.synchronize()
    return p, pb, m, v, int(step.item())


class _Rank:
    def __init__(self, rank, world, n, params, m, v, dev, grid):
        from pulse_b200 import _lib
        self.grads = torch.zeros(n, device=dev)
        self.params, self.m, self.v = params.clone(), m.clone(), v.clone()
        self.pb = self.params.bfloat16()
        self.sig = torch.zeros(256, dtype=torch.uint8, device=dev)
        self.step = torch.zeros(1, dtype=torch.int32, device=dev)
        self.epoch = torch.zeros(1, dtype=torch.int32, device=dev)
        self.partials = torch.zeros(_lib.PEER_MAX_GRID, dtype=torch.float64, device=dev)
        self.bar = torch.zeros(1, dtype=torch.int64, device=dev)
        self.stream = torch.cuda.Stream(dev)
        self.args = _lib.PeerAdamArgs(rank=rank, world=world, count=n, grid=grid, timeout_ms=5000)

    def bind(self, ranks, max_norm, lr):
        a = self.args
        for p, r in enumerate(ranks):
            a.grads[p], a.params[p], a.params_bf16[p], a.signals[p] = r.grads.data_ptr(), r.params.data_ptr(), r.pb.data_ptr(), r.sig.data_ptr()
        a.exp_avg, a.exp_avg_sq, a.step, a.epoch = self.m.data_ptr(), self.v.data_ptr(), self.step.data_ptr(), self.epoch.data_ptr()
        a.cta_partials, a.grid_bar = self.partials.data_ptr(), self.bar.data_ptr()
        a.max_norm, a.lr, a.beta1, a.beta2, a.eps = max_norm, lr, 0.9, 0.999, 1e-8

    def launch(self):
        from pulse_b200 import _lib
        with torch.cuda.stream(self.stream):
            _lib.check(_lib.load().pulse_peer_reduce_adam(C.byref(self.args), C.c_void_p(self.stream.cuda_stream)), "pulse_peer_reduce_adam")


CASES = [(1, 4096 + 64, 0.0), (2, 300000, 50.0), (2, 1028, 0.5), (4, 1000000 + 4, 1.0), (8, 5529600, 50.0)]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_peer_reduce_adam_matches_allreduce_adam(case):
    _isolated(f"match:{case}")


def test_peer_reduce_adam_graph_replay():
    _isolated("graph")


def _case_match(world, n, max_norm):
    dev = torch.device("cuda:0")
    grid = max(1, 128 // world)          # all `world` launches must be co-resident on the one GPU: they wait for each other
    params, m, v, grads = _rank_state(world, n, dev, seed=world * 1000 + n % 97)
    ranks = [_Rank(r, world, n, params, m, v, dev, grid) for r in range(world)]
    lr = 3e-3
    for r in ranks:
        r.bind(ranks, max_norm, lr)
    ref_p, ref_m, ref_v = params, m, v
    for it in range(3):                  # several optimizer steps: epoch flags, step counter, moments carried in the slices
        gs = [g * (1.0 + 0.25 * it) for g in grads]
        ref_p, ref_pb, ref_m, ref_v, ref_step = _reference(ref_p, ref_m, ref_v, gs, it, max_norm, lr)
        for r, g in zip(ranks, gs):
            r.grads.copy_(g)
        torch.cuda.synchronize()
        for r in ranks:
            r.launch()
        torch.cuda.synchronize()
        n4 = n // 4
        per = (n4 + world - 1) // world
        for k, r in enumerate(ranks):
            torch.testing.assert_close(r.params, ref_p, rtol=2e-6, atol=1e-7)
            assert torch.equal(r.params, ranks[0].params) and torch.equal(r.pb, ranks[0].pb)      # replicas stay bit-identical
            assert torch.equal(r.pb, r.params.bfloat16())
            assert int(r.step.item()) == ref_step == it + 1 and int(r.epoch.item()) == it + 1
            assert float(r.grads.abs().max()) == 0.0                                              # consumed and cleared
            s0, s1 = 4 * min(per * k, n4), 4 * min(per * k + per, n4)
            torch.testing.assert_close(r.m[s0:s1], ref_m[s0:s1], rtol=2e-6, atol=1e-9)            # sharded moments: the rank's slice only
            torch.testing.assert_close(r.v[s0:s1], ref_v[s0:s1], rtol=2e-6, atol=1e-12)


def _case_graph():
    """The launch is CUDA-graph replayable (device-side epoch / step counters): two emulated ranks, each captured on its own stream."""
    dev = torch.device("cuda:0")
    world, n = 2, 65536
    params, m, v, grads = _rank_state(world, n, dev, seed=5)
    ranks = [_Rank(r, world, n, params, m, v, dev, 32) for r in range(world)]
    for r in ranks:
        r.bind(ranks, 50.0, 1e-3)
    graphs = []
    for r in ranks:      # capture only (nothing runs): one graph per emulated rank
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=r.stream):
            r.launch()
        graphs.append(g)
    ref_p, ref_m, ref_v = params, m, v
    for it in range(3):
        ref_p, _, ref_m, ref_v, _ = _reference(ref_p, ref_m, ref_v, grads, it, 50.0, 1e-3)
        for r, g in zip(ranks, grads):
            r.grads.copy_(g)
        torch.cuda.synchronize()
        for r, g in zip(ranks, graphs):
            with torch.cuda.stream(r.stream):
                g.replay()
        torch.cuda.synchronize()
        for r in ranks:
            torch.testing.assert_close(r.params, ref_p, rtol=2e-6, atol=1e-7)
            assert int(r.step.item()) == it + 1


if __name__ == "__main__":
    what = sys.argv[1]
    if what.startswith("match:"):
        _case_match(*CASES[int(what.split(":")[1])])
    else:
        _case_graph()
    print("ok", what)
