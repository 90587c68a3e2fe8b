"""Host side of the evaluation loop without a GPU: `EvalLoopB200` (chunking, wrapped last chunk, polling, success rate, failed keys,
frame-weighted metric means) driven through a NUMPY MODEL of the device state machine of csrc/eval_metrics.cu (same counters, same
stopping rule, same frame-counting rule; the per-frame metric values come from the oracle's formulas), against the oracle's restatement
of IMAmpAgent._post_step_eval (im_amp.py:244-363).  The CUDA kernel itself is compared with the same oracle in tests/test_gpu_eval.py."""
import numpy as np
import pytest

from oracle.eval_oracle import EvalOracle, p_mpjpe
from pulse_b200.evaluation import EvalLoopB200, summarise


class NumpyMetrics:
    """eval_accumulate_kernel + eval_advance_kernel, one call = one launch pair."""

    def __init__(self, N):
        self.N = N

    def begin_chunk(self, num_steps, bound=None):
        self.num_steps = np.asarray(num_steps).astype(np.int64)
        self.bound = self.N if bound is None else int(bound)
        self.max_steps = int(self.num_steps.max())
        self.step_count, self.done = 0, False
        self.term = np.zeros(self.N, dtype=bool)
        self.sums, self.counts = np.zeros((self.N, 5)), np.zeros((self.N, 3), dtype=np.int64)
        self.h1, self.h2 = np.zeros((self.N, 24, 3)), np.zeros((self.N, 24, 3))

    def step(self, pred, gt, terminate):
        if self.done:
            return
        s = self.step_count
        self.term |= (s <= self.num_steps - 1) & terminate.astype(bool)
        d = pred - gt
        counted = s < self.num_steps - 1
        mg = np.linalg.norm(d, axis=-1).mean(-1)
        pr, gr = pred - pred[:, :1], gt - gt[:, :1]
        ml = np.linalg.norm(pr - gr, axis=-1).mean(-1)
        mpa = p_mpjpe(pr, gr)
        mv = np.linalg.norm(d - self.h1, axis=-1).mean(-1)
        ma = np.linalg.norm(d - 2 * self.h1 + self.h2, axis=-1).mean(-1)
        self.h2, self.h1 = self.h1, d
        for k, val in enumerate((mg, ml, mpa)):
            self.sums[counted, k] += val[counted]
        self.counts[counted, 0] += 1
        if s >= 1:
            self.sums[counted, 3] += mv[counted]
            self.counts[counted, 1] += 1
        if s >= 2:
            self.sums[counted, 4] += ma[counted]
            self.counts[counted, 2] += 1
        running = ~self.term
        if running.any():
            rb = running[:self.bound]
            curr_max = int(self.num_steps[:self.bound][rb].max()) if rb.any() else s - 1
            if s >= curr_max:
                curr_max = s + 1
        else:
            curr_max = self.max_steps
        self.step_count = s + 1
        if s + 1 >= curr_max or not running.any():
            self.done = True

    def finished(self):
        return self.done

    def read(self):
        return {"sums": self.sums.copy(), "counts": self.counts.copy(), "terminated": self.term.copy(), "steps": self.step_count}


class Sim:
    def __init__(self, N, U, seed, fail_all_chunk=None):
        self.N, self.U = N, U
        rng = np.random.default_rng(seed)
        self.num_steps_all = rng.integers(8, 30, size=U)
        self.keys = np.array([f"clip_{i:03d}" for i in range(U)])
        self.fail_step = rng.integers(0, 50, size=U)
        self.fail_step[rng.random(U) < 0.5] = 10 ** 6
        if fail_all_chunk is not None:
            self.fail_step[fail_all_chunk * N:(fail_all_chunk + 1) * N] = 3
        self.s = 0

    def load_chunk(self, start_idx):
        self.s = 0
        self.ids = (start_idx + np.arange(self.N)) % self.U
        S = int(self.num_steps_all[self.ids].max()) + 12
        rng = np.random.default_rng(100 + start_idx)
        self.gt = rng.normal(size=(1, self.N, 24, 3)) + np.cumsum(rng.normal(scale=0.02, size=(S, self.N, 24, 3)), axis=0)
        self.pred = self.gt + np.cumsum(rng.normal(scale=0.004, size=(S, self.N, 24, 3)), axis=0) + rng.normal(scale=0.01, size=(S, self.N, 24, 3))
        return self.num_steps_all[self.ids], self.ids

    def frame(self):
        s = min(self.s, self.pred.shape[0] - 1)
        term = self.fail_step[self.ids] == s
        self.s += 1
        return self.pred[s], self.gt[s], term


@pytest.mark.parametrize("N,U,poll,fail_all", [(16, 40, 1, None), (16, 40, 5, None), (8, 16, 1, 0), (32, 20, 3, None), (4, 9, 1, 1)])
def test_eval_loop_host_logic_matches_oracle(N, U, poll, fail_all):
    sim = Sim(N, U, seed=N * 31 + U, fail_all_chunk=fail_all)
    orc = EvalOracle(N, U, sim.keys)
    steps = 0
    while True:
        num_steps, ids = sim.load_chunk(orc.start_idx)
        while True:
            pred, gt, term = sim.frame()
            done, end, info = orc.post_step(term, np.linalg.norm(pred - gt, axis=-1).mean(-1), pred, gt, num_steps, ids)
            steps += 1
            if done or end:
                break
        if end:
            break
    sim2 = Sim(N, U, seed=N * 31 + U, fail_all_chunk=fail_all)
    loop = EvalLoopB200(N, U, sim2.keys, load_chunk=sim2.load_chunk, reset_all=lambda: None, step=sim2.frame, poll_every=poll, metrics=NumpyMetrics(N))
    out = loop.run()
    assert out["steps"] == steps
    assert sorted(out["failed_keys"].tolist()) == sorted(np.asarray(info["failed_keys"]).tolist())
    for k, v in info["eval_info"].items():
        assert abs(out["eval_info"][k] - v) <= 1e-9 * max(1.0, abs(v)), (k, out["eval_info"][k], v)


def test_summarise_is_a_frame_weighted_mean():
    sums = np.array([[2.0, 1.0, 0.5, 0.3, 0.1], [6.0, 3.0, 1.5, 0.9, 0.4]])
    counts = np.array([[2, 1, 0], [6, 5, 4]])
    m = summarise(sums, counts)
    assert abs(m["mpjpe_g"] - 1000.0) < 1e-9 and abs(m["vel_dist"] - 200.0) < 1e-9 and abs(m["accel_dist"] - 125.0) < 1e-9
    assert abs(summarise(sums, counts, np.array([False, True]))["mpjpe_l"] - 500.0) < 1e-9
