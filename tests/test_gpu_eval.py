"""Evaluation loop with device-side metrics (pulse_eval_step, pulse_b200/evaluation.py) against the CPU restatement of
IMAmpAgent._post_step_eval + compute_metrics_lite (oracle/eval_oracle.py) on a synthetic simulator: several chunks, a wrapped last
chunk, envs that fail inside / after their sequence, chunks that end early because every env failed."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class _SynthSim:
    """Deterministic stand-in for (task + MotionLib) during evaluation: positions are functions of (chunk, env, step)."""

    def __init__(self, N, U, seed, fail_all_chunk=None):
        self.N, self.U = N, U
        self.rng = np.random.default_rng(seed)
        self.num_steps_all = self.rng.integers(12, 48, size=U).astype(np.int64)
        self.keys = np.array([f"clip_{i:04d}" for i in range(U)])
        self.fail_step = self.rng.integers(0, 80, size=U)                 # terminate flag raised at this step (>= num_steps: not a failure)
        self.fail_step[self.rng.random(U) < 0.5] = 10 ** 6               # half of the clips never fail
        if fail_all_chunk is not None:
            lo = fail_all_chunk * N
            self.fail_step[lo:lo + N] = self.rng.integers(2, 6, size=min(N, U - lo))
        self.start, self.s = 0, 0

    def load_chunk(self, start_idx):
        self.start, self.s = start_idx, 0
        self.ids = (start_idx + np.arange(self.N)) % self.U              # sequential sampling wraps in the last chunk
        S = int(self.num_steps_all[self.ids].max()) + 12
        rng = np.random.default_rng(1000 + start_idx)
        base = rng.normal(size=(1, self.N, 24, 3)).astype(np.float32) + np.cumsum(rng.normal(scale=0.02, size=(S, self.N, 24, 3)), axis=0).astype(np.float32)
        noise = np.cumsum(rng.normal(scale=0.004, size=(S, self.N, 24, 3)), axis=0).astype(np.float32) + rng.normal(scale=0.01, size=(S, self.N, 24, 3)).astype(np.float32)
        self.gt, self.pred = base, base + noise
        return self.num_steps_all[self.ids], self.ids

    def frame(self):
        s = self.s
        term = (self.fail_step[self.ids] == s) | ((self.fail_step[self.ids] < s) & (s % 7 == 0))
        out = self.pred[s], self.gt[s], term
        self.s += 1
        return out


def _oracle_run(sim):
    from oracle.eval_oracle import EvalOracle
    orc = EvalOracle(sim.N, sim.U, sim.keys)
    steps = 0
    while True:
        num_steps, ids = sim.load_chunk(orc.start_idx)
        while True:
            pred, gt, term = sim.frame()
            mpjpe = np.linalg.norm(pred - gt, axis=-1).mean(-1)
            chunk_done, end, info = orc.post_step(term, mpjpe, pred, gt, num_steps, ids)
            steps += 1
            if chunk_done or end:
                break
        if end:
            return info, steps


@pytest.mark.parametrize("N,U,poll,fail_all", [(64, 150, 1, None), (64, 150, 8, None), (32, 64, 1, 0), (128, 100, 4, None), (16, 33, 1, 1)])
def test_eval_loop_matches_oracle(N, U, poll, fail_all):
    from pulse_b200.evaluation import EvalLoopB200
    dev = torch.device("cuda:0")
    ref_info, ref_steps = _oracle_run(_SynthSim(N, U, seed=N + U, fail_all_chunk=fail_all))
    sim = _SynthSim(N, U, seed=N + U, fail_all_chunk=fail_all)
    state = torch.zeros(N, 26, 13, device=dev)                           # body positions read in place from a [N, B, 13] rigid-body view
    gt_buf = torch.zeros(N, 24, 3, device=dev)
    term_buf = torch.zeros(N, dtype=torch.int64, device=dev)

    def step():
        if sim.s >= sim.pred.shape[0]:                                   # polling may run a few steps past the end of a chunk: ignored by the kernel
            return state, gt_buf, term_buf
        pred, gt, term = sim.frame()
        state[:, :24, :3].copy_(torch.from_numpy(pred))
        gt_buf.copy_(torch.from_numpy(gt))
        term_buf.copy_(torch.from_numpy(term.astype(np.int64)))
        return state, gt_buf, term_buf

    loop = EvalLoopB200(N, U, sim.keys, load_chunk=sim.load_chunk, reset_all=lambda: None, step=step, device=dev, poll_every=poll)
    out = loop.run()
    assert out["steps"] == ref_steps                                     # the device-side stopping rule ended every chunk where the reference does
    assert sorted(out["failed_keys"].tolist()) == sorted(np.asarray(ref_info["failed_keys"]).tolist())
    assert sorted(out["success_keys"].tolist()) == sorted(np.asarray(ref_info["success_keys"]).tolist())
    for k, v in ref_info["eval_info"].items():
        got = out["eval_info"][k]
        assert abs(got - v) <= 2e-4 * max(1.0, abs(v)), (k, got, v)
