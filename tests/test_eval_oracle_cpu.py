"""The evaluation oracle's metric restatements (smpl_sim compute_metrics_lite [3P-memory], parity unpinned) checked through the
properties that define them, and the bookkeeping restatement (im_amp.py:244-363) on a hand-checkable case."""
import numpy as np

from oracle.eval_oracle import EvalOracle, compute_error_accel, compute_error_vel, compute_metrics_lite, p_mpjpe


def _rot(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def test_procrustes_is_invariant_to_similarity_transforms():
    rng = np.random.default_rng(0)
    target = rng.normal(size=(5, 24, 3))
    pred = np.stack([2.5 * (target[t] @ _rot(rng)) + rng.normal(size=(1, 3)) for t in range(5)])
    assert np.all(p_mpjpe(pred, target) < 1e-9)
    noisy = pred + rng.normal(scale=0.01, size=pred.shape)
    e = p_mpjpe(noisy, target)
    assert np.all(e > 0) and np.all(e < 0.05)
    # a reflection is NOT removed (proper rotations only)
    assert np.all(p_mpjpe(target * np.array([1.0, 1.0, -1.0]), target) > 0.1)


def test_velocity_and_acceleration_errors():
    rng = np.random.default_rng(1)
    gt = np.cumsum(rng.normal(size=(9, 24, 3)), axis=0)
    assert np.allclose(compute_error_vel(gt + 0.3, gt), 0) and np.allclose(compute_error_accel(gt + 0.3, gt), 0)
    drift = gt + np.arange(9)[:, None, None] * np.array([0.1, 0.0, 0.0])
    assert np.allclose(compute_error_vel(drift, gt), 0.1) and np.allclose(compute_error_accel(drift, gt), 0, atol=1e-12)
    m = compute_metrics_lite([drift], [gt])
    assert m["mpjpe_g"].shape == (9, 24) and m["vel_dist"].shape == (8,) and m["accel_dist"].shape == (7,) and m["mpjpe_pa"].shape == (9,)
    assert np.allclose(m["mpjpe_l"], 0, atol=1e-9)          # a pure translation vanishes root-relative


def test_bookkeeping_small_case():
    """2 envs, 3 unique clips -> two chunks, the second wrapped (env 1 repeats clip 0): success rate over the first 3 sequences only."""
    orc = EvalOracle(2, 3, ["a", "b", "c"])
    rng = np.random.default_rng(3)
    pos = rng.normal(size=(2, 24, 3))
    steps = {0: np.array([4, 3]), 2: np.array([5, 4])}
    ids = {0: np.array([0, 1]), 2: np.array([2, 0])}
    n = 0
    while True:
        ns, ci = steps[orc.start_idx], ids[orc.start_idx]
        term = np.array([False, orc.start_idx == 0 and orc.curr_stpes == 1])      # clip "b" fails at step 1 of chunk 0
        done, end, info = orc.post_step(term, np.zeros(2), pos + rng.normal(scale=0.01, size=pos.shape), pos, ns, ci)
        n += 1
        if end:
            break
    assert list(info["failed_keys"]) == ["b"] and list(info["success_keys"]) == ["a", "c"]
    assert abs(info["eval_info"]["eval_success_rate"] - 2 / 3) < 1e-12
    assert n == 4 + 5        # chunk 0 runs to max steps of the non-terminated env (4); the wrapped chunk to clip c's 5 steps
