"""Device-side MotionLib loader (MotionLibB200.from_clips) timed against the reference's ~60 ms / clip of per-frame Python
loops + mp.Process fan-out (motion_lib_base.py:179-323): N AMASS-shaped synthetic clips (lognormal lengths, median 150
frames @ 30 fps) in the on-disk schema -> packed device tables, end to end (host concatenation, H2D copies, three kernels, record packing).

    python tools/bench_loader.py [--clips 2048] [--json out.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pulse_b200.motion_lib import MotionLibB200  # noqa: E402
from tests.helpers import load_npz  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=2048)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    z = load_npz("loader.npz")               # skeleton (parents, local translations) of the reference's smpl_humanoid.xml
    rng = np.random.default_rng(0)
    nf = np.clip(np.exp(rng.normal(np.log(150), 0.6, size=a.clips)).astype(np.int64), 5, 1800)
    clips = []
    for n in nf:
        q = rng.normal(size=(n, 24, 4))
        q /= np.linalg.norm(q, axis=-1, keepdims=True)
        clips.append({"pose_quat_global": q, "root_trans_offset": torch.from_numpy(rng.normal(size=(n, 3))), "pose_aa": rng.normal(size=(n, 72)),
                      "fps": 30.0})
    heads = np.pi * (2 * rng.random(a.clips) - 1)
    parents, loc = z["parents"].tolist(), z["local_translation"].numpy()
    MotionLibB200.from_clips(clips[:8], parents, loc, "cuda:0", headings=heads[:8])      # warm-up (library load, kernel attributes)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ml = MotionLibB200.from_clips(clips, parents, loc, "cuda:0", headings=heads)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out = {"clips": a.clips, "frames": int(nf.sum()), "seconds": dt, "ms_per_clip": 1e3 * dt / a.clips, "reference_ms_per_clip": 60.0,
           "speedup_vs_reference_loader": 60.0 / (1e3 * dt / a.clips), "tables_mb": ml.frame_rec.numel() * 4 / 1e6 + ml.aux_rec.numel() * 4 / 1e6,
           "note": "wall clock incl. host-side concatenation of the clip arrays and the H2D copies; reference figure from SURVEY.md 8f-1"}
    print(json.dumps(out))
    if a.json:
        open(a.json, "w").write(json.dumps(out) + "\n")


if __name__ == "__main__":
    main()
