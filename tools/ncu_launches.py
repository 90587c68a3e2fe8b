"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel; optionally print one window."""
import collections
import csv
import re
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    gi = hdr.index("Grid Size") if "Grid Size" in hdr else None
    seq = []
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
        name = re.sub(r"\(.*", "", r[ki]).split("::")[-1][:48]
        seq.append((name, v, r[gi] if gi is not None else ""))
    return seq


def main(path, window_kernel=None, out=None):
    seq = load(path)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, v, _ in seq:
        agg[n][0] += 1
        agg[n][1] += v
    tot = sum(v[1] for v in agg.values())
    lines = [f"launches {len(seq)}  total {tot:.1f} us (cold-cache, serialised: compare SHARES)"]
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        lines.append(f"{t / tot * 100:5.1f}%  {t:10.1f} us  n={c:5d}  avg {t / c:8.2f} us  {k}")
    if window_kernel:
        idx = [i for i, (n, _, _) in enumerate(seq) if n.startswith(window_kernel)]
        if len(idx) > 3:
            i0, i1 = idx[2], idx[3]
            lines.append(f"--- one period between consecutive {window_kernel} launches ({i1 - i0} launches, "
                         f"{sum(v for _, v, _ in seq[i0:i1]):.1f} us)")
            for n, v, g in seq[i0:i1]:
                lines.append(f"{v:8.1f} us  {g:>14s}  {n}")
    text = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(text)
    print(text)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None, sys.argv[3] if len(sys.argv) > 3 else None)
