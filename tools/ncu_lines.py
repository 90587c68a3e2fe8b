"""Join ncu per-SASS-instruction counts with nvdisasm -g line info: executed instructions and stall
samples per source line (needs the same build as the profile)."""
import collections
import csv
import re
import subprocess
import sys


def main(rep, disasm, top=40):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[1]
    si, ii, wi = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    prof = []
    for r in rows[2:]:
        if len(r) > max(si, ii, wi):
            prof.append((r[si].strip(), int(r[ii] or 0), int(r[wi] or 0)))
    # nvdisasm listing: sequence of (line marker | instruction)
    cur, seq = ("?", 0), []
    for ln in open(disasm):
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            seq.append((cur, m.group(2).strip()))
    if len(seq) != len(prof):
        print(f"warning: {len(seq)} disasm instrs vs {len(prof)} profiled", file=sys.stderr)
    per = collections.defaultdict(lambda: [0, 0])
    n = min(len(seq), len(prof))
    for k in range(n):
        per[seq[k][0]][0] += prof[k][1]
        per[seq[k][0]][1] += prof[k][2]
    ti = sum(v[0] for v in per.values()) or 1
    ts = sum(v[1] for v in per.values()) or 1
    src_cache = {}
    print(f"total warp-instructions {ti}, stall samples {ts}")
    for (f, l), (i, s) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        try:
            if f not in src_cache:
                src_cache[f] = open(f"pulse_b200/csrc/{f}").read().splitlines()
            text = src_cache[f][l - 1].strip()[:90]
        except Exception:
            text = ""
        print(f"{100*i/ti:5.1f}% inst {100*s/ts:5.1f}% stall  {f}:{l:<4d} {text}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
