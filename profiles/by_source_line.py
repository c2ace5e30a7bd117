#!/usr/bin/env python
"""Aggregates `ncu -i REP --page source --csv --print-source cuda,sass --kernel-name regex:K` by CUDA source line:
share of executed warp-instructions and of stall samples.  Usage: by_source_line.py REP KERNEL_REGEX [TOP]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
data, tot, ts, name = [], 0.0, 0.0, ""
hdr = None
for r in rows:
    if r and r[0] == "Function Name" and not name:
        name = r[1]
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-":      # source-line rows carry "-" in the address column
        continue
    try:
        n = float(r[hdr.index("Instructions Executed")]); s = float(r[hdr.index("# Samples")])
    except ValueError:
        continue
    data.append((n, s, r[0], r[1].strip()[:120])); tot += n; ts += s
print("# %s\n# total warp-instructions %d  samples %d\n# %%inst  %%samples  line  source" % (name[:100], tot, ts))
for n, s, ln, src in sorted(data, key=lambda x: -x[0])[:top]:
    print("%5.1f  %5.1f  %4s  %s" % (100 * n / max(tot, 1), 100 * s / max(ts, 1), ln, src))
