#!/usr/bin/env python
"""Summarise an ncu source-page CSV of one kernel: warp instructions and stall samples
per opcode and per barrier-delimited region (normalised per unit of work)."""
import collections
import csv
import subprocess
import sys

rep, units = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
sel = ["--kernel-name", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else []
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", *sel], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[hdr_i], rows[hdr_i + 1:]
ia, isrc, ismp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
recs = []
for r in data:
    try:
        n, s = int(r[ia]), int(r[ismp])
    except (ValueError, IndexError):
        continue
    toks = r[isrc].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    recs.append((op, r[isrc], n, s))
tot = sum(r[2] for r in recs)
tots = sum(r[3] for r in recs)
print(f"total warp instr {tot:.4g}  per unit {tot / units:.2f}  samples {tots}")
byop, smp = collections.Counter(), collections.Counter()
for op, _, n, s in recs:
    byop[op.split(".")[0]] += n
    smp[op.split(".")[0]] += s
for op, n in byop.most_common(22):
    print(f"  {op:10s} {n / units:8.3f}/unit {100 * n / tot:5.1f}%  samples {100 * smp[op] / tots:5.1f}%")
bars = [i for i, r in enumerate(recs) if "BAR.SYNC" in r[1]]
prev = 0
for b in bars + [len(recs)]:
    n = sum(r[2] for r in recs[prev:b])
    s = sum(r[3] for r in recs[prev:b])
    ops = collections.Counter(r[0].split(".")[0] for r in recs[prev:b] if r[2] > 0)
    top = ",".join(k for k, _ in ops.most_common(4))
    print(f"  region [{prev:5d},{b:5d}) {n / units:7.2f}/unit  samples {100 * s / tots:5.1f}%   {top}")
    prev = b
