#!/usr/bin/env python
"""Times SynchrotronSpectrum kernels for a list of env settings: tools/sweep_spec.py VAR v1 v2 ..."""
import os, sys, statistics
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ragnar_b200 import cabi
var, vals = sys.argv[1], sys.argv[2:]
n = int(os.environ.get("N", "100000000")); nbins = int(os.environ.get("NBINS", "200"))
cabi.init(0)
lo, hi = (0.01, 1e5) if nbins <= 200 else (1e-3, 1e6)
bins = cabi.logspace(lo, hi, nbins); table = cabi.tabulate_ffunc()
p = cabi.Particles(3).allocate(n); p.generate(int(os.environ.get("KIND", "0")), 123, 0, 0, n, 1.0, 100.0); cabi.synchronize()
for v in vals:
    os.environ[var] = v
    ts = []
    for i in range(4):
        cabi.sync_spectrum_particles(p, bins, 1.0, 1.0, 1.0, table=table)
        ts.append(cabi.last_kernel_times())
    ts = ts[1:]
    print(f"{var}={v}: total {statistics.mean(t[0] for t in ts):.3f} ms  pair {statistics.mean(t[1] for t in ts):.3f} ms  prologue {statistics.mean(t[2] for t in ts):.3f} ms  -> {n*nbins/statistics.mean(t[0] for t in ts)/1e9:.2f} Tevals/s", flush=True)
