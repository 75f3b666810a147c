"""Bitwise run-to-run reproducibility of the spectrum (stress for the atomic-rank sort)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from ragnar_b200 import cabi

cabi.init(0)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
bins = cabi.logspace(0.01, 1e5, 200)
table = cabi.tabulate_ffunc()
for kind in (0, 1):
    p = cabi.Particles(3).allocate(n)
    p.generate(kind, 123, 0, 0, n, 1.0, 100.0)
    ref = cabi.sync_spectrum_particles(p, bins, 1.0, 1.0, 1.0, table=table)[1]
    bad = 0
    for _ in range(reps):
        got = cabi.sync_spectrum_particles(p, bins, 1.0, 1.0, 1.0, table=table)[1]
        bad += int(not np.array_equal(got, ref))
    print(f"kind {kind}: {reps} repeats, {bad} differ bitwise")
    p.release()
