"""Wall-clock vs device time of the two hot-path calls (host overhead per call)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from ragnar_b200 import cabi

cabi.init(0)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 200
bins = cabi.logspace(0.01, 1e5, nb)
gbins = cabi.logspace(1e-2, 1e3, 200)
table = cabi.tabulate_ffunc()
p = cabi.Particles(3).allocate(n)
p.generate(0, 123, 0, 0, n, 1.0, 100.0)
cabi.synchronize()
for _ in range(3):
    cabi.energy_histogram(p, gbins, True, True, want_counts=False)
    cabi.sync_spectrum_particles(p, bins, 1.0, 1.0, 1.0, table=table)
for name, fn in (("hist", lambda: cabi.energy_histogram(p, gbins, True, True, want_counts=False)),
                 ("spec", lambda: cabi.sync_spectrum_particles(p, bins, 1.0, 1.0, 1.0, table=table))):
    wall, dev = [], []
    for _ in range(10):
        t0 = time.perf_counter()
        fn()
        wall.append(1e3 * (time.perf_counter() - t0))
        dev.append(list(cabi.last_kernel_times()))
    print(name, "wall ms", np.round(np.median(wall), 4), "device [total, main, prologue, -]",
          np.round(np.median(np.array(dev), axis=0), 4))
