"""Where the module-level end-to-end step spends its time: fromArrays (allocation, zero fill,
staged H2D of pageable columns) for several staging-pool widths.
    RGC_COPY_THREADS=8 python tools/bench_fromarrays.py [n]"""
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np

import ragnar_b200
from ragnar_b200 import cabi

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
rg = ragnar_b200.load()
rg.Initialize()
rng = np.random.default_rng(0)
cols = {f"{q}{d + 1}": rng.random(n, dtype=np.float32) for q in "UEB" for d in range(3)}
os.dup2(2, 1)
for rep in range(3):
    t0 = time.perf_counter()
    p = rg.Particles_3D("e-")
    p.fromArrays(cols)
    t1 = time.perf_counter()
    del p
    cabi.synchronize()
    t2 = time.perf_counter()
    print(f"[fromArrays] n={n} threads={os.environ.get('RGC_COPY_THREADS', 'default')}: fromArrays "
          f"{1e3 * (t1 - t0):.1f} ms ({36 * n / (t1 - t0) / 1e9:.1f} GB/s), release {1e3 * (t2 - t1):.1f} ms",
          file=sys.stderr, flush=True)
# C-ABI pageable write into an existing container (no allocation)
tgt = cabi.Particles(3).allocate(n)
arrs = list(cols.values())
for rep in range(2):
    t0 = time.perf_counter()
    k = 0
    for q in (cabi.Q_U, cabi.Q_E, cabi.Q_B):
        for d in range(3):
            tgt.write(q, d, 0, arrs[k])
            k += 1
    cabi.synchronize()
    t1 = time.perf_counter()
    print(f"[write pageable] {1e3 * (t1 - t0):.1f} ms ({36 * n / (t1 - t0) / 1e9:.1f} GB/s)", file=sys.stderr, flush=True)
t0 = time.perf_counter()
p2 = cabi.Particles(3).allocate(n)
cabi.synchronize()
t1 = time.perf_counter()
p2.release()
cabi.synchronize()
t2 = time.perf_counter()
print(f"[allocate] {1e3 * (t1 - t0):.1f} ms, release {1e3 * (t2 - t1):.1f} ms", file=sys.stderr, flush=True)
