"""Latency of the literal path (rgc_sync_literal.cu) against the hinge pipeline on the same
populations, and FromDist per call:  python tools/bench_literal.py"""
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ragnar_b200 import cabi

cabi.init(0)
table = cabi.tabulate_ffunc()
for n in (2, 4096, 65_536, 524_288):
    p = cabi.Particles(3).allocate(n).generate(1, 7, 0, 0, n, 0.05, 500.0)
    for M, lo, hi in ((200, 0.01, 1e5), (1000, 1e-3, 1e6), (2500, 1e-4, 1e7)):
        bins = cabi.logspace(lo, hi, M)
        row = [f"n={n:>7} M={M:>4}"]
        for mode in ("literal", "hinge"):
            if mode == "hinge":
                os.environ["RGC_LITERAL_MAX_N"] = "0"
            else:
                os.environ.pop("RGC_LITERAL_MAX_N", None)
            cabi.sync_spectrum_particles(p, bins, 1.0, 1.0, 1.0, table=table)
            t0 = time.perf_counter()
            for _ in range(5):
                cabi.sync_spectrum_particles(p, bins, 1.0, 1.0, 1.0, table=table)
            wall = (time.perf_counter() - t0) / 5
            row.append(f"{mode}: call {1e3 * wall:8.3f} ms, kernel {cabi.last_kernel_ms()[1]:8.3f} ms, "
                       f"{n * M / wall:.2e} evals/s")
        print(" | ".join(row), flush=True)
    p.release()
os.environ.pop("RGC_LITERAL_MAX_N", None)
gb = cabi.logspace(1, 100, 200)
fd = cabi.generator_eval(0, [-2.0, 1.0, 100.0], gb)
b1 = cabi.logspace(0.01, 1e7, 200)
cabi.sync_spectrum_dist(gb, fd, True, b1, 1.0, 1.0, table=table)
t0 = time.perf_counter()
for _ in range(50):
    cabi.sync_spectrum_dist(gb, fd, True, b1, 1.0, 1.0, table=table)
print(f"FromDist config 0: {1e6 * (time.perf_counter() - t0) / 50:.1f} us per call, kernel "
      f"{1e3 * cabi.last_kernel_ms()[1]:.1f} us")
