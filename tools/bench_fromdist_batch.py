"""Batched SynchrotronSpectrumFromDist: literal terms against the kernel-matrix contraction
(SURVEY 8d last row: is the binned Distribution x kernel-matrix form a real dense contraction?)
    python tools/bench_fromdist_batch.py"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np

from ragnar_b200 import cabi

cabi.init(0)
table = cabi.tabulate_ffunc()
G, M = 200, 200
gb = cabi.logspace(1, 100, G)
base = cabi.generator_eval(0, [-2.0, 1.0, 100.0], gb)
bins = cabi.logspace(0.01, 1e7, M)
print(f"G = {G} distribution bins, M = {M} photon bins; per call: host wall (us) / kernels (us)")
for nbatch in (1, 4, 16, 64, 256, 1024, 4096, 16384):
    fb = np.tile(base, (nbatch, 1)) * np.linspace(1, 2, nbatch, dtype=np.float32)[:, None]
    row = [f"batch {nbatch:>6}"]
    for mode, name in ((0, "literal"), (1, "contraction")):
        cabi.sync_spectrum_dist_batch(gb, fb, True, bins, 1.0, 1.0, table=table, mode=mode)
        reps = 20 if nbatch <= 1024 else 5
        t0 = time.perf_counter()
        for _ in range(reps):
            cabi.sync_spectrum_dist_batch(gb, fb, True, bins, 1.0, 1.0, table=table, mode=mode)
        wall = (time.perf_counter() - t0) / reps
        kern = cabi.last_kernel_ms()[1]
        flops = 2.0 * nbatch * G * M
        row.append(f"{name}: {1e6 * wall:9.1f} / {1e3 * kern:8.1f} us"
                   + (f" ({flops / (kern * 1e-3) / 1e12:.2f} TFLOP/s fp64 incl. K build)" if mode == 1 else
                      f" ({nbatch * G * M / (kern * 1e-3) / 1e9:.1f} G pairs/s)"))
    print(" | ".join(row), flush=True)
