"""Batched driver (ragnar_b200/pipeline.py) over several steps x 2 species, with and
without the read of species k+1 overlapped with the compute of species k.

    python tools/bench_pipeline.py [particles_per_species] [nsteps] [workdir]"""
import json
import shutil
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np

from ragnar_b200 import cabi, pipeline

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
work = Path(sys.argv[3] if len(sys.argv) > 3 else "/tmp/rgc_pipeline_bench")
shutil.rmtree(work, ignore_errors=True)
(work / "output" / "prtl").mkdir(parents=True)
cabi.init(0)
pbins = cabi.logspace(1e-3, 1e3, 200)
gbins = cabi.logspace(1e-1, 200, 200)
consts = (0.45 * 0.45 * np.sqrt(100.0) / 2.0, 50.0, (27.0 / 8.0) * 0.1 * 137.0)
t0 = time.perf_counter()
for sp in (1, 2):
    p = cabi.Particles(3).allocate(n)
    p.generate(1, 2000 + sp, 0, 0, n, 1.0, 100.0)
    cols = [p.read(q, d, 0, n) for q in (cabi.Q_U, cabi.Q_E, cabi.Q_B) for d in range(3)]
    p.release()
    for st in range(1, nsteps + 1):
        cabi.tristan_write_species(str(work), st, sp, cols, with_coords=False, append=(sp == 2))
    del cols
t_write = time.perf_counter() - t0
species = [("e-", 1), ("e+", 2)]
steps = list(range(1, nsteps + 1))
pipeline.process_steps(str(work), steps[:1], species, pbins, gbins, *consts)  # warm-up (plans, lanes)
seq = pipeline.process_steps(str(work), steps, species, pbins, gbins, *consts, prefetch=False)
ovl = pipeline.process_steps(str(work), steps, species, pbins, gbins, *consts, prefetch=True,
                             out_file=str(work / "spec.h5"))
same = all(np.array_equal(a.spectrum64, b.spectrum64) and np.array_equal(a.distribution, b.distribution)
           for a, b in zip(seq.results, ovl.results))
nbytes = nsteps * 2 * 9 * 4 * n
print(json.dumps({
    "workload": "batched driver: steps x (e-, e+) read -> energyDistribution -> SynchrotronSpectrum_3D "
                "-> write (legacy/simulation.cpp.bak), page-cache-resident files",
    "particles_per_species": n, "steps": nsteps, "species": 2, "bytes_read": nbytes,
    "write_fixture_s": round(t_write, 2),
    "sequential": {"wall_s": round(seq.wall_s, 4), "read_s": round(seq.read_s, 4),
                   "compute_s": round(seq.compute_s, 4)},
    "overlapped": {"wall_s": round(ovl.wall_s, 4), "read_s": round(ovl.read_s, 4),
                   "compute_s": round(ovl.compute_s, 4)},
    "overlap_of_max(read,compute)": round(max(ovl.read_s, ovl.compute_s) / ovl.wall_s, 3),
    "read_GBps_overlapped": round(nbytes / 1e9 / ovl.wall_s, 2),
    "evals_per_s_end_to_end": nsteps * 2 * n * 200 / ovl.wall_s,
    "identical_results": bool(same)}), flush=True)
shutil.rmtree(work, ignore_errors=True)
sys.exit(0 if same else 1)
