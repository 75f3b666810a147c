"""BASELINE configs[3]: TristanV2_3D plugin on a synthetic HDF5 step with e-/e+ species,
read -> energy histogram + synchrotron spectrum, timed end to end.

    python tools/bench_tristan.py [particles_per_species] [workdir]

The species are generated on the device (Philox, full-3D population), copied to the host
and written with the library's own HDF5 writer (superblock v0, contiguous float32
datasets u_,v_,w_,ex_..bz_<sp>, as Tristan-v2 writes them).  Timed region = what a user
of the plugin runs: readParticles(e-), readParticles(e+) (disk -> pinned lanes ->
device), energyDistribution and SynchrotronSpectrum_3D of both.  Checks: array identity of
sampled slices, results bitwise equal to the same calls on the generated particles."""
import json
import os
import shutil
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np

from ragnar_b200 import cabi


def mem_gb():
    for line in open("/proc/meminfo"):
        if line.startswith("MemAvailable"):
            return int(line.split()[1]) / 1e6
    return 0.0


n_req = int(float(sys.argv[1])) if len(sys.argv) > 1 else 500_000_000
work = Path(sys.argv[2] if len(sys.argv) > 2 else "/tmp/rgc_tristan_bench")
shutil.rmtree(work, ignore_errors=True)
(work / "output" / "prtl").mkdir(parents=True)
disk_gb = shutil.disk_usage(work).free / 1e9
ram_gb = mem_gb()
# one species' host copy (36 B/particle) must fit comfortably, the file (2 species) on disk
n = int(min(n_req, 0.35 * ram_gb * 1e9 / 36, 0.8 * disk_gb * 1e9 / 72))
n -= n % 1024
cabi.init(0)
bins = cabi.logspace(0.01, 1e5, 200)
gbins = cabi.logspace(1e-2, 1e3, 200)
table = cabi.tabulate_ffunc()
consts = (1.0, 1.0, 1.0)

want = {}
samples = {}
t_write = time.perf_counter()
for sp in (1, 2):
    p = cabi.Particles(3).allocate(n)
    p.generate(1, 1000 + sp, 0, 0, n, 1.0, 100.0)
    want[sp] = (cabi.energy_histogram(p, gbins, True, True)[1:],
                cabi.sync_spectrum_particles(p, bins, *consts, table=table)[1])
    cols = [p.read(q, d, 0, n) for q in (cabi.Q_U, cabi.Q_E, cabi.Q_B) for d in range(3)]
    samples[sp] = [c[12345:12345 + 4096].copy() for c in cols] + [c[-777:].copy() for c in cols]
    cabi.tristan_write_species(str(work), 1, sp, cols, with_coords=False, append=(sp == 2))
    del cols
    p.release()
t_write = time.perf_counter() - t_write
fname = work / "output" / "prtl" / "prtl.tot.00001"
file_gb = fname.stat().st_size / 1e9
os.sync()
cold = False
try:
    with open("/proc/sys/vm/drop_caches", "w") as f:
        f.write("3\n")
    cold = True
except OSError:
    pass


def run():
    t0 = time.perf_counter()
    prtls = {}
    for sp in (1, 2):
        prtls[sp], ntotal = cabi.tristan_read_particles(str(work), 1, sp, ignore_coords=True)
        assert ntotal == n and prtls[sp].n == n
    cabi.synchronize()
    t_read = time.perf_counter() - t0
    res = {}
    for sp in (1, 2):
        res[sp] = (cabi.energy_histogram(prtls[sp], gbins, True, True)[1:],
                   cabi.sync_spectrum_particles(prtls[sp], bins, *consts, table=table)[1])
    t_all = time.perf_counter() - t0
    return prtls, res, t_read, t_all


prtls, res, t_read_cold, t_all_cold = run()
ok = True
for sp in (1, 2):
    k = 0
    for q in (cabi.Q_U, cabi.Q_E, cabi.Q_B):
        for d in range(3):
            ok &= np.array_equal(prtls[sp].read(q, d, 12345, 4096), samples[sp][k])
            ok &= np.array_equal(prtls[sp].read(q, d, n - 777, 777), samples[sp][9 + k])
            k += 1
    ok &= np.array_equal(res[sp][0][0], want[sp][0][0]) and np.array_equal(res[sp][1], want[sp][1])
    prtls[sp].release()
prtls, res, t_read_warm, t_all_warm = run()
for sp in (1, 2):
    prtls[sp].release()
read_bytes = 2 * 9 * 4 * n
line = {
    "workload": "TristanV2_3D plugin, BASELINE configs[3]", "particles_per_species": n,
    "requested_particles_per_species": n_req, "species": 2, "file_GB": round(file_gb, 2),
    "host_ram_available_GB": round(ram_gb, 1), "disk_free_GB": round(disk_gb, 1),
    "write_fixture_s": round(t_write, 2),
    "cold_page_cache": cold,
    "read_s_first": round(t_read_cold, 3), "read_GBps_first": round(read_bytes / 1e9 / t_read_cold, 2),
    "end_to_end_s_first": round(t_all_cold, 3),
    "read_s_cached": round(t_read_warm, 3), "read_GBps_cached": round(read_bytes / 1e9 / t_read_warm, 2),
    "end_to_end_s_cached": round(t_all_warm, 3),
    "compute_s": round(t_all_warm - t_read_warm, 4),
    "evals": 2 * n * 200,
    "evals_per_s_end_to_end_cached": 2 * n * 200 / t_all_warm,
    "array_identity_and_bitwise_results": bool(ok),
}
print(json.dumps(line), flush=True)
shutil.rmtree(work, ignore_errors=True)
sys.exit(0 if ok else 1)
