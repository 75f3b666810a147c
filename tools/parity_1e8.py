"""SURVEY 8d full-size parity, recorded once per round under profiles/: BASELINE configs[1] and
configs[2] at their full 1e8 particles against the reference's own sources.

  * spectrum: the GPU result (hinge pipeline) against oracle/_ref `ragnar_ref64` — the reference's
    unmodified sources, per-pair float terms summed in double — on the SAME 1e8 particles
    (generated on the device, copied back), per-bin relative error on bins >= 1e-6 max;
  * histogram: u64 counts against the reference run in chunks of 1.6e7 particles (its float
    counts stay exact below 2^24) with `log_spaced = False`, bit for bit, and sum = N.

    python tools/parity_1e8.py [n]        (~3 min on 16 host threads)
"""
import contextlib
import io
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np

import oracle
from ragnar_b200 import cabi

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
cabi.init(0)
threads = oracle.port.set_num_threads(len(os.sched_getaffinity(0)))
out = {"particles": n, "host_threads": threads}
ref = oracle.ref64() if oracle.ref_available() else None
out["oracle"] = "oracle/_ref ragnar_ref64 (reference sources, double ScatterView)" if ref else "oracle port (f64 sums)"

# ---- configs[2]: spectrum
p = cabi.Particles(3).allocate(n).generate(0, 123, 0, 0, n, 1.0, 100.0)
bins = cabi.logspace(0.01, 1e5, 200)
t0 = time.perf_counter()
_, got = cabi.sync_spectrum_particles(p, bins, 1.0, 1.0, 1.0)
out["gpu_spectrum_s"] = time.perf_counter() - t0
cols = {f"{q}{d + 1}": p.read(qi, d, 0, n) for q, qi in (("U", cabi.Q_U), ("E", cabi.Q_E), ("B", cabi.Q_B))
        for d in range(3)}
t0 = time.perf_counter()
if ref is not None:
    with contextlib.redirect_stdout(io.StringIO()):
        rp = ref.Particles_3D("e-")
        rp.fromArrays(cols)
        want = ref.SynchrotronSpectrum_3D(rp, ref.Logbins(0.01, 1e5, 200, "mec2"), 1.0, 1.0, 1.0).as_array()
        del rp
    want = want.astype(np.float64)
    got_cmp = got.astype(np.float32).astype(np.float64)  # the reference returns float32
else:
    U, E, B = ([cols[f"{q}{d}"] for d in (1, 2, 3)] for q in "UEB")
    _, want = oracle.port.sync_spectrum_particles(U, E, B, bins, 1.0, 1.0, 1.0)
    got_cmp = got
out["oracle_spectrum_s"] = time.perf_counter() - t0
big = want >= 1e-6 * want.max()
out["spectrum_rel_err_max"] = float(np.max(np.abs(got_cmp[big] - want[big]) / want[big]))
out["spectrum_bins_compared"] = int(big.sum())
out["spectrum_zero_bins_equal"] = bool(np.array_equal(got_cmp == 0, want == 0))
p.release()

# ---- configs[1]: histogram on the contended population (about half of it below the first bin)
p = cabi.Particles(3).allocate(n).generate(2, 123, 0, 0, n, 5e-3, 2e3)
gb = cabi.logspace(1e-2, 1e3, 200)
_, counts, _ = cabi.energy_histogram(p, gb, log_spaced=False, fourvel=True)
want_c = np.zeros(200, np.uint64)
chunk = 16_000_000
t0 = time.perf_counter()
for off in range(0, n, chunk):
    m = min(chunk, n - off)
    u = [p.read(cabi.Q_U, d, off, m) for d in range(3)]
    if ref is not None:
        with contextlib.redirect_stdout(io.StringIO()):
            rp = ref.Particles_3D("e-")
            rp.fromArrays({"U1": u[0], "U2": u[1], "U3": u[2]})
            b = ref.Logbins(1e-2, 1e3, 200)
            b.log_spaced = False
            want_c += rp.energyDistribution(b).F().as_array().astype(np.uint64)
            del rp
    else:
        want_c += oracle.port.energy_distribution(*u, gb, False, True)[2]
out["oracle_histogram_s"] = time.perf_counter() - t0
out["counts_bit_exact"] = bool(np.array_equal(counts, want_c))
out["counts_sum_equals_n"] = bool(int(counts.sum()) == n)
out["counts_in_clamp_bin_0"] = int(counts[0])
out["ok"] = bool(out["spectrum_rel_err_max"] < 1e-5 and out["counts_bit_exact"] and out["counts_sum_equals_n"]
                 and out["spectrum_zero_bins_equal"])
print(json.dumps(out, indent=1))
sys.exit(0 if out["ok"] else 1)
