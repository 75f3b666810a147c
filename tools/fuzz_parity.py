"""Randomised parity sweep of the CUDA path against the CPU oracle (beyond the fixed
cases of tests/): random particle counts, bin counts and ranges, constants and
populations (power laws, mono-energetic, zeros / NaNs / infs mixed in).

    python tools/fuzz_parity.py [cases] [seed]

Bars: spectrum <= 1e-5 per bin on bins >= 1e-6 * max (1e-4 for the mono-energetic
population, see below) with exact zeros and NaN-poisoned results preserved; histogram
counts bit-exact."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np

import oracle
from ragnar_b200 import cabi
from tests import synth

ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 150
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
cabi.init(0)
port = oracle.port
worst = 0.0
fails = []
for case in range(ncases):
    n = int(rng.choice([1, 2, 31, 4095, 4096, 4097, 8192, 20_000, 65_537, 150_000, 400_000]))
    kind = rng.choice(["config3", "full3d", "mono", "dirty", "narrow"])
    if kind == "config3":
        U, E, B = synth.config3(n, seed=int(rng.integers(1 << 30)))
    else:
        U, E, B = synth.full3d(n, seed=int(rng.integers(1 << 30)))
    if kind == "mono":  # every particle in one bucket
        U = [np.full(n, 30.0, np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)]
        E = [np.zeros(n, np.float32)] * 3
        B = [np.zeros(n, np.float32), np.full(n, 1.0, np.float32), np.zeros(n, np.float32)]
    if kind == "narrow":  # a few adjacent buckets
        U[0] = (30.0 * (1 + 0.01 * rng.random(n))).astype(np.float32)
    if kind == "dirty" and n > 8:
        for arr in (U[0], B[1], E[2]):
            idx = rng.integers(0, n, max(1, n // 50))
            arr[idx] = rng.choice(np.array([0.0, np.nan, np.inf, -np.inf, 1e-30, 1e30, -5.0], np.float32),
                                  len(idx))
    M = int(rng.choice([1, 2, 5, 37, 200, 254, 255, 500, 1000, 2033, 2500]))
    lo = 10 ** rng.uniform(-6, 1)
    hi = lo * 10 ** rng.uniform(0.5, 9)
    bins = cabi.logspace(lo, hi, M) if rng.random() < 0.8 else cabi.linspace(lo, hi, M)
    consts = (float(10 ** rng.uniform(-1, 1)), float(10 ** rng.uniform(-0.5, 2)),
              float(10 ** rng.uniform(-2, 2)))
    p = cabi.Particles(3).from_columns(U=U, E=E, B=B)
    _, got = cabi.sync_spectrum_particles(p, bins, *consts)
    _, want = port.sync_spectrum_particles(U, E, B, bins, *consts)
    ok = True
    why = []
    finite = np.isfinite(want)
    if not np.array_equal(np.isfinite(got), finite):
        ok = False
        why.append(f"finite mask differs ({np.count_nonzero(~np.isfinite(got))} vs {np.count_nonzero(~finite)} non-finite)")
    if finite.any() and np.nanmax(np.abs(want[finite])) > 0:
        big = finite & (np.abs(want) >= 1e-6 * np.nanmax(np.abs(want[finite])))
        err = float(np.max(np.abs(got[big] - want[big]) / np.abs(want[big])))
        worst = max(worst, err)
        # a mono-energetic population puts every particle at ONE table coordinate: the
        # reference's own float rounding of log10f(x0) (~3e-6 cell) is then not averaged
        # and shows, in the steep tail of F just above the 1e-6 floor, as up to ~2e-5
        tol = 1e-4 if (kind == "mono" or n <= 2) else 1e-5
        if not err < tol:
            ok = False
            j = int(np.argmax(np.where(big, np.abs(got - want) / np.abs(np.where(want == 0, 1, want)), 0)))
            why.append(f"rel err {err:.2e} at bin {j}: got {got[j]:.9e} want {want[j]:.9e} max {np.nanmax(np.abs(want[finite])):.3e}")
    if not np.array_equal(got[finite] == 0, want[finite] == 0):
        ok = False
        zg, zw = got[finite] == 0, want[finite] == 0
        why.append(f"zero mask differs: got {np.count_nonzero(zg)} zeros, want {np.count_nonzero(zw)}; "
                   f"largest |got| where want==0: {np.max(np.abs(got[finite][zw]), initial=0):.3e}; "
                   f"largest |want| where got==0: {np.max(np.abs(want[finite][zg]), initial=0):.3e}")
    # histogram on the same particles
    n_g = int(rng.choice([1, 2, 6, 50, 200, 777]))
    glo = 10 ** rng.uniform(-3, 0.5)
    gbins = cabi.logspace(glo, glo * 10 ** rng.uniform(0.3, 6), n_g)
    fourvel = bool(rng.random() < 0.5)
    _, counts, _ = cabi.energy_histogram(p, gbins, log_spaced=False, fourvel=fourvel)
    _, _, want_c = port.energy_distribution(*U, gbins, False, fourvel)
    if not np.array_equal(counts, want_c):
        ok = False
        why.append(f"hist counts differ in {np.count_nonzero(counts != want_c)} bins")
    p.release()
    if not ok:
        fails.append((case, n, kind, M, float(lo), float(hi), consts, n_g, fourvel))
        print("FAIL", fails[-1], "|", "; ".join(why), flush=True)
print(f"[fuzz_parity] seed={seed} cases={ncases}: {len(fails)} failures, worst spectrum rel err {worst:.2e}",
      flush=True)
sys.exit(1 if fails else 0)
