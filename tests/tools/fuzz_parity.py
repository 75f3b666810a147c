"""Randomised parity sweep of the CUDA path against the CPU oracle (beyond the fixed
cases of tests/): random particle counts, bin counts and ranges, constants and
populations (power laws, mono-energetic, zeros / NaNs / infs mixed in).

    python tests/tools/fuzz_parity.py [cases] [seed]

Bars: spectrum <= 1e-5 per bin on bins >= 1e-3 * max and <= 1e-4 on bins in
[1e-6, 1e-3) * max (see two_tier_err); degenerate populations (fewer than 4095 particles,
mono-energetic, 1 % spread: nothing averages the float rounding of the per-particle
coordinate and of runs of identical addends) 1e-4 / 1e-2 (the first bin after the forced
zero F(xmin) = 0 has an unbounded relative slope); exact zeros and NaN-poisoned
results preserved; FromDist <= 1e-5 / 1e-4; histogram counts bit-exact, weighted sums
<= 1e-5; ICSpectrum <= 1e-5.  The fixed-seed BASELINE populations of tests/ meet 1e-5 on
every bin >= 1e-6 * max."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import numpy as np

import oracle
from ragnar_b200 import cabi
from tests import synth



def two_tier_err(got, want, fin):
    """(err on bins >= 1e-3 * max, err on bins in [1e-6, 1e-3) * max): the reference's own
    float rounding of the table coordinate (~3e-6 cell) is amplified without bound next
    to the zeros of the interpolant (first cell after the forced F(xmin) = 0, last cell
    before the zero tail), which is where the smallest bins of a spectrum come from"""
    mx = np.max(np.abs(want[fin]))
    rel = np.abs(got - want) / np.where(want == 0, 1.0, np.abs(want))
    main = fin & (np.abs(want) >= 1e-3 * mx)
    tail = fin & (np.abs(want) >= 1e-6 * mx) & ~main
    return (float(np.max(rel[main])) if main.any() else 0.0,
            float(np.max(rel[tail])) if tail.any() else 0.0)


def exact_fromdist(gb, fd, islog, bins, g_syn, e_at, tx, ty):
    """tabulation.hpp:29-41 / synchrotron.hpp:78-93 evaluated in float64 on the float table"""
    x, y, n = tx.astype(np.float64), ty.astype(np.float64), len(tx)
    e_peak = (np.float32(e_at) * gb * gb / (np.float32(g_syn) * np.float32(g_syn))).astype(np.float64)
    out = np.zeros(len(bins))
    b64 = bins.astype(np.float64)
    with np.errstate(all="ignore"):
        for g in range(len(gb)):
            if not e_peak[g] > 0:
                continue
            x0 = b64 / e_peak[g]
            inside = (x0 >= x[0]) & (x0 < x[-1])
            xi = np.clip(np.floor((n - 1) * np.abs(np.log10(x0 / x[0])) / np.log10(x[-1] / x[0])), 0, n - 2).astype(int)
            F = (y[xi + 1] * np.log10(x0 / x[xi]) + y[xi] * np.log10(x[xi + 1] / x0)) / np.log10(x[xi + 1] / x[xi])
            F = np.where(inside, F, 0.0)
            out += float(fd[g]) * b64 * (float(gb[g]) if islog else 1.0) * F
    return out


ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 150
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
cabi.init(0)
port = oracle.port
worst = 0.0
worst_main = worst_deg = 0.0
edge_bins = 0
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
        emain, etail = two_tier_err(got, want, finite)
        degenerate = kind in ("mono", "narrow") or n < 4095  # no averaging over particles
        worst_main = max(worst_main, emain if not degenerate else 0.0)
        worst_deg = max(worst_deg, emain if degenerate else 0.0)
        if not (err < 1e-5 or (not degenerate and emain < 1e-5 and etail < 1e-4)
                or (degenerate and emain < 1e-4 and etail < 1e-2)):
            ok = False
            j = int(np.argmax(np.where(big, np.abs(got - want) / np.abs(np.where(want == 0, 1, want)), 0)))
            why.append(f"rel err {err:.2e} at bin {j}: got {got[j]:.9e} want {want[j]:.9e} max {np.nanmax(np.abs(want[finite])):.3e}")
            dump = Path(__file__).resolve().parents[2] / "gpurun_out"
            dump.mkdir(exist_ok=True)
            np.savez_compressed(dump / f"fuzz_fail_s{seed}_c{case}.npz", U=np.array(U), E=np.array(E), B=np.array(B),
                                bins=bins, consts=np.array(consts), got=got, want=want)
    if not np.array_equal(got[finite] == 0, want[finite] == 0):
        zg, zw = got[finite] == 0, want[finite] == 0
        mx_all = np.max(np.abs(want[finite])) if finite.any() else 0.0
        stray = max(np.max(np.abs(got[finite][zw]), initial=0), np.max(np.abs(want[finite][zg]), initial=0))
        if stray > 1e-9 * mx_all:
            ok = False
            why.append(f"zero mask differs: got {np.count_nonzero(zg)} zeros, want {np.count_nonzero(zw)}; "
                       f"largest stray value {stray:.3e} (max {mx_all:.3e})")
        else:
            # a bin whose x0 = e_syn / e_peak sits within float rounding of the table's last
            # non-zero node for the (few) particles that reach it: the reference's float x0
            # lands in the zero cell, the fp64 coordinate just before it (or vice versa)
            edge_bins += 1
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
    # weighted histogram (log-spaced bins: sum of 1/energy), fp64 sums of the float terms
    _, wcounts, h64 = cabi.energy_histogram(p, gbins, log_spaced=True, fourvel=fourvel)
    _, want_h64, _ = port.energy_distribution(*U, gbins, True, fourvel)
    with np.errstate(invalid="ignore", divide="ignore"):
        hf = np.isfinite(want_h64)
        if not np.array_equal(np.isfinite(h64), hf) or not np.array_equal(np.isnan(h64), np.isnan(want_h64)):
            ok = False
            why.append("weighted hist finite/NaN mask differs")
        nzh = hf & (want_h64 > 0)
        if nzh.any():
            herr = float(np.max(np.abs(h64[nzh] - want_h64[nzh]) / want_h64[nzh]))
            if not herr < 1e-5:
                ok = False
                why.append(f"weighted hist rel err {herr:.2e}")
        if not np.array_equal(h64[hf] == 0, want_h64[hf] == 0) or not np.array_equal(wcounts, want_c):
            ok = False
            why.append("weighted hist zero mask / counts differ")
    p.release()
    # FromDist and IC on random tabulated distributions (every few cases)
    if case % 3 == 0:
        G = int(rng.choice([1, 2, 33, 200, 1000]))
        glo_d = 10 ** rng.uniform(-1, 2)
        gb = (cabi.logspace if rng.random() < 0.6 else cabi.linspace)(glo_d, glo_d * 10 ** rng.uniform(0.5, 4), G)
        islog = bool(rng.random() < 0.6)
        fd = cabi.generator_eval(0, [float(rng.uniform(-3.5, -1.1)), float(gb.min()), float(gb.max())], gb)
        s_got = cabi.sync_spectrum_dist(gb, fd, islog, bins, consts[1], consts[2])[1]
        _, s_want = port.sync_spectrum_dist(gb, fd, islog, bins, consts[1], consts[2])
        fin = np.isfinite(s_want)
        if fin.any() and np.max(np.abs(s_want[fin])) > 0:
            dmain, dtail = two_tier_err(s_got, s_want, fin)
            if not (dmain < 1e-5 and dtail < 1e-4) or not np.array_equal(s_got[fin] == 0, s_want[fin] == 0):
                # the 200-term float sums of the reference are themselves ~1e-5 noisy next to
                # a zero of F: the fp64 kernel is then judged against the float64 evaluation
                # of the reference formula
                ex = exact_fromdist(gb, fd, islog, bins, consts[1], consts[2], *cabi.tabulate_ffunc())
                keep = fin & (np.abs(s_want) >= 1e-6 * np.max(np.abs(s_want[fin]))) & (ex > 0)
                e_ours = float(np.max(np.abs(s_got[keep] - ex[keep]) / ex[keep]))
                e_ref = float(np.max(np.abs(s_want[keep] - ex[keep]) / ex[keep]))
                if not (e_ours < 1e-6 and np.array_equal(s_got[fin] == 0, s_want[fin] == 0)):
                    ok = False
                    why.append(f"FromDist rel err {dmain:.2e} / tail {dtail:.2e} (G={G}, islog={islog}); vs float64 "
                               f"evaluation of the formula: ours {e_ours:.2e}, reference float {e_ref:.2e}")
        S = int(rng.choice([1, 7, 64, 300]))
        es = np.sort(10 ** rng.uniform(-10, -3, S)).astype(np.float32)
        fs = rng.uniform(0, 1, S).astype(np.float32)
        eic = np.sort(10 ** rng.uniform(-6, 6, min(M, 400))).astype(np.float32)
        i_got = cabi.ic_spectrum(gb, fd, islog, es, fs, eic)[1]
        _, i_want = port.ic_spectrum(gb, fd, islog, es, fs, eic)
        with np.errstate(invalid="ignore"):
            fin = np.isfinite(i_want)
            if not np.array_equal(np.isfinite(i_got), fin):
                ok = False
                why.append("IC finite mask differs")
            elif fin.any() and np.max(np.abs(i_want[fin])) > 0:
                bigi = fin & (np.abs(i_want) >= 1e-6 * np.max(np.abs(i_want[fin])))
                ierr = float(np.max(np.abs(i_got[bigi] - i_want[bigi]) / np.abs(i_want[bigi])))
                if not ierr < 1e-5 or not np.array_equal(i_got[fin] == 0, i_want[fin] == 0):
                    ok = False
                    why.append(f"IC rel err {ierr:.2e}")
    if not ok:
        fails.append((case, n, kind, M, float(lo), float(hi), consts, n_g, fourvel))
        print("FAIL", fails[-1], "|", "; ".join(why), flush=True)
print(f"[fuzz_parity] seed={seed} cases={ncases}: {len(fails)} failures; worst spectrum rel err on bins >= "
      f"1e-3 max: statistical populations {worst_main:.2e}, degenerate {worst_deg:.2e}; on bins >= 1e-6 max: "
      f"{worst:.2e}; cases with a table-edge bin that is 0 on one side and < 1e-9 max on the other: {edge_bins}",
      flush=True)
sys.exit(1 if fails else 0)
