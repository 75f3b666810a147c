"""Randomised parity sweep of the CUDA path against the CPU oracle: random particle
counts, bin counts and ranges, constants and populations (power laws, mono-energetic,
1 % spread, zeros / NaNs / infs mixed in).  Shared by tests/test_gpu_fuzz.py (driver-run,
`-m gpu`) and tests/tools/fuzz_parity.py (longer sweeps by hand).

Bars
  literal path (n <= 2^19 particles, and every SynchrotronSpectrumFromDist): the
    reference's float term per pair bit for bit -> every non-zero bin <= 1e-11 (fp64
    summation order), zero / NaN masks identical;
  hinge pipeline (n > 2^19, or forced with RGC_LITERAL_MAX_N=0): <= 1e-5 per bin on bins
    >= 1e-3 * max and <= 1e-4 on bins in [1e-6, 1e-3) * max for statistical populations;
    forced onto degenerate populations (fewer than 4095 particles, mono-energetic, 1 %
    spread: nothing averages the reference's own float rounding of the table coordinate)
    1e-4 / 1e-2 — the bar of round 1, kept to watch the pipeline, not the product path;
  histogram counts bit-exact, weighted sums <= 1e-5; ICSpectrum <= 1e-5."""
import os

import numpy as np

from tests import synth

LITERAL_MAX_N = 1 << 19
LITERAL_RTOL = 1e-11


def two_tier_err(got, want, fin):
    """(err on bins >= 1e-3 * max, err on bins in [1e-6, 1e-3) * max)"""
    mx = np.max(np.abs(want[fin]))
    rel = np.abs(got - want) / np.where(want == 0, 1.0, np.abs(want))
    main = fin & (np.abs(want) >= 1e-3 * mx)
    tail = fin & (np.abs(want) >= 1e-6 * mx) & ~main
    return (float(np.max(rel[main])) if main.any() else 0.0,
            float(np.max(rel[tail])) if tail.any() else 0.0)


def make_population(rng, n, kind):
    if kind == "config3":
        U, E, B = synth.config3(n, seed=int(rng.integers(1 << 30)))
    else:
        U, E, B = synth.full3d(n, seed=int(rng.integers(1 << 30)))
    if kind == "mono":  # every particle identical
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
    return U, E, B


def check_spectrum(got, want, literal, degenerate):
    """-> list of failure strings (empty = pass), worst relative error on bins >= 1e-6 max"""
    why = []
    finite = np.isfinite(want)
    if not np.array_equal(np.isfinite(got), finite) or not np.array_equal(np.isnan(got), np.isnan(want)):
        why.append(f"finite mask differs ({np.count_nonzero(~np.isfinite(got))} vs "
                   f"{np.count_nonzero(~finite)} non-finite)")
        return why, 0.0
    if not finite.any() or np.max(np.abs(want[finite])) == 0:
        if finite.any() and np.max(np.abs(got[finite])) != 0:
            why.append("reference is identically zero, result is not")
        return why, 0.0
    mx = np.max(np.abs(want[finite]))
    nz = finite & (want != 0)
    rel = np.zeros_like(want)
    rel[nz] = np.abs(got[nz] - want[nz]) / np.abs(want[nz])
    big = finite & (np.abs(want) >= 1e-6 * mx)
    err = float(np.max(rel[big]))
    if literal:
        if not np.array_equal(got[finite] == 0, want[finite] == 0):
            why.append("zero mask differs on the literal path")
        worst_all = float(np.max(rel[nz]))
        if not worst_all < LITERAL_RTOL:
            j = int(np.argmax(rel))
            why.append(f"literal path rel err {worst_all:.2e} at bin {j}: got {got[j]:.17e} want {want[j]:.17e}")
        return why, err
    emain, etail = two_tier_err(got, want, finite)
    if not (err < 1e-5 or (not degenerate and emain < 1e-5 and etail < 1e-4)
            or (degenerate and emain < 1e-4 and etail < 1e-2)):
        j = int(np.argmax(np.where(big, rel, 0)))
        why.append(f"rel err {err:.2e} at bin {j}: got {got[j]:.9e} want {want[j]:.9e} max {mx:.3e}")
    zg, zw = got[finite] == 0, want[finite] == 0
    if not np.array_equal(zg, zw):
        stray = max(np.max(np.abs(got[finite][zw]), initial=0), np.max(np.abs(want[finite][zg]), initial=0))
        # a bin whose x0 = e_syn / e_peak sits within float rounding of the table's last
        # non-zero node: the reference's float x0 lands in the zero cell, the fp64
        # coordinate just before it (or vice versa)
        if stray > 1e-9 * mx:
            why.append(f"zero mask differs: largest stray value {stray:.3e} (max {mx:.3e})")
    return why, err


def run(cabi, port, ncases, seed, sizes=None, dump_dir=None, log=print):
    """-> (failures, stats)"""
    rng = np.random.default_rng(seed)
    sizes = sizes or [1, 2, 31, 100, 4095, 4096, 4097, 8192, 20_000, 65_537, 150_000, 400_000,
                      600_000, 1_000_000]
    stats = {"literal_cases": 0, "hinge_cases": 0, "worst_hinge_stat": 0.0, "worst_hinge_degenerate": 0.0,
             "worst_literal": 0.0, "fromdist_cases": 0}
    fails = []
    for case in range(ncases):
        n = int(rng.choice(sizes))
        kind = str(rng.choice(["config3", "full3d", "mono", "dirty", "narrow"]))
        U, E, B = make_population(rng, n, kind)
        M = int(rng.choice([1, 2, 5, 37, 200, 254, 255, 500, 1000, 2033, 2500]))
        if n > LITERAL_MAX_N:
            M = min(M, 1000)  # bounds the oracle's time (n x M pairs on the host)
        lo = 10 ** rng.uniform(-6, 1)
        hi = lo * 10 ** rng.uniform(0.5, 9)
        bins = cabi.logspace(lo, hi, M) if rng.random() < 0.8 else cabi.linspace(lo, hi, M)
        consts = (float(10 ** rng.uniform(-1, 1)), float(10 ** rng.uniform(-0.5, 2)),
                  float(10 ** rng.uniform(-2, 2)))
        force_hinge = n <= LITERAL_MAX_N and n >= 4095 and rng.random() < 0.25
        literal = n <= LITERAL_MAX_N and not force_hinge
        p = cabi.Particles(3).from_columns(U=U, E=E, B=B)
        if force_hinge:
            os.environ["RGC_LITERAL_MAX_N"] = "0"
        try:
            _, got = cabi.sync_spectrum_particles(p, bins, *consts)
        finally:
            os.environ.pop("RGC_LITERAL_MAX_N", None)
        _, want = port.sync_spectrum_particles(U, E, B, bins, *consts)
        degenerate = kind in ("mono", "narrow") or n < 4095
        why, err = check_spectrum(got, want, literal, degenerate)
        if literal:
            stats["literal_cases"] += 1
            stats["worst_literal"] = max(stats["worst_literal"], err)
        else:
            stats["hinge_cases"] += 1
            key = "worst_hinge_degenerate" if degenerate else "worst_hinge_stat"
            fin = np.isfinite(want)
            if fin.any() and np.max(np.abs(want[fin])) > 0 and np.array_equal(np.isfinite(got), fin):
                stats[key] = max(stats[key], two_tier_err(got, want, fin)[0])
        if why and dump_dir is not None:
            os.makedirs(dump_dir, exist_ok=True)
            np.savez_compressed(os.path.join(dump_dir, f"fuzz_fail_s{seed}_c{case}.npz"), U=np.array(U),
                                E=np.array(E), B=np.array(B), bins=bins, consts=np.array(consts),
                                got=got, want=want)
        # histogram on the same particles
        n_g = int(rng.choice([1, 2, 6, 50, 200, 777]))
        glo = 10 ** rng.uniform(-3, 0.5)
        gbins = cabi.logspace(glo, glo * 10 ** rng.uniform(0.3, 6), n_g)
        fourvel = bool(rng.random() < 0.5)
        _, counts, _ = cabi.energy_histogram(p, gbins, log_spaced=False, fourvel=fourvel)
        _, _, want_c = port.energy_distribution(*U, gbins, False, fourvel)
        if not np.array_equal(counts, want_c):
            why.append(f"hist counts differ in {np.count_nonzero(counts != want_c)} bins")
        # weighted histogram (log-spaced bins: sum of 1/energy), fp64 sums of the float terms
        _, wcounts, h64 = cabi.energy_histogram(p, gbins, log_spaced=True, fourvel=fourvel)
        _, want_h64, _ = port.energy_distribution(*U, gbins, True, fourvel)
        with np.errstate(invalid="ignore", divide="ignore"):
            hf = np.isfinite(want_h64)
            if not np.array_equal(np.isfinite(h64), hf) or not np.array_equal(np.isnan(h64), np.isnan(want_h64)):
                why.append("weighted hist finite/NaN mask differs")
            nzh = hf & (want_h64 > 0)
            if nzh.any():
                herr = float(np.max(np.abs(h64[nzh] - want_h64[nzh]) / want_h64[nzh]))
                if not herr < 1e-5:
                    why.append(f"weighted hist rel err {herr:.2e}")
            if not np.array_equal(h64[hf] == 0, want_h64[hf] == 0) or not np.array_equal(wcounts, want_c):
                why.append("weighted hist zero mask / counts differ")
        p.release()
        # FromDist and IC on random tabulated distributions (every few cases)
        if case % 3 == 0:
            stats["fromdist_cases"] += 1
            G = int(rng.choice([1, 2, 33, 200, 1000]))
            glo_d = 10 ** rng.uniform(-1, 2)
            gb = (cabi.logspace if rng.random() < 0.6 else cabi.linspace)(glo_d, glo_d * 10 ** rng.uniform(0.5, 4), G)
            islog = bool(rng.random() < 0.6)
            fd = cabi.generator_eval(0, [float(rng.uniform(-3.5, -1.1)), float(gb.min()), float(gb.max())], gb)
            s_got = cabi.sync_spectrum_dist(gb, fd, islog, bins, consts[1], consts[2])[1]
            _, s_want = port.sync_spectrum_dist(gb, fd, islog, bins, consts[1], consts[2])
            dwhy, _ = check_spectrum(s_got, s_want, True, False)
            why += [f"FromDist (G={G}, islog={islog}): {w}" for w in dwhy]
            S = int(rng.choice([1, 7, 64, 300]))
            es = np.sort(10 ** rng.uniform(-10, -3, S)).astype(np.float32)
            fs = rng.uniform(0, 1, S).astype(np.float32)
            eic = np.sort(10 ** rng.uniform(-6, 6, min(M, 400))).astype(np.float32)
            i_got = cabi.ic_spectrum(gb, fd, islog, es, fs, eic)[1]
            _, i_want = port.ic_spectrum(gb, fd, islog, es, fs, eic)
            with np.errstate(invalid="ignore"):
                fin = np.isfinite(i_want)
                if not np.array_equal(np.isfinite(i_got), fin):
                    why.append("IC finite mask differs")
                elif fin.any() and np.max(np.abs(i_want[fin])) > 0:
                    bigi = fin & (np.abs(i_want) >= 1e-6 * np.max(np.abs(i_want[fin])))
                    ierr = float(np.max(np.abs(i_got[bigi] - i_want[bigi]) / np.abs(i_want[bigi])))
                    if not ierr < 1e-5 or not np.array_equal(i_got[fin] == 0, i_want[fin] == 0):
                        why.append(f"IC rel err {ierr:.2e}")
        if why:
            fails.append((case, n, kind, M, float(lo), float(hi), consts, n_g, fourvel, "; ".join(why)))
            log("FAIL", fails[-1])
    return fails, stats
