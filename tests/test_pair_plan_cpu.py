"""CPU-side checks of the hinge path's plan (rgc_pair_plan_describe, host-only): how photon bins
are laid out in 32-lane groups by the sub-bucket of their hinge threshold (DESIGN.md 3.2,
ragnar_b200/csrc/rgc_sync_pair.cu make_pair_plan).  The reference has no counterpart: its
MDRange visits every (particle, bin) pair (src/physics/synchrotron.cpp:124-139)."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def cabi_cpu():
    from ragnar_b200 import cabi

    cabi.lib()
    return cabi


@pytest.fixture(scope="module")
def ftable(cabi_cpu):
    return cabi_cpu.tabulate_ffunc()  # the reference's 200-point table over 8 decades


def _describe(cabi, bins, ftable):
    return cabi.pair_plan_describe(bins, *ftable)


def test_every_bin_has_exactly_one_lane(cabi_cpu, ftable):
    for nbins, lo, hi in ((200, 0.01, 1e5), (1000, 1e-3, 1e6), (37, 1e-3, 1e3), (1, 1.0, 2.0), (2032, 1e-3, 1e6)):
        bins = cabi_cpu.logspace(lo, hi, nbins) if nbins > 1 else np.array([1.5], np.float32)
        d = _describe(cabi_cpu, bins, ftable)
        assert d["eligible"] == 1
        assert d["slots"] == 32 * d["groups"]
        sb = d["slot_bin"]
        assert sorted(sb[sb >= 0].tolist()) == list(range(nbins))
        # every sub-bucket owns at least one lane group whose first two lanes are the moment lanes
        assert d["groups"] >= d["sub_buckets"] == 9
        assert np.all(sb.reshape(-1, 32)[:, :2][sb.reshape(-1, 32)[:, 0] == -1] == -1)
        assert 0.0 <= d["phase"] < 1.0


def test_baseline_bins_sit_mid_sub_bucket(cabi_cpu, ftable):
    """BASELINE's 200 bins over 7 decades on the 200-point table over 8 decades: the bin spacing is
    7/8 of a table cell, so every hinge threshold is a multiple of 1/8 of a cell; the plan must put
    the sub-bucket boundaries half-way between them, and no lane group may need a neighbouring run"""
    d = _describe(cabi_cpu, cabi_cpu.logspace(0.01, 1e5, 200), ftable)
    assert abs(d["phase"] - 0.5) < 0.02
    assert d["buckets_with_extension"] == 0
    assert d["most_groups_per_sub_bucket"] == 1  # ~25 bins per sub-bucket: one pair in eight is evaluated
    assert d["groups"] == 9


def test_forced_phase_puts_thresholds_on_boundaries(cabi_cpu, ftable, monkeypatch):
    monkeypatch.setenv("RGC_PAIR_PHI", "0")
    d = _describe(cabi_cpu, cabi_cpu.logspace(0.01, 1e5, 200), ftable)
    assert d["phase"] == 0.0
    assert d["eligible"] == 1 and d["buckets_with_extension"] > 0  # neighbouring runs, still exact


def test_group_counts_scale_with_bins(cabi_cpu, ftable):
    d1000 = _describe(cabi_cpu, cabi_cpu.logspace(1e-3, 1e6, 1000), ftable)
    assert 4 <= d1000["most_groups_per_sub_bucket"] <= 6  # ~125 bins per sub-bucket
    assert d1000["buckets_with_extension"] == 0
    d2032 = _describe(cabi_cpu, cabi_cpu.logspace(1e-3, 1e6, 2032), ftable)
    assert d2032["groups"] <= 96 and d2032["eligible"] == 1


def test_random_bins_and_rejections(cabi_cpu, ftable):
    rng = np.random.default_rng(7)
    for _ in range(20):
        n = int(rng.integers(1, 400))
        bins = np.sort(10 ** rng.uniform(-3, 5, n)).astype(np.float32)
        d = _describe(cabi_cpu, bins, ftable)
        sb = d["slot_bin"]
        assert d["eligible"] == 1 and sorted(sb[sb >= 0].tolist()) == list(range(n))
    # a table that does not vanish at both ends is not the hinge path's (gather kernel instead)
    tx, ty = ftable
    ty2 = ty.copy()
    ty2[-1] = 1.0
    assert cabi_cpu.pair_plan_describe(cabi_cpu.logspace(0.01, 1e5, 200), tx, ty2)["eligible"] == 0
    # more bins than one launch takes
    assert _describe(cabi_cpu, cabi_cpu.logspace(1e-3, 1e6, 2500), ftable)["eligible"] == 0


def test_sub_bucket_decomposition_identity():
    """the identity the pair kernel rests on (rgc_sync_pair.cu header), in numpy fp64: for a bucket's
    particles (fc_i, w_i) and a bin with threshold t, sum_i w_i max(0, fc_i - t) equals the pair sum over
    the ONE sub-bucket s = floor(8 fc + phi) that contains t plus S1_s - t S0_s of the sub-buckets above
    it; and for the cell next to the table's zero tail, sum_i w_i max(0, t - fc_i) equals
    S0_run - sum w sat(fc + (1 - t)) on that sub-bucket plus t S0_s - S1_s of the sub-buckets below"""
    rng = np.random.default_rng(3)
    for _ in range(50):
        n = int(rng.integers(1, 5000))
        fc = rng.random(n)
        w = rng.random(n) * 10 ** rng.uniform(-3, 3)
        phi = rng.random()
        s = np.minimum(np.floor(8 * fc + phi).astype(int), 8)
        S0 = np.bincount(s, weights=w, minlength=9)
        S1 = np.bincount(s, weights=w * fc, minlength=9)
        for t in rng.random(8):
            s0 = min(int(np.floor(8 * t + phi)), 8)
            own = s == s0
            # L form: hinge active above the threshold
            direct = np.sum(w * np.maximum(0.0, fc - t))
            pair = np.sum(w[own] * np.clip(fc[own] - t, 0.0, 1.0))
            lin = np.sum(S1[s0 + 1:] - t * S0[s0 + 1:])
            assert abs(direct - (pair + lin)) <= 1e-12 * max(1.0, abs(direct))
            # zero-tail form: S0 of the run minus the saturated sum, moments of the sub-buckets below
            direct_r = np.sum(w * np.maximum(0.0, t - fc))
            pair_r = np.sum(w[own]) - np.sum(w[own] * np.clip(fc[own] + (1.0 - t), 0.0, 1.0))
            lin_r = np.sum(t * S0[:s0] - S1[:s0])
            assert abs(direct_r - (pair_r + lin_r)) <= 1e-12 * max(1.0, abs(direct_r))
