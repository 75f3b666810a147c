"""GPU parity tests proper: the CUDA path, called through the C-ABI, against the
CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): histogram counts bit-exact; spectra and weighted
histograms within 1e-5 relative per bin of the reference's float terms summed in
double (oracle 'f64' outputs == oracle/_ref ragnar_ref64)."""
import numpy as np
import pytest

from tests import synth

pytestmark = pytest.mark.gpu

SPEC_RTOL = 1e-5  # per-bin, on bins >= 1e-6 * max (hinge pipeline: exact interpolant, float pair terms)
# literal path (n <= 2^19 and FromDist): the reference's float term per pair, bit for bit;
# only the fp64 summation order differs from the oracle's
LITERAL_RTOL = 1e-11
HIST_RTOL = 1e-5


@pytest.fixture
def hinge(monkeypatch):
    """force the bucketed hinge pipeline for populations the literal path would take"""
    monkeypatch.setenv("RGC_LITERAL_MAX_N", "0")


@pytest.fixture(params=["literal", "hinge"])
def path_rtol(request, monkeypatch):
    """both evaluation paths of SynchrotronSpectrum_<D>D with their parity bars"""
    if request.param == "hinge":
        monkeypatch.setenv("RGC_LITERAL_MAX_N", "0")
        return SPEC_RTOL
    return LITERAL_RTOL


def _particles(cabi, U, E=None, B=None):
    return cabi.Particles(3).from_columns(U=U, E=E, B=B)


@pytest.mark.parametrize("fourvel", [True, False])
@pytest.mark.parametrize("n", [1, 7, 1000, 2_000_003])
def test_histogram_counts_bit_exact(cabi, port, fourvel, n):
    U = synth.config2(n)
    bins = cabi.logspace(1e-2, 1e3, 200)
    p = _particles(cabi, U)
    hist, counts, _ = cabi.energy_histogram(p, bins, log_spaced=False, fourvel=fourvel)
    _, _, want = port.energy_distribution(*U, bins, log_spaced=False, fourvel=fourvel)
    assert counts.sum() == n
    assert np.array_equal(counts, want)
    assert np.array_equal(hist, want.astype(np.float32))


@pytest.mark.parametrize("fourvel", [True, False])
def test_histogram_weighted(cabi, port, fourvel):
    n = 1_000_000
    U = synth.config2(n, seed=7)
    bins = cabi.logspace(1e-2, 1e3, 200)
    p = _particles(cabi, U)
    hist, counts, s64 = cabi.energy_histogram(p, bins, log_spaced=True, fourvel=fourvel)
    _, want64, want_counts = port.energy_distribution(*U, bins, log_spaced=True, fourvel=fourvel)
    assert np.array_equal(counts, want_counts)
    nz = want64 > 0
    assert np.all(s64[~nz] == 0)
    assert np.max(np.abs(s64[nz] - want64[nz]) / want64[nz]) < HIST_RTOL
    assert np.allclose(hist, want64.astype(np.float32), rtol=HIST_RTOL, atol=0)


def test_histogram_linbins_and_gamma_bins(cabi, port):
    """index formula is logarithmic even for Linbins (reference particles.cpp:239-242)"""
    n = 300_000
    U = synth.config2(n, seed=11, umin=0.05, umax=50)
    p = _particles(cabi, U)
    for bins, fourvel in ((cabi.linspace(0.5, 40, 64), True), (cabi.logspace(1, 60, 33), False)):
        _, counts, _ = cabi.energy_histogram(p, bins, log_spaced=False, fourvel=fourvel)
        _, _, want = port.energy_distribution(*U, bins, log_spaced=False, fourvel=fourvel)
        assert np.array_equal(counts, want)


def test_histogram_edge_values(cabi, port):
    """the survey's probe vector: values on / around the clamps"""
    u = np.array([0.001, 0.5, 1, 5, 999, 1000, 5000, 0.0099999, 0.0], np.float32)
    z = np.zeros_like(u)
    bins = cabi.logspace(1e-2, 1e3, 6)
    p = _particles(cabi, [u, z, z])
    _, counts, _ = cabi.energy_histogram(p, bins, log_spaced=False)
    _, _, want = port.energy_distribution(u, z, z, bins, log_spaced=False)
    assert np.array_equal(counts, want)
    # nactive smaller than the allocation: only the active range is binned
    _, counts5, _ = cabi.energy_histogram(p, bins, log_spaced=False, nactive=5)
    _, _, want5 = port.energy_distribution(u[:5], z[:5], z[:5], bins, log_spaced=False)
    assert np.array_equal(counts5, want5)


@pytest.mark.parametrize("maker,consts", [
    (synth.config3, (1.0, 1.0, 1.0)),
    (synth.full3d, (1.3, 2.0, 0.7)),
    # production-like constants (reference legacy/simulation.cpp.bak:110-118,133-136)
    (synth.full3d, (0.45**2 * 10 / 2, 50.0, (27 / 8) * 0.1 * 137)),
])
@pytest.mark.parametrize("nbins,lo,hi", [(200, 0.01, 1e5), (1000, 1e-3, 1e6), (37, 1e-3, 1e3)])
def test_spectrum_particles(cabi, port, maker, consts, nbins, lo, hi, path_rtol):
    n = 100_000 if nbins <= 200 else 30_000
    U, E, B = maker(n)
    bins = cabi.logspace(lo, hi, nbins)
    p = _particles(cabi, U, E, B)
    s32, s64 = cabi.sync_spectrum_particles(p, bins, *consts)
    _, want = port.sync_spectrum_particles(U, E, B, bins, *consts)
    assert want.max() > 0
    assert synth.rel_err(s64, want) < path_rtol
    assert np.all(s64[want == 0] == 0), "bins the reference leaves at zero must stay zero"
    if path_rtol == LITERAL_RTOL:  # every bin, not only those above 1e-6 of the maximum
        assert np.array_equal(s64 == 0, want == 0)
        assert synth.rel_err(s64, want, floor_frac=0.0) < path_rtol
    assert np.allclose(s32, s64.astype(np.float32), rtol=1e-7, atol=0)


@pytest.mark.parametrize("n", [1, 3, 1023, 1024, 1025, 4099])
def test_spectrum_ragged_sizes(cabi, port, n):
    U, E, B = synth.full3d(n, seed=n)
    bins = cabi.logspace(0.01, 1e5, 200)
    p = _particles(cabi, U, E, B)
    _, s64 = cabi.sync_spectrum_particles(p, bins, 1.0, 1.0, 1.0)
    _, want = port.sync_spectrum_particles(U, E, B, bins, 1.0, 1.0, 1.0)
    assert synth.rel_err(s64, want, floor_frac=0.0) < LITERAL_RTOL
    assert np.array_equal(s64 == 0, want == 0)


@pytest.mark.parametrize("n", [1023, 1024, 1025, 4099])
def test_spectrum_ragged_sizes_hinge(cabi, port, n, hinge):
    """the hinge pipeline on sizes around its tile / piece boundaries; a few thousand
    particles do not average the reference's per-pair rounding: 1e-4 on bins >= 1e-3 max"""
    U, E, B = synth.full3d(n, seed=n)
    bins = cabi.logspace(0.01, 1e5, 200)
    p = _particles(cabi, U, E, B)
    _, s64 = cabi.sync_spectrum_particles(p, bins, 1.0, 1.0, 1.0)
    _, want = port.sync_spectrum_particles(U, E, B, bins, 1.0, 1.0, 1.0)
    assert synth.rel_err(s64, want, floor_frac=1e-3) < 1e-4


def test_spectrum_degenerate_particles(cabi, port):
    """u = 0, B = 0, E-dominated (negative radicand -> NaN -> skipped), huge gamma"""
    U = [np.array([0, 0, 1e4, 3, 0.1, 1e-3], np.float32), np.zeros(6, np.float32),
         np.array([0, 2, 0, 4, 0, 0], np.float32)]
    E = [np.array([0, 0, 0, 5, 0, 0], np.float32), np.zeros(6, np.float32),
         np.array([0, 0, 0, 5, 0, 0], np.float32)]
    B = [np.array([1, 0, 1, 0.1, 1, 1], np.float32), np.array([0, 0, 1, 0, 0, 0], np.float32),
         np.zeros(6, np.float32)]
    bins = cabi.logspace(1e-4, 1e9, 300)
    p = _particles(cabi, U, E, B)
    _, s64 = cabi.sync_spectrum_particles(p, bins, 1.0, 1.0, 1.0)
    _, want = port.sync_spectrum_particles(U, E, B, bins, 1.0, 1.0, 1.0)
    assert np.all(np.isfinite(s64))
    assert synth.rel_err(s64, want, floor_frac=0.0) < LITERAL_RTOL
    assert np.array_equal(s64 == 0, want == 0)


def test_spectrum_unsorted_and_invalid_bins(cabi, port, path_rtol):
    U, E, B = synth.config3(20_000)
    rng = np.random.default_rng(3)
    bins = rng.permutation(cabi.logspace(0.01, 1e5, 100)).astype(np.float32)
    bins[5] = 0.0
    bins[17] = -3.0
    p = _particles(cabi, U, E, B)
    _, s64 = cabi.sync_spectrum_particles(p, bins, 1.0, 1.0, 1.0)
    _, want = port.sync_spectrum_particles(U, E, B, bins, 1.0, 1.0, 1.0)
    assert s64[5] == 0 and s64[17] == 0
    assert synth.rel_err(s64, want) < path_rtol


@pytest.mark.parametrize("case", ["config1", "sync_log", "sync_lin"])
def test_spectrum_from_dist(cabi, port, case):
    if case == "config1":  # BASELINE config 1 / reference README.md:47-60
        gb = cabi.logspace(1, 100, 200)
        f = cabi.generator_eval(0, [-2, 1, 100], gb)
        bins, islog = cabi.logspace(0.01, 1e7, 200), True
    elif case == "sync_log":  # reference src/tests/synchrotron.py:14-35
        gb = cabi.logspace(1, 1000, 200)
        f = cabi.generator_eval(0, [-2.23, 1, 1000], gb)
        bins, islog = cabi.logspace(0.01, 1e7, 200), True
    else:  # reference src/tests/synchrotron.py:38-59
        gb = cabi.linspace(1, 1000, 10000)
        f = cabi.generator_eval(0, [-2.5, 1, 1000], gb)
        bins, islog = cabi.logspace(0.01, 1e6, 500), False
    s32, s64 = cabi.sync_spectrum_dist(gb, f, islog, bins, 1.0, 1.0)
    _, want = port.sync_spectrum_dist(gb, f, islog, bins, 1.0, 1.0)
    # the reference's float term per pair, summed in double in distribution order on both sides
    assert synth.rel_err(s64, want, floor_frac=0.0) < LITERAL_RTOL
    assert np.array_equal(s64 == 0, want == 0)


def test_device_generator_sharding_and_additivity(cabi, port, hinge):
    """Philox particles depend only on (seed, global index): a 2-way split of the
    index range reproduces the single-range spectrum (fp64 round-off) and counts."""
    n = 200_000
    bins = cabi.logspace(1e-3, 1e6, 1000)
    gbins = cabi.logspace(1e-2, 1e3, 200)
    whole = cabi.Particles(3).allocate(n).generate(1, 42, 0, 0, n, 0.05, 500.0)
    _, s_whole = cabi.sync_spectrum_particles(whole, bins, 1.0, 1.0, 1.0)
    _, c_whole, _ = cabi.energy_histogram(whole, gbins, False)
    parts, cparts = np.zeros_like(s_whole), np.zeros_like(c_whole)
    for off, cnt in ((0, 70_001), (70_001, n - 70_001)):
        shard = cabi.Particles(3).allocate(cnt).generate(1, 42, off, 0, cnt, 0.05, 500.0)
        parts += cabi.sync_spectrum_particles(shard, bins, 1.0, 1.0, 1.0)[1]
        cparts += cabi.energy_histogram(shard, gbins, False)[1]
        # shard data are bit-identical to the matching slice of the whole
        for q in (cabi.Q_U, cabi.Q_E, cabi.Q_B):
            for d in range(3):
                assert np.array_equal(shard.read(q, d, 0, cnt), whole.read(q, d, off, cnt))
    assert np.array_equal(cparts, c_whole)
    # per-tile partials are float sums of <= 128 terms before they enter the fp64
    # accumulators; a different split moves the tile boundaries, hence ~1e-8
    assert np.allclose(parts, s_whole, rtol=1e-6, atol=0)
    # and the device-generated sample agrees with the oracle
    cols = [[whole.read(q, d, 0, 20_000) for d in range(3)] for q in (cabi.Q_U, cabi.Q_E, cabi.Q_B)]
    sub = cabi.Particles(3).from_columns(*cols)
    _, s_sub = cabi.sync_spectrum_particles(sub, bins, 1.0, 1.0, 1.0)
    _, want = port.sync_spectrum_particles(*cols, bins, 1.0, 1.0, 1.0)
    assert synth.rel_err(s_sub, want) < SPEC_RTOL


def test_repeatable(cabi, hinge):
    U, E, B = synth.full3d(50_000)
    bins = cabi.logspace(0.01, 1e5, 200)
    p = _particles(cabi, U, E, B)
    a = cabi.sync_spectrum_particles(p, bins, 1.0, 1.0, 1.0)[1]
    b = cabi.sync_spectrum_particles(p, bins, 1.0, 1.0, 1.0)[1]
    assert np.array_equal(a, b), "fixed-order reductions: bitwise reproducible"
    assert cabi.launch_count() > 0
    # the default ranking of the sort kernel is only used after the on-device probe verified
    # that same-address shared atomics of one instruction are served in lane order
    assert cabi.sort_rank_mode() == 1


@pytest.mark.parametrize("nbins,lo,hi", [(200, 0.01, 1e5), (1000, 1e-3, 1e6), (5, 1.0, 10.0),
                                          (254, 1e-2, 1e4), (255, 1e-2, 1e4)])
def test_spectrum_pair_path_matches_gather_path(cabi, port, nbins, lo, hi, monkeypatch, hinge):
    """the bucketed hinge kernel (default) and the gather kernel are two
    formulations of the same sum: they must agree far inside the parity bar"""
    U, E, B = synth.full3d(60_000, seed=5)
    bins = cabi.logspace(lo, hi, nbins)
    p = _particles(cabi, U, E, B)
    monkeypatch.setenv("RGC_SPECTRUM_PATH", "gather")
    _, gather = cabi.sync_spectrum_particles(p, bins, 1.0, 2.0, 3.0)
    monkeypatch.delenv("RGC_SPECTRUM_PATH")
    _, pair = cabi.sync_spectrum_particles(p, bins, 1.0, 2.0, 3.0)
    _, want = port.sync_spectrum_particles(U, E, B, bins, 1.0, 2.0, 3.0)
    assert synth.rel_err(pair, gather) < 2e-6
    assert synth.rel_err(pair, want) < SPEC_RTOL
    assert np.all(pair[want == 0] == 0)


def test_zero_active_particles(cabi):
    """nactive == 0 (a rank that owns no particles): no launch, zeros, no fault"""
    p = cabi.Particles(3).allocate(64)
    p.n = 0
    bins = cabi.logspace(0.01, 1e5, 200)
    s32, s64 = cabi.sync_spectrum_particles(p, bins, 1.0, 1.0, 1.0)
    assert not s32.any() and not s64.any()
    hist, counts, _ = cabi.energy_histogram(p, cabi.logspace(1e-2, 1e3, 50), log_spaced=False)
    assert counts.sum() == 0 and not hist.any()


def test_spectrum_multi_pass(cabi, port, monkeypatch, hinge):
    """populations larger than one pipeline pass (2^27 particles in production; forced
    down to 8192 here): several passes accumulate on the device into the same result"""
    n = 50_000
    U, E, B = synth.full3d(n, seed=5)
    bins = cabi.logspace(0.01, 1e5, 200)
    p = _particles(cabi, U, E, B)
    one = cabi.sync_spectrum_particles(p, bins, 1.3, 2.0, 0.7)[1]
    monkeypatch.setenv("RGC_PAIR_PASS_MAX", "8192")
    many = cabi.sync_spectrum_particles(p, bins, 1.3, 2.0, 0.7)[1]
    monkeypatch.delenv("RGC_PAIR_PASS_MAX")
    _, want = port.sync_spectrum_particles(U, E, B, bins, 1.3, 2.0, 0.7)
    big = want >= 1e-6 * want.max()
    assert np.max(np.abs(many[big] - want[big]) / want[big]) < SPEC_RTOL
    assert np.max(np.abs(many[big] - one[big]) / one[big]) < 1e-6
    assert np.array_equal(many == 0, want == 0)


def test_spectrum_infinite_field_poisons_every_bin(cabi, port, path_rtol):
    """chiR = +inf (an infinite field component): the reference's term is e_syn * inf * F(0)
    = inf * 0 = NaN in every bin (synchrotron.hpp:162-171); NaN inputs are skipped"""
    U, E, B = synth.full3d(5000, seed=8)
    bins = cabi.logspace(0.01, 1e5, 200)
    B[1][123] = np.inf
    got = cabi.sync_spectrum_particles(_particles(cabi, U, E, B), bins, 1.0, 1.0, 1.0)[1]
    _, want = port.sync_spectrum_particles(U, E, B, bins, 1.0, 1.0, 1.0)
    assert np.all(np.isnan(want)) and np.all(np.isnan(got))
    B[1][123] = np.nan  # NaN chiR fails `e_peak > 0`: that particle is skipped
    U[0][7] = np.inf    # beta = inf / inf = NaN: skipped as well
    got = cabi.sync_spectrum_particles(_particles(cabi, U, E, B), bins, 1.0, 1.0, 1.0)[1]
    _, want = port.sync_spectrum_particles(U, E, B, bins, 1.0, 1.0, 1.0)
    assert np.all(np.isfinite(want)) and synth.rel_err(got, want) < path_rtol


def test_sort_rank_variants_agree(cabi, monkeypatch, hinge):
    """RGC_SORT_RANK=ballot (order guaranteed by construction) against the default
    shared-atomic ranking of the sort kernel: same buckets, same sums"""
    U, E, B = synth.full3d(300_000, seed=12)
    bins = cabi.logspace(0.01, 1e5, 200)
    p = _particles(cabi, U, E, B)
    a = cabi.sync_spectrum_particles(p, bins, 1.0, 1.0, 1.0)[1]
    monkeypatch.setenv("RGC_SORT_RANK", "ballot")
    b = cabi.sync_spectrum_particles(p, bins, 1.0, 1.0, 1.0)[1]
    b2 = cabi.sync_spectrum_particles(p, bins, 1.0, 1.0, 1.0)[1]
    monkeypatch.delenv("RGC_SORT_RANK")
    assert np.array_equal(b, b2)
    big = a >= 1e-6 * a.max()
    assert np.max(np.abs(a[big] - b[big]) / a[big]) < 1e-6
    assert np.array_equal(a == 0, b == 0)
    # the hardware replays same-address lanes of one shared-atomic instruction in ascending
    # lane order, which is the order the ballot ranking builds: bit-identical results
    assert np.array_equal(a, b)


@pytest.mark.parametrize("nbins,lo,hi", [(200, 0.01, 1e5), (1000, 1e-3, 1e6)])
def test_zero_group_skipping_is_exact(cabi, monkeypatch, nbins, lo, hi, hinge):
    """lane groups whose bins all lie beyond the table's zero tail for a bucket are not
    evaluated: bit-identical to evaluating every group; and only the bins whose hinge threshold
    lies in a particle's own sub-bucket are evaluated pair by pair at all (about one pair in
    eight, rounded up to whole 32-lane groups), every particle at least once (moment lanes)"""
    U, E, B = synth.config3(400_000, seed=4)
    bins = cabi.logspace(lo, hi, nbins)
    p = _particles(cabi, U, E, B)
    a = cabi.sync_spectrum_particles(p, bins, 1.0, 1.0, 1.0)[1]
    issued = cabi.last_pair_lane_evals()
    monkeypatch.setenv("RGC_PAIR_NO_SKIP", "1")
    b = cabi.sync_spectrum_particles(p, bins, 1.0, 1.0, 1.0)[1]
    issued_all = cabi.last_pair_lane_evals()
    monkeypatch.delenv("RGC_PAIR_NO_SKIP")
    assert np.array_equal(a, b)
    assert 0 < issued <= issued_all
    if nbins > 256:
        assert issued < issued_all  # several groups per sub-bucket: the trailing ones are skipped
    assert 0.9 * 400_000 * 32 <= issued_all < 0.5 * 400_000 * nbins


@pytest.mark.parametrize("phi", ["0", "0.5", "0.999"])
def test_sub_bucket_boundaries_on_thresholds(cabi, port, monkeypatch, hinge, phi):
    """BASELINE's 200 bins over 7 decades on the 200-point table over 8 decades put every hinge
    threshold on a multiple of 1/8 of a cell.  The plan moves the sub-bucket boundaries away from
    them (phase 0.5 there); forcing the phase to 0 puts every threshold ON a boundary, where the
    table nodes' 1e-5 wiggle decides per bucket which neighbouring run a lane group must also
    see (extmask): same spectrum, same exact zeros, more evaluations"""
    U, E, B = synth.config3(300_000, seed=11)
    bins = cabi.logspace(0.01, 1e5, 200)
    p = _particles(cabi, U, E, B)
    base = cabi.sync_spectrum_particles(p, bins, 1.0, 1.0, 1.0)[1]
    issued0 = cabi.last_pair_lane_evals()
    monkeypatch.setenv("RGC_PAIR_PHI", phi)
    got = cabi.sync_spectrum_particles(p, bins, 1.0, 1.0, 1.0)[1]  # (the plan cache is keyed by the knob too)
    issued = cabi.last_pair_lane_evals()
    monkeypatch.delenv("RGC_PAIR_PHI")
    _, want = port.sync_spectrum_particles(U, E, B, bins, 1.0, 1.0, 1.0)
    assert synth.rel_err(got, want) < 1e-5
    assert np.all(got[want == 0] == 0)
    assert synth.rel_err(got, base) < 2e-6
    if phi == "0":
        assert issued > 1.5 * issued0  # lane groups also stream the neighbouring runs


@pytest.mark.parametrize("nbins,lo,hi", [(1, 1.0, 2.0), (2, 0.1, 10.0), (31, 1e-2, 1e4), (33, 1e-2, 1e4),
                                         (257, 1e-3, 1e6), (2032, 1e-3, 1e6), (2500, 1e-3, 1e6)])
def test_hinge_bin_counts(cabi, port, hinge, nbins, lo, hi):
    """lane-group bookkeeping of the sub-bucket plan: a single bin, bins that do not fill a lane
    group, several groups per sub-bucket, the per-launch maximum and beyond it (two launches)"""
    U, E, B = synth.full3d(40_000, seed=5)
    bins = cabi.logspace(lo, hi, nbins) if nbins > 1 else np.array([1.5], np.float32)
    p = _particles(cabi, U, E, B)
    got = cabi.sync_spectrum_particles(p, bins, 1.3, 2.0, 0.7)[1]
    _, want = port.sync_spectrum_particles(U, E, B, bins, 1.3, 2.0, 0.7)
    assert synth.rel_err(got, want, floor_frac=1e-3) < 1e-4  # 4e4 particles do not average the rounding
    assert np.all(got[want == 0] == 0)


def test_config3_at_1e7(cabi, port):
    """SURVEY.md 8d configs 2-3 at 1e7 particles (the oracle needs ~10 s on the box's host
    threads): spectrum <= 1e-5 per bin against the reference's float terms summed in double,
    histogram counts bit-exact"""
    n = 10_000_000
    U, E, B = synth.config3(n, seed=123)
    bins = cabi.logspace(0.01, 1e5, 200)
    p = _particles(cabi, U, E, B)
    _, s64 = cabi.sync_spectrum_particles(p, bins, 1.0, 1.0, 1.0)
    _, want = port.sync_spectrum_particles(U, E, B, bins, 1.0, 1.0, 1.0)
    assert synth.rel_err(s64, want) < SPEC_RTOL
    assert np.array_equal(s64 == 0, want == 0)
    gb = cabi.logspace(1e-2, 1e3, 200)
    _, counts, _ = cabi.energy_histogram(p, gb, log_spaced=False, fourvel=True)
    _, _, want_c = port.energy_distribution(*U, gb, False, True)
    assert counts.sum() == n and np.array_equal(counts, want_c)


@pytest.mark.parametrize("fourvel", [True, False])
@pytest.mark.parametrize("log_spaced", [True, False])
def test_fused_histogram_and_spectrum_is_bit_identical(cabi, port, fourvel, log_spaced, monkeypatch):
    """rgc_hist_and_spectrum (histogram enqueued ahead of the spectrum pipeline, one wait for
    both): bit-identical to the two separate calls on the hinge pipeline, across several
    passes and on the literal path"""
    n = 700_000
    U, E, B = synth.full3d(n, seed=31)
    U[0][:5] = [0.0, 1e-9, 1e9, np.nan, np.inf]  # clamp bins, NaN -> last bin
    p = _particles(cabi, U, E, B)
    gb = cabi.logspace(1e-2, 1e3, 200) if log_spaced else cabi.linspace(0.5, 40, 64)
    bins = cabi.logspace(0.01, 1e5, 200)
    for nactive, passes in ((n, None), (n, "131072"), (5000, None)):
        if passes:
            monkeypatch.setenv("RGC_PAIR_PASS_MAX", passes)
        h32, _, h64 = cabi.energy_histogram(p, gb, log_spaced, fourvel, nactive=nactive)
        s32, s64 = cabi.sync_spectrum_particles(p, bins, 1.3, 2.0, 0.7, nactive=nactive)
        f32, f64, t32, t64 = cabi.hist_and_spectrum(p, gb, log_spaced, fourvel, bins, 1.3, 2.0, 0.7, nactive=nactive)
        if passes:
            monkeypatch.delenv("RGC_PAIR_PASS_MAX")
        assert np.array_equal(f64, h64, equal_nan=True) and np.array_equal(f32, h32, equal_nan=True)
        assert np.array_equal(t64, s64, equal_nan=True) and np.array_equal(t32, s32, equal_nan=True)
        assert f64[np.isfinite(f64)].sum() > 0
    _, want_h, _ = port.energy_distribution(*U, gb, log_spaced, fourvel)
    f32, f64, _, _ = cabi.hist_and_spectrum(p, gb, log_spaced, fourvel, bins, 1.3, 2.0, 0.7)
    ok = np.isfinite(want_h) & (want_h > 0)
    assert np.max(np.abs(f64[ok] - want_h[ok]) / want_h[ok]) < HIST_RTOL
