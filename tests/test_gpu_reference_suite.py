"""The reference's own pytest suite (src/tests/*.py @ fceb6b08), restated against
the drop-in `ragnar` module: same calls, same property checks and tolerances.
The reference files cannot be read on the GPU box, so the checks are written out
here, each citing the test it mirrors."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def check_slope(xs, ys, p, xmin, xmax, atol=1e-2):
    mask = (xs > xmin) & (xs < xmax)
    pfit, _ = np.polyfit(np.log10(xs[mask]), np.log10(ys[mask]), 1)
    return np.isclose(p, pfit, atol=atol)


def test_arrays_bins(rg):  # src/tests/arrays_bins.py:7-28
    array = np.arange(123)
    assert np.allclose(rg.Array1D_i(array.astype(int)).as_array(), array)
    assert np.allclose(rg.Array1D_f(array.astype(np.float32)).as_array(), array)
    assert np.allclose(rg.Array1D_d(array.astype(np.float64)).as_array(), array)
    assert rg.Array1D_d(array.astype(np.float64)).extent() == 123
    bins = rg.Bins(np.logspace(-2.5, 2.5, 213))
    bins.log_spaced = True
    bins.unit = rg.EnergyUnits.mec2
    assert bins.unit == rg.EnergyUnits.mec2
    assert bins.extent() == 213
    assert bins.log_spaced
    assert repr(rg.Array1D_f(np.zeros(5, np.float32))) == "1D Array [ size: 5 ]"
    with pytest.raises(ValueError):  # std::range_error -> ValueError
        rg.Array1D_f(np.zeros(5, np.float32)).head(10)


def test_linspace_logspace(rg):  # src/tests/linspace_logspace.py:7-36
    assert np.allclose(rg.Linspace(-32.0, 4832.0, 56).as_array(), np.linspace(-32.0, 4832.0, 56))
    assert np.allclose(rg.Logspace(10**-2.5, 10**2.5, 213).as_array(), np.logspace(-2.5, 2.5, 213))
    linbins = rg.Linbins(55.0, 56.0, 123)
    assert np.allclose(linbins.as_array(), np.linspace(55.0, 56.0, 123))
    assert not linbins.log_spaced
    logbins = rg.Logbins(1, 1e3, 22)
    assert np.allclose(logbins.as_array(), np.logspace(0, 3, 22))
    assert logbins.log_spaced
    with pytest.raises(RuntimeError, match="Linspace start must be < stop"):
        rg.Linspace(2.0, 1.0, 4)
    with pytest.raises(RuntimeError, match="strictly positive"):
        rg.Logspace(0.0, 1.0, 4)


def test_distributions(rg):  # src/tests/distributions.py:7-29
    trapz = getattr(np, "trapezoid", None) or np.trapz
    for norm in [rg.Logspace, rg.Linspace]:
        bins = rg.Bins(norm(1e-2, 1, int(1e6)))
        dist = rg.TabulatedDistribution(bins, rg.DeltaGenerator(2e-2, 0.01))
        assert np.isclose(trapz(dist.F().as_array(), dist.EnergyBins().as_array()), 1)
    for norm in [rg.Logspace, rg.Linspace]:
        bins = rg.Bins(norm(1e-3, 10, int(1e6)))
        dist = rg.TabulatedDistribution(bins, rg.BrokenPlawGenerator(0.3, 0.23, -1.0, 1e-2, 2))
        assert np.isclose(trapz(dist.F().as_array(), dist.EnergyBins().as_array()), 1)
    for norm in [rg.Logspace, rg.Linspace]:
        bins = rg.Bins(norm(1e-2, 1, int(1e6)))
        dist = rg.TabulatedDistribution(bins, rg.PlawGenerator(-1.2, 1e-2, 1))
        assert np.isclose(trapz(dist.F().as_array(), dist.EnergyBins().as_array()), 1)
    with pytest.raises(RuntimeError, match="normalization diverges"):
        rg.PlawGenerator(-2.0)


KEYS = ["U1", "U3", "E1", "E2", "E3", "B1", "B2", "B3"]


def _cols(p):
    return [p.U(1), p.U(3), p.E(1), p.E(2), p.E(3), p.B(1), p.B(2), p.B(3)]


def test_prtls(rg):  # src/tests/particles.py:7-99
    prtls = rg.Particles_3D("electrons")
    assert prtls.label() == "electrons"
    assert not prtls.is_allocated()
    rng = np.random.default_rng(123)
    first = {k: rng.random(10) for k in KEYS}
    prtls.fromArrays(first)
    for k, f in zip(KEYS, _cols(prtls)):
        assert np.allclose(first[k], f.as_array())
    assert prtls.is_allocated() and prtls.nactive() == 10 and prtls.nalloc() == 10
    with pytest.raises(RuntimeError, match="already allocated"):
        prtls.fromArrays(first)
    second = {k: rng.random(10) for k in KEYS}
    prtls.fromArrays(second, append=True)
    for k, f in zip(KEYS, _cols(prtls)):
        assert np.allclose(first[k], f.as_array()[:10])
        assert np.allclose(second[k], f.as_array()[10:])
    assert prtls.nactive() == 20 and prtls.nalloc() == 20 and len(prtls) == 20
    assert repr(prtls) == "Particles<3D> (electrons) : 2.00·10^1"
    with pytest.raises(IndexError):  # std::out_of_range("Invalid component")
        prtls.U(4)
    with pytest.raises(RuntimeError, match="Inconsistent number of particles"):
        rg.Particles_3D("x").fromArrays({"U1": np.zeros(3), "U2": np.zeros(4)})
    with pytest.raises(RuntimeError, match="No particles provided"):
        rg.Particles_3D("x").fromArrays({})


def test_prtls_preallocated(rg):  # src/tests/particles.py:102-190
    rng = np.random.default_rng(123)
    prtls = rg.Particles_3D("electrons")
    prtls.allocate(20)
    assert prtls.nactive() == 0 and prtls.nalloc() == 20
    first = {k: rng.random(10) for k in KEYS}
    prtls.fromArrays(first, append=True)
    assert prtls.nactive() == 10 and prtls.nalloc() == 20
    for k, f in zip(KEYS, _cols(prtls)):
        assert np.allclose(first[k], f.as_array())
    second = {k: rng.random(10) for k in KEYS}
    prtls.fromArrays(second, append=True)
    assert prtls.nactive() == 20 and prtls.nalloc() == 20
    for k, f in zip(KEYS, _cols(prtls)):
        assert np.allclose(first[k], f.as_array()[:10])
        assert np.allclose(second[k], f.as_array()[10:])


def test_prtls_coords(rg):  # src/tests/particles.py:193-211
    rng = np.random.default_rng(123)
    prtls = rg.Particles_3D("electrons")
    X1s, U3s = rng.random(10), rng.random(10)
    prtls.fromArrays({"X1": X1s, "U3": U3s}, append=True)
    assert np.allclose(prtls.X(1).as_array(), X1s)
    assert np.allclose(prtls.U(3).as_array(), U3s)
    assert np.allclose(prtls.U(1).as_array(), np.zeros_like(U3s))
    assert np.allclose(prtls.E(1).as_array(), np.zeros_like(U3s))


def test_sync_log(rg):  # src/tests/synchrotron.py:14-35
    p = 2.23
    prtl_dist = rg.TabulatedDistribution(rg.Logbins(1, 1000, 200), rg.PlawGenerator(-p, 1, 1000))
    esync_bins = rg.Logbins(0.01, 1e7, 200)
    esync_bins.unit = rg.EnergyUnits.mec2
    spec = rg.SynchrotronSpectrumFromDist(prtl_dist, esync_bins, 1, 1)
    x_prtls, y_prtls = prtl_dist.EnergyBins().as_array(), prtl_dist.F().as_array()
    x_sync, y_sync = esync_bins.as_array(), spec.as_array()
    assert check_slope(x_prtls, y_prtls, -p, 2, 800)
    assert check_slope(x_sync, y_sync, -p / 2 + 3 / 2, 10, 1e4)
    assert check_slope(x_sync, y_sync, 1 / 3 + 1, 3e-2, 2e-1, atol=0.1)


def test_sync_lin(rg):  # src/tests/synchrotron.py:38-59
    p = 2.5
    dist = rg.TabulatedDistribution(rg.Linbins(1, 1000, 10000), rg.PlawGenerator(-p, 1, 1000))
    bins = rg.Logbins(0.01, 1e6, 500)
    bins.unit = rg.EnergyUnits.mec2
    spec = rg.SynchrotronSpectrumFromDist(dist, bins, 1, 1)
    assert check_slope(dist.EnergyBins().as_array(), dist.F().as_array(), -p, 2, 800)
    assert check_slope(bins.as_array(), spec.as_array(), -p / 2 + 3 / 2, 10, 1e4)
    assert check_slope(bins.as_array(), spec.as_array(), 1 / 3 + 1, 3e-2, 2e-1, atol=0.1)


def test_sync_prtls(rg):  # src/tests/synchrotron.py:62-104
    rng = np.random.default_rng(123)

    def random_plaw(size, xmin, xmax, p, rng):
        return ((xmax ** (p + 1) - xmin ** (p + 1)) * rng.random(size) + xmin ** (p + 1)) ** (
            1 / (p + 1))

    nprtls = int(1e5)
    rnd1 = 2 * (rng.random(nprtls) - 0.5)
    rnd2 = 2 * np.pi * rng.random(nprtls)
    prtls = rg.Particles_3D("pairs")
    prtls.fromArrays({
        "U1": random_plaw(nprtls, 1, 100, -2, rng),
        "B1": np.sqrt(1 - rnd1**2) * np.cos(rnd2),
        "B2": np.sqrt(1 - rnd1**2) * np.sin(rnd2),
        "B3": rnd1,
    })
    dist_prtls = prtls.energyDistribution(rg.Logbins(1, 1e3, 100))
    bins_e_syn = rg.Logbins(0.01, 1e5, 200, rg.EnergyUnits.mec2)
    from_prtls = rg.SynchrotronSpectrum_3D(prtls, bins_e_syn, 1, 1, 1).as_array()
    from_dist = rg.SynchrotronSpectrumFromDist(dist_prtls, bins_e_syn, 1, 1).as_array()
    assert np.abs((from_prtls * 1.3 - from_dist) / from_dist).max() < 0.15
    # unit check of sync::Kernel's constructor (synchrotron.hpp:139-142)
    with pytest.raises(RuntimeError, match="bins_e_syn must be in units of mc\\^2"):
        rg.SynchrotronSpectrum_3D(prtls, rg.Logbins(0.01, 1e5, 200), 1, 1, 1)


def test_module_vs_oracle_end_to_end(rg, port):
    """the pybind path returns exactly what the C-ABI path is tested for"""
    from tests import synth

    U, E, B = synth.full3d(40_000)
    prtls = rg.Particles_3D("e-")
    prtls.fromArrays({f"{q}{d + 1}": a[d] for q, a in (("U", U), ("E", E), ("B", B)) for d in range(3)})
    bins = rg.Logbins(0.01, 1e5, 200, rg.EnergyUnits.mec2)
    spec = rg.SynchrotronSpectrum_3D(prtls, bins, 1.3, 2.0, 0.7).as_array()
    _, want = port.sync_spectrum_particles(U, E, B, bins.as_array(), 1.3, 2.0, 0.7)
    assert synth.rel_err(spec, want) < 1e-5 + 6e-8  # + one float rounding of the result
    gb = rg.Logbins(1e-2, 1e3, 200)
    h = prtls.energyDistribution(gb).F().as_array()
    _, want_h, _ = port.energy_distribution(*U, gb.as_array(), True, True)
    nz = want_h > 0
    assert np.max(np.abs(h[nz] - want_h[nz]) / want_h[nz]) < 1e-5 + 6e-8
    gb.log_spaced = False
    counts = prtls.energyDistribution(gb).F().as_array()
    _, _, want_c = port.energy_distribution(*U, gb.as_array(), False, True)
    assert np.array_equal(counts, want_c.astype(np.float32))


def test_module_edge_cases_follow_the_reference(rg, port):
    """an empty photon Bins gives an empty Array1D (the reference launches N x 0 threads);
    unallocated particles are N = 0: a zero spectrum, no error; the module's spectrum of a
    small population is the reference's float value bit for bit (literal path)"""
    from tests import synth

    U, E, B = synth.full3d(3_000, seed=2)
    prtls = rg.Particles_3D("e-")
    prtls.fromArrays({f"{q}{d + 1}": a[d] for q, a in (("U", U), ("E", E), ("B", B)) for d in range(3)})
    empty = rg.Bins(np.zeros(0, np.float32), rg.EnergyUnits.mec2)
    assert rg.SynchrotronSpectrum_3D(prtls, empty, 1, 1, 1).as_array().shape == (0,)
    bins = rg.Logbins(0.01, 1e5, 50, rg.EnergyUnits.mec2)
    blank = rg.Particles_3D("none")
    assert not blank.is_allocated()
    assert not rg.SynchrotronSpectrum_3D(blank, bins, 1, 1, 1).as_array().any()
    spec = rg.SynchrotronSpectrum_3D(prtls, bins, 1.3, 2.0, 0.7).as_array()
    _, want = port.sync_spectrum_particles(U, E, B, bins.as_array(), 1.3, 2.0, 0.7)
    assert np.array_equal(spec, want.astype(np.float32))
    dist = rg.TabulatedDistribution(rg.Logbins(1, 100, 200), rg.PlawGenerator(-2, 1, 100))
    b2 = rg.Logbins(0.01, 1e7, 200, rg.EnergyUnits.mec2)
    s2 = rg.SynchrotronSpectrumFromDist(dist, b2, 1, 1).as_array()
    _, w2 = port.sync_spectrum_dist(dist.EnergyBins().as_array(), dist.F().as_array(), True, b2.as_array(), 1, 1)
    assert np.array_equal(s2, w2.astype(np.float32))
