"""device-side Linspace / Logspace for grids too large for the host
path, the MinMax reduction of TabulatedFunction and the batched evaluator of
InterpolateTabulatedFunction — all bit-identical to the reference's host arithmetic
(reference src/utils/snippets.cpp:21-62, src/containers/tabulation.{hpp,cpp})."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("num", [1, 2, 200, 65_537, 3_000_001])
@pytest.mark.parametrize("lo,hi", [(1e-6, 100.0), (0.01, 1e7), (3.0, 3.5), (1e-30, 1e30)])
def test_device_spaces_bit_exact(cabi, port, num, lo, hi):
    def same_bits(got, want):
        # (1e-30, 1e30): stop / start overflows float, the reference's own grid is NaN / inf
        # there; NaNs match as NaNs (their payload is not arithmetic), all else bit for bit
        nan = np.isnan(want)
        return np.array_equal(np.isnan(got), nan) and \
            np.array_equal(got[~nan].view(np.uint32), want[~nan].view(np.uint32))

    assert same_bits(cabi.logspace_device(lo, hi, num).to_host(), port.logspace(lo, hi, num))
    assert same_bits(cabi.linspace_device(lo, hi, num).to_host(), port.linspace(lo, hi, num))


def test_device_spaces_errors(cabi):
    with pytest.raises(RuntimeError, match="strictly positive"):
        cabi.logspace_device(0.0, 1.0, 10)
    with pytest.raises(RuntimeError, match="start must be < stop"):
        cabi.logspace_device(2.0, 1.0, 10)
    with pytest.raises(RuntimeError, match="start must be < stop"):
        cabi.linspace_device(2.0, 2.0, 10)


def test_module_builds_large_grids_on_the_device(rg, port):
    """rg.Logspace / rg.Linspace beyond 65536 points never touch a host vector"""
    n = 2_000_003
    assert np.array_equal(rg.Logspace(1e-3, 1e6, n).as_array(), port.logspace(1e-3, 1e6, n))
    assert np.array_equal(rg.Linspace(-5.0, 7.0, n).as_array(), port.linspace(-5.0, 7.0, n))
    b = rg.Logbins(1.0, 1e4, 100_000, rg.EnergyUnits.mec2)
    assert b.log_spaced and np.array_equal(b.as_array(), port.logspace(1.0, 1e4, 100_000))
    # TabulatedFunction on a 1e6-point table: min / max and verify() run on the device
    x = rg.Logspace(1e-4, 1e3, 1_000_000)
    y = rg.Linspace(0.0, 1.0, 1_000_000)
    tf = rg.TabulatedFunction_log(x, y)
    xa = x.as_array()
    assert tf.xMin() == xa.min() and tf.xMax() == xa.max() and tf.nPoints() == 1_000_000
    with pytest.raises(ValueError, match="xmin <= 0.0"):
        rg.TabulatedFunction_log(rg.Linspace(-1.0, 1.0, 100_000), rg.Linspace(0.0, 1.0, 100_000))


def test_minmax_ignores_nans(cabi):
    rng = np.random.default_rng(5)
    a = rng.normal(size=1_234_567).astype(np.float32)
    a[[3, 77, 99_999]] = np.nan
    mn, mx = cabi.DeviceArray.from_host(a).minmax()
    assert mn == np.nanmin(a) and mx == np.nanmax(a)


@pytest.mark.parametrize("loggrid", [True, False])
def test_tabulated_eval_matches_reference_interpolation(cabi, port, loggrid):
    """every point bit-identical to InterpolateTabulatedFunction (the oracle's scalar port),
    on the synchrotron F table and on a 1e6-point table"""
    rng = np.random.default_rng(11)
    tx, ty = port.tabulate_ffunc()
    if not loggrid:
        tx = port.linspace(0.5, 20.0, 200)
    x0 = np.concatenate([
        (10 ** rng.uniform(-7, 2.5, 20_000)).astype(np.float32),
        tx, tx * np.float32(1 + 1e-7), tx * np.float32(1 - 1e-7),
        np.array([0.0, -1.0, np.inf, -np.inf, np.nan, 1e-38, 3e38], np.float32)])
    got = cabi.tabulated_eval(loggrid, cabi.DeviceArray.from_host(tx), cabi.DeviceArray.from_host(ty),
                              cabi.DeviceArray.from_host(x0), yfill=-2.5).to_host()
    want = np.array([port.interp(float(v), tx, ty, loggrid, -2.5) for v in x0], np.float32)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # a large device-resident table
    T = 1_000_000
    bx = cabi.logspace_device(1e-3, 1e3, T) if loggrid else cabi.linspace_device(1e-3, 1e3, T)
    by = cabi.linspace_device(1.0, 2.0, T)
    hx, hy = bx.to_host(), by.to_host()
    pts = (10 ** rng.uniform(-3.2, 3.2, 1_500)).astype(np.float32)
    got = cabi.tabulated_eval(loggrid, bx, by, cabi.DeviceArray.from_host(pts)).to_host()
    want = np.array([port.interp(float(v), hx, hy, loggrid, 0.0) for v in pts], np.float32)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    with pytest.raises(RuntimeError, match="xmin >= xmax"):
        one = cabi.DeviceArray.from_host(np.ones(4, np.float32))
        cabi.tabulated_eval(loggrid, one, one, cabi.DeviceArray.from_host(pts))
