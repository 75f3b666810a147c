"""TristanV2 plugin and H5read/H5write1DArray on the GPU.

No reference-side oracle exists for this boundary (no libhdf5 here, and the
reference's host read path is defective — SURVEY a15), so parity is ARRAY IDENTITY:
what the synthetic file holds must come back bit-identical through readParticles;
spectrum / histogram parity then reduces to the oracle on the same arrays."""
import numpy as np
import pytest

from tests import synth

pytestmark = pytest.mark.gpu

NAMES = ["x", "y", "z", "u", "v", "w", "ex", "ey", "ez", "bx", "by", "bz"]


def _write_step(cabi, root, step, n1, n2, seed=3):
    (root / "output" / "prtl").mkdir(parents=True, exist_ok=True)
    U, E, B = synth.full3d(n1, seed=seed)
    rng = np.random.default_rng(seed)
    X = [rng.random(n1, dtype=np.float32) * 100 for _ in range(3)]
    sp1 = [*X, *U, *E, *B]
    U2, E2, B2 = synth.full3d(n2, seed=seed + 1)
    sp2 = [*U2, *E2, *B2]
    cabi.tristan_write_species(str(root), step, 1, sp1, with_coords=True, append=False)
    cabi.tristan_write_species(str(root), step, 2, sp2, with_coords=False, append=True)
    return sp1, sp2


def _columns(p, D, with_coords):
    out = []
    if with_coords:
        out += [p.X(d + 1).as_array() for d in range(D)]
    for get in (p.U, p.E, p.B):
        out += [get(d + 1).as_array() for d in range(3)]
    return out


@pytest.mark.parametrize("D", [1, 2, 3])
def test_read_particles_array_identity(cabi, rg, tmp_path, D, capfd):
    n1, n2 = 100_003, 5_000_017  # the second spans several 16 MiB slabs per column
    sp1, sp2 = _write_step(cabi, tmp_path, 12, n1, n2)
    plug = getattr(rg, f"TristanV2_{D}D")()
    assert plug.label() == "Tristan V2"
    with pytest.raises(RuntimeError, match="Path not set"):
        plug.getPath()
    with pytest.raises(RuntimeError, match="Step not set"):
        plug.getStep()
    plug.setPath(str(tmp_path))
    plug.setStep(12)
    p = plug.readParticles("e-", 1)
    out = capfd.readouterr().out
    assert f"Reading particles # 1 from {tmp_path}/output/prtl/prtl.tot.00012 ..." in out
    assert "found 1.00·10^5 particles, reading 1.00·10^5 starting from 0" in out
    assert out.count(": OK") == D + 9
    assert p.nactive() == n1 and p.label() == "e-"
    got = _columns(p, D, True)
    want = sp1[:D] + sp1[3:]
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    q = plug.readParticles("e+", 2, ignore_coordinates=True)
    assert q.nactive() == n2
    for g, w in zip(_columns(q, D, False), sp2):
        assert np.array_equal(g, w)


def test_read_particles_selections_and_errors(cabi, rg, tmp_path):
    n1 = 10_000
    sp1, _ = _write_step(cabi, tmp_path, 3, n1, 16)
    plug = rg.TristanV2_3D()
    plug.setPath(str(tmp_path))
    plug.setStep(3)
    # explicit window (reference: nparticles = size, hyperslab {start},{size},{1})
    p = plug.readParticles("e-", 1, start=17, size=4000)
    assert p.nactive() == 4000
    assert np.array_equal(p.U(1).as_array(), sp1[3][17:4017])
    assert np.array_equal(p.B(3).as_array(), sp1[11][17:4017])
    # stride: nparticles = ntotal / stride (tristan-v2.cpp:129)
    p = plug.readParticles("e-", 1, stride=7)
    assert p.nactive() == n1 // 7
    assert np.array_equal(p.E(2).as_array(), sp1[7][::7][: n1 // 7])
    assert np.array_equal(p.X(2).as_array(), sp1[1][::7][: n1 // 7])
    p = plug.readParticles("e-", 1, start=1, stride=2)
    assert np.array_equal(p.U(3).as_array(), sp1[5][1::2][: n1 // 2])
    # validation order and messages of tristan-v2.cpp:102-107,126-128
    with pytest.raises(RuntimeError, match="Stride must be greater than 0"):
        plug.readParticles("e-", 1, stride=0)
    with pytest.raises(RuntimeError, match=r"Size must be determined automatically \(0\) when stride != 1"):
        plug.readParticles("e-", 1, size=10, stride=2)
    with pytest.raises(RuntimeError, match="start \\+ size >= total number of particles"):
        plug.readParticles("e-", 1, start=0, size=n1)  # an explicit full-length size is rejected
    with pytest.raises(RuntimeError, match="start \\+ size >= total number of particles"):
        plug.readParticles("e-", 1, start=n1)
    with pytest.raises(RuntimeError, match="hyperslab"):
        plug.readParticles("e-", 1, start=5)  # size := ntotal overruns the extent (HighFive throws)
    with pytest.raises(RuntimeError, match="doesn't exist"):
        plug.readParticles("ions", 9)
    plug.setStep(4)
    with pytest.raises(RuntimeError, match="Unable to open file"):
        plug.readParticles("e-", 1)


def test_read_particles_converts_f64_and_mismatch(cabi, rg, tmp_path):
    """double-precision / integer datasets convert on read (HighFive read<float>)"""
    d = tmp_path / "output" / "prtl"
    d.mkdir(parents=True)
    n = 3_000_001
    rng = np.random.default_rng(0)
    cols = {f"{nm}_1": rng.standard_normal(n) for nm in NAMES}
    with cabi.H5File(d / "prtl.tot.00001", "w") as f:
        for k, (name, a) in enumerate(cols.items()):
            if k == 0:
                f.create_dataset(name, np.int32, n)
                cols[name] = rng.integers(-50, 50, n).astype(np.int32)
                f.write(name, cols[name])
            else:
                f.create_dataset(name, np.float64, n)
                f.write(name, a)
        f.create_dataset("x_2", np.float32, 10)
        for nm in NAMES[1:]:
            f.create_dataset(f"{nm}_2", np.float32, 9 if nm == "ey" else 10)
    plug = rg.TristanV2_3D()
    plug.setPath(str(tmp_path))
    plug.setStep(1)
    p = plug.readParticles("e-", 1)
    assert np.array_equal(p.X(1).as_array(), cols["x_1"].astype(np.float32))
    assert np.array_equal(p.U(2).as_array(), cols["v_1"].astype(np.float32))
    assert np.array_equal(p.B(1).as_array(), cols["bx_1"].astype(np.float32))
    q = plug.readParticles("e-", 1, stride=3)
    assert np.array_equal(q.E(3).as_array(), cols["ez_1"][::3][: n // 3].astype(np.float32))
    with pytest.raises(RuntimeError, match="Number of particles mismatch"):
        plug.readParticles("e+", 2)


def test_plugin_feeds_the_hot_path(cabi, rg, port, tmp_path):
    """config 4 in miniature: read -> energyDistribution + SynchrotronSpectrum_3D equals
    the oracle on the arrays that were written"""
    n = 400_000
    sp1, sp2 = _write_step(cabi, tmp_path, 1, n, n // 2, seed=9)
    plug = rg.TristanV2_3D()
    plug.setPath(str(tmp_path))
    plug.setStep(1)
    bins = rg.Logbins(0.01, 1e5, 200, rg.EnergyUnits.mec2)
    gb = rg.Logbins(1e-2, 1e3, 200)
    gb.log_spaced = False
    for label, sp, cols, off in (("e-", 1, sp1, 3), ("e+", 2, sp2, 0)):
        p = plug.readParticles(label, sp, ignore_coordinates=(sp == 2))
        U, E, B = cols[off:off + 3], cols[off + 3:off + 6], cols[off + 6:off + 9]
        spec = rg.SynchrotronSpectrum_3D(p, bins, 1.0, 1.0, 1.0).as_array()
        _, want = port.sync_spectrum_particles(U, E, B, bins.as_array(), 1.0, 1.0, 1.0)
        assert synth.rel_err(spec, want) < 1e-5
        counts = p.energyDistribution(gb).F().as_array()
        _, _, want_c = port.energy_distribution(*U, gb.as_array(), False, True)
        assert np.array_equal(counts, want_c.astype(np.float32))


def test_h5_array_io_through_device(cabi, rg, tmp_path, capfd):
    """H5write1DArray_* / H5read1DArray_* (reference src/io/h5.cpp:16-68)"""
    p = str(tmp_path / "arrays.h5")
    f32 = np.random.default_rng(2).random(5_000_000, dtype=np.float32)
    rg.H5write1DArray_f(p, "spec", rg.Array1D_f(f32))
    assert "Writing spec to" in capfd.readouterr().out
    rg.H5write1DArray_d(p, "dbl", rg.Array1D_d(np.arange(10, dtype=np.float64) / 3))
    rg.H5write1DArray_i(p, "ints", rg.Array1D_i(np.arange(-3, 4, dtype=np.int32)))
    rg.H5write1DArray_f(p, "bins", rg.Logbins(1, 100, 7))  # Bins is an Array1D_f
    with pytest.raises(RuntimeError, match="already exists"):
        rg.H5write1DArray_f(p, "spec", rg.Array1D_f(f32))
    back = rg.H5read1DArray_f(p, "spec")
    assert "Reading spec from" in capfd.readouterr().out
    assert np.array_equal(back.as_array(), f32)
    assert np.array_equal(rg.H5read1DArray_d(p, "dbl").as_array(), np.arange(10) / 3)
    assert np.array_equal(rg.H5read1DArray_i(p, "ints").as_array(), np.arange(-3, 4))
    assert np.array_equal(rg.H5read1DArray_f(p, "bins").as_array(), rg.Logbins(1, 100, 7).as_array())
    # size: 0 -> extent; dims[0]/stride > size -> error; stride > 1 needs a fitting size
    assert np.array_equal(rg.H5read1DArray_f(p, "ints", size=4, stride=2).as_array(),
                          np.float32([-3, -1, 1, 3]))
    with pytest.raises(RuntimeError, match="Number of read quantity exceeds allocated space"):
        rg.H5read1DArray_f(p, "ints", size=2)
    with pytest.raises(RuntimeError, match="Stride must be greater than 0"):
        rg.H5read1DArray_f(p, "ints", stride=0)
    with pytest.raises(RuntimeError, match="hyperslab"):
        rg.H5read1DArray_f(p, "ints", stride=2)  # size := 7, 7 x stride 2 leaves the extent
    with pytest.raises(RuntimeError, match="doesn't exist"):
        rg.H5read1DArray_f(p, "nope")
    with cabi.H5File(p) as f:
        assert f.list() == ["bins", "dbl", "ints", "spec"]
