"""Interoperability of the self-contained HDF5 layer (rgc_h5.cpp) with libhdf5, through h5py.
Skipped where h5py is not installed (this image and the GPU boxes have neither h5py nor
libhdf5: the only libhdf5-written file available offline is scipy's MATLAB v7.3 fixture,
tests/test_h5_cpu.py); runs wherever the user-facing files will actually be read."""
import numpy as np
import pytest

h5py = pytest.importorskip("h5py")

from ragnar_b200 import cabi, pipeline  # noqa: E402


def test_files_we_write_open_with_libhdf5(tmp_path):
    """every dataset of write_results / H5File.write / tristan_write_species, read by h5py"""
    pb = np.geomspace(1e-3, 1e3, 200).astype(np.float32)
    gb = np.geomspace(0.1, 200, 200).astype(np.float32)
    rep = pipeline.PipelineReport()
    rng = np.random.default_rng(0)
    for st in (3, 4):
        for label, sp in (("e-", 1), ("e+", 2)):
            rep.results.append(pipeline.SpeciesResult(
                st, label, sp, 100, rng.random(200, dtype=np.float32), rng.random(200, dtype=np.float32),
                rng.random(200), spectrum_from_dist=rng.random(200, dtype=np.float32)))
    out = tmp_path / "spec.h5"
    pipeline.write_results(str(out), rep, pb, gb, multi_step=True)
    with h5py.File(out, "r") as f:
        assert np.array_equal(f["sync_photon_energy_mec2"][:], pb)
        for r in rep.results:
            tag = f"{r.label}_{r.step}"
            assert np.array_equal(f[f"distribution_{tag}"][:], r.distribution)
            assert np.array_equal(f[f"sync_intensity_{tag}"][:], r.spectrum)
            assert np.array_equal(f[f"sync_intensity_dist_{tag}"][:], r.spectrum_from_dist)
    (tmp_path / "output" / "prtl").mkdir(parents=True)
    cols = [rng.random(1000, dtype=np.float32) for _ in range(12)]
    cabi.tristan_write_species(str(tmp_path), 7, 1, cols, with_coords=True, append=False)
    cabi.tristan_write_species(str(tmp_path), 7, 2, cols[3:], with_coords=False, append=True)
    names = ["x", "y", "z", "u", "v", "w", "ex", "ey", "ez", "bx", "by", "bz"]
    with h5py.File(tmp_path / "output" / "prtl" / "prtl.tot.00007", "r") as f:
        for nm, c in zip(names, cols):
            assert np.array_equal(f[f"{nm}_1"][:], c)
    for dt in (np.int32, np.float32, np.float64):
        p = tmp_path / f"arr_{np.dtype(dt).name}.h5"
        a = (rng.random(777) * 100).astype(dt)
        with cabi.H5File(str(p), "w") as f:
            f.create_dataset("a", dt, a.size)
            f.write("a", a)
        with h5py.File(p, "r") as f:
            assert f["a"].dtype == np.dtype(dt) and np.array_equal(f["a"][:], a)


@pytest.mark.parametrize("libver", ["earliest", "latest"])
def test_files_libhdf5_writes_are_read(tmp_path, libver):
    """contiguous / chunked / deflate / shuffle / fletcher32 / f64-stored datasets and both
    superblock generations, written by libhdf5, through the reader and the streaming plugin"""
    rng = np.random.default_rng(1)
    n = 100_003
    a = rng.standard_normal(n).astype(np.float32)
    d = rng.standard_normal(n)
    p = tmp_path / f"lib_{libver}.h5"
    with h5py.File(p, "w", libver=libver) as f:
        f.create_dataset("plain", data=a)
        f.create_dataset("chunked", data=a, chunks=(4096,))
        f.create_dataset("gz", data=a, chunks=(10_000,), compression="gzip", compression_opts=4)
        f.create_dataset("gz_shuf_f32", data=a, chunks=(8192,), compression="gzip", shuffle=True, fletcher32=True)
        f.create_dataset("f64", data=d, chunks=(5000,), compression="gzip", shuffle=True)
        f.create_dataset("i32", data=np.arange(n, dtype=np.int32))
    with cabi.H5File(str(p)) as f:
        for name in ("plain", "chunked", "gz", "gz_shuf_f32"):
            assert np.array_equal(f.read(name), a), name
            assert np.array_equal(f.read(name, 17, 1000, 7), a[17:17 + 7000:7]), name
        assert np.array_equal(f.read("f64", dtype=np.float64), d)
        assert np.array_equal(f.read("f64"), d.astype(np.float32))
        assert np.array_equal(f.read("i32", dtype=np.int32), np.arange(n, dtype=np.int32))
    # appending our datasets to a libhdf5-created file keeps it readable by libhdf5
    if libver == "earliest":
        with cabi.H5File(str(p), "a") as f:
            f.create_dataset("ours", np.float32, 50)
            f.write("ours", a[:50])
        with h5py.File(p, "r") as f:
            assert np.array_equal(f["ours"][:], a[:50]) and np.array_equal(f["gz"][:], a)
