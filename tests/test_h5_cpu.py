"""Host-side HDF5 layer (rgc_h5_*: no GPU needed) — the stand-in for HighFive/libhdf5
at the plugin boundary (reference src/plugins/tristan-v2.cpp:51-74, src/io/h5.cpp).

Pinned against the one libhdf5-written file available offline
(tests/golden/libhdf5_testdouble.mat: MATLAB v7.3 = HDF5 1.6-era superblock v0 behind
a 512-byte user block; it ships with scipy's test data, where scipy documents its
content as the 9 doubles linspace(0, 2*pi, 9)), then by write -> read round trips
and by spec-crafted files for the format features the writer does not emit."""
import shutil
from pathlib import Path

import numpy as np
import pytest

from ragnar_b200 import cabi
from tests import h5craft

GOLDEN = Path(__file__).parent / "golden"


def test_reads_libhdf5_written_fixture():
    with cabi.H5File(GOLDEN / "libhdf5_testdouble.mat") as f:
        assert f.list() == ["testdouble"]
        info = f.info("testdouble")
        assert info == {"dims": [9, 1], "class": 1, "elem_size": 8, "layout": 1}
        got = f.read("testdouble", dtype=np.float64)
        assert np.array_equal(got, np.linspace(0, 2 * np.pi, 9))
        # select({1},{4},{2}) with f64 -> f32 conversion
        assert np.array_equal(f.read("testdouble", 1, 4, 2), got[1:8:2].astype(np.float32))
        with pytest.raises(cabi.RagnarCudaError, match="doesn't exist"):
            f.info("nope")


def test_round_trip_many_datasets(tmp_path):
    """> 256 links needs a two-level group B-tree (8 symbols per node, 32 per tree node)"""
    p = tmp_path / "many.h5"
    rng = np.random.default_rng(0)
    want = {}
    with cabi.H5File(p, "w") as f:
        for k in range(300):
            want[f"d{k:03d}"] = rng.random(10 + k, dtype=np.float32)
            f.create_dataset(f"d{k:03d}", np.float32, 10 + k)
            f.write(f"d{k:03d}", want[f"d{k:03d}"])
        f.create_dataset("ints", np.int32, 5)
        f.write("ints", np.arange(5, dtype=np.int32) - 2)
        f.create_dataset("dbl", np.float64, 70_000)
        f.write("dbl", np.arange(70_000, dtype=np.float64))
        f.create_dataset("empty", np.float32, 0)
        f.create_dataset("zeros", np.float32, 1000)  # never written: sparse zeros
        with pytest.raises(cabi.RagnarCudaError, match="already exists"):
            f.create_dataset("ints", np.int32, 5)
    with cabi.H5File(p) as f:
        assert f.list() == sorted([*want, "ints", "dbl", "empty", "zeros"])
        for name, arr in want.items():
            assert np.array_equal(f.read(name), arr)
        assert np.array_equal(f.read("ints", dtype=np.int32), np.arange(5) - 2)
        assert np.array_equal(f.read("ints"), np.float32([-2, -1, 0, 1, 2]))
        assert np.array_equal(f.read("dbl", 5, 3, 1000), np.float32([5, 1005, 2005]))
        assert f.info("dbl")["dims"] == [70_000] and f.info("dbl")["elem_size"] == 8
        assert f.read("empty").size == 0
        assert not f.read("zeros").any()
        with pytest.raises(cabi.RagnarCudaError, match="exceeds the extent"):
            f.read("ints", 3, 3, 1)
    # slab writes + append mode keep earlier datasets intact
    with cabi.H5File(p, "a") as f:
        f.create_dataset("late", np.float32, 6)
        f.write("late", np.float32([4, 5, 6]), start=3)
        f.write("late", np.float32([1, 2, 3]), start=0)
    with cabi.H5File(p) as f:
        assert len(f.list()) == 305
        assert np.array_equal(f.read("late"), np.float32([1, 2, 3, 4, 5, 6]))
        assert np.array_equal(f.read("d299"), want["d299"])


def test_append_to_libhdf5_file(tmp_path):
    p = tmp_path / "foreign.mat"
    shutil.copy(GOLDEN / "libhdf5_testdouble.mat", p)
    p.chmod(0o644)
    with cabi.H5File(p, "a") as f:
        f.create_dataset("extra", np.float32, 4)
        f.write("extra", np.float32([9, 8, 7, 6]))
    with cabi.H5File(p) as f:
        assert f.list() == ["extra", "testdouble"]
        assert np.array_equal(f.read("extra"), np.float32([9, 8, 7, 6]))
        assert np.array_equal(f.read("testdouble", dtype=np.float64), np.linspace(0, 2 * np.pi, 9))
    assert p.read_bytes()[:19] == b"MATLAB 7.0 MAT-file"  # user block untouched


def test_not_hdf5(tmp_path):
    p = tmp_path / "junk"
    p.write_bytes(b"not an hdf5 file" * 100)
    with pytest.raises(cabi.RagnarCudaError, match="not an HDF5 file"):
        cabi.H5File(p)
    with pytest.raises(cabi.RagnarCudaError, match="Unable to open file"):
        cabi.H5File(tmp_path / "missing.h5")


@pytest.mark.parametrize("deflate,shuffle", [(False, False), (True, False), (True, True)])
def test_spec_crafted_features(tmp_path, deflate, shuffle):
    rng = np.random.default_rng(5)
    a32 = rng.standard_normal(1000).astype(np.float32)
    a64be = rng.standard_normal(77).astype(">f8")
    i64 = (rng.integers(-1000, 1000, 33)).astype(np.int64)
    u16 = rng.integers(0, 60000, 9).astype(np.uint16)
    c = h5craft.Crafter()
    c.chunked("chunky", a32, 96, deflate=deflate, shuffle=shuffle)  # 11 chunks, ragged tail
    c.chunked("holes", a32, 100, deflate=deflate, shuffle=shuffle, skip_chunks=(2, 9))
    c.contiguous("be64", a64be)
    c.compact("small", i64)
    c.contiguous("u16", u16)
    p = tmp_path / "crafted.h5"
    c.save(p)
    with cabi.H5File(p) as f:
        assert f.list() == ["be64", "chunky", "holes", "small", "u16"]
        assert f.info("chunky")["layout"] == 2 and f.info("small")["layout"] == 0
        assert np.array_equal(f.read("chunky"), a32)
        assert np.array_equal(f.read("chunky", 90, 300, 3), a32[90:990:3])
        holes = a32.copy()
        holes[200:300] = 0
        holes[900:1000] = 0
        assert np.array_equal(f.read("holes"), holes)
        assert np.array_equal(f.read("be64", dtype=np.float64), a64be.astype(np.float64))
        assert np.array_equal(f.read("be64"), a64be.astype(np.float32))
        assert np.array_equal(f.read("small", dtype=np.int32), i64.astype(np.int32))
        assert np.array_equal(f.read("u16"), u16.astype(np.float32))


def test_tristan_file_layout(tmp_path):
    """the fixture writer: <path>/output/prtl/prtl.tot.%05d with x_,y_,z_,u_,..,bz_<sp>
    (reference tristan-v2.cpp:110-113,148-185); two species appended into one file"""
    (tmp_path / "output" / "prtl").mkdir(parents=True)
    rng = np.random.default_rng(1)
    cols1 = [rng.random(100, dtype=np.float32) for _ in range(12)]
    cols2 = [rng.random(40, dtype=np.float32) for _ in range(9)]
    cabi.tristan_write_species(str(tmp_path), 7, 1, cols1, with_coords=True, append=False)
    cabi.tristan_write_species(str(tmp_path), 7, 2, cols2, with_coords=False, append=True)
    names = ["x", "y", "z", "u", "v", "w", "ex", "ey", "ez", "bx", "by", "bz"]
    with cabi.H5File(tmp_path / "output" / "prtl" / "prtl.tot.00007") as f:
        assert f.list() == sorted(f"{n}_{s}" for n in names for s in (1, 2))
        for n, c in zip(names, cols1):
            assert np.array_equal(f.read(f"{n}_1"), c)
        for n, c in zip(names[3:], cols2):
            assert np.array_equal(f.read(f"{n}_2"), c)
        assert f.info("x_2")["dims"] == [40] and not f.read("x_2").any()


def test_pipeline_write_results_legacy_dataset_names(tmp_path):
    """ragnar_b200/pipeline.py writes what the retired driver wrote
    (legacy/simulation.cpp.bak:39,60-64,156-159); host-only, no GPU needed"""
    from ragnar_b200 import cabi, pipeline

    pb = np.geomspace(1e-3, 1e3, 20).astype(np.float32)
    gb = np.geomspace(0.1, 200, 10).astype(np.float32)
    rep = pipeline.PipelineReport()
    for st in (3, 4):
        for label, sp in (("e-", 1), ("e+", 2)):
            rep.results.append(pipeline.SpeciesResult(st, label, sp, 100, (gb * st).astype(np.float32),
                                                      (pb * sp).astype(np.float32), (pb * sp).astype(np.float64)))
    out = str(tmp_path / "spec.h5")
    pipeline.write_results(out, rep, pb, gb, multi_step=True)
    with cabi.H5File(out, "r") as f:
        names = set(f.list("/"))
        assert names == {"sync_photon_energy_mec2"} | {f"{k}_{lab}_{st}" for k in ("gammaM1", "distribution", "sync_intensity")
                                                        for lab in ("e-", "e+") for st in (3, 4)}
        assert np.array_equal(f.read("sync_intensity_e+_4"), pb * 2)
        assert np.array_equal(f.read("distribution_e-_3"), gb * 3)
    pipeline.write_results(out, rep.__class__(results=rep.results[:2]), pb, gb, multi_step=False)
    with cabi.H5File(out, "r") as f:
        assert {"sync_intensity_e-", "gammaM1_e+", "distribution_e+"} <= set(f.list("/"))
