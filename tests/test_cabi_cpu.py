"""CPU-side checks of the drop-in boundary: libragnar_cuda.so loads and exports
every symbol include/ragnar_cuda.h declares, the host-exact entry points match
the reference-generated golden vectors bit-for-bit, compute entry points fail
loudly without a GPU (no CPU fallback), and the pybind11 module exposes the
reference's API surface with its docstrings."""
import hashlib
import json
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
G = ROOT / "tests" / "golden"


@pytest.fixture(scope="module")
def cabi_cpu():
    from ragnar_b200 import cabi

    cabi.lib()
    return cabi


def test_library_exports_every_declared_symbol(cabi_cpu):
    header = (ROOT / "include" / "ragnar_cuda.h").read_text()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(rgc_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 45
    lib = cabi_cpu.lib()
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, f"declared in the header but not exported: {missing}"
    assert declared == set(cabi_cpu.SIGNATURES), "ctypes table out of sync with the header"


def test_host_exact_entry_points_match_reference_golden(cabi_cpu):
    g = np.load(G / "spaces.npz")
    for i, (a, b, n) in enumerate(g["cases"]):
        assert np.array_equal(cabi_cpu.logspace(a, b, int(n)), g[f"log_{i}"])
        assert np.array_equal(cabi_cpu.linspace(a, b, int(n)), g[f"lin_{i}"])
    f = np.load(G / "ffunc.npz")
    got = np.array([cabi_cpu.ffunc_integrand(float(x)) for x in f["x"]], np.float32)
    assert np.array_equal(got, f["f"])
    tx, ty = cabi_cpu.tabulate_ffunc()
    assert np.array_equal(tx, f["tab_x"]) and np.array_equal(ty, f["tab_y"])
    gen = np.load(G / "generators.npz")
    for sp in ("log", "lin"):
        e = gen[f"bins_{sp}"]
        assert np.array_equal(cabi_cpu.generator_eval(0, [-1.2, 1e-2, 1], e), gen[f"plaw_{sp}"])
        assert np.array_equal(cabi_cpu.generator_eval(0, [-2.5, 0.5, 0], e), gen[f"plaw_inf_{sp}"])
        assert np.array_equal(cabi_cpu.generator_eval(0, [-1.0, 1e-2, 5], e), gen[f"plaw_m1_{sp}"])
        assert np.array_equal(cabi_cpu.generator_eval(1, [0.3, 0.23, -1.0, 1e-2, 2], e), gen[f"broken_{sp}"])
        assert np.array_equal(cabi_cpu.generator_eval(1, [0.3, 1.5, -2.2, 0, 0], e), gen[f"broken_inf_{sp}"])
        assert np.array_equal(cabi_cpu.generator_eval(2, [2e-2, 0.01], e), gen[f"delta_{sp}"])


def test_host_interpolation_matches_oracle(cabi_cpu):
    import oracle

    tx, ty = cabi_cpu.tabulate_ffunc()
    rng = np.random.default_rng(0)
    for x0 in np.concatenate([10 ** rng.uniform(-7, 2.2, 200), [1e-6, 100.0, tx[0], tx[-1], tx[57]]]):
        assert cabi_cpu.interpolate(float(x0), tx, ty) == oracle.port.interp(float(x0), tx, ty)
    lx = np.linspace(1, 9, 33).astype(np.float32)
    ly = np.sin(lx).astype(np.float32)
    for x0 in rng.uniform(0.5, 9.5, 50):
        assert cabi_cpu.interpolate(float(x0), lx, ly, loggrid=False, yfill=-1.0) == \
            oracle.port.interp(float(x0), lx, ly, loggrid=False, yfill=-1.0)


def test_argument_errors_mirror_reference(cabi_cpu):
    with pytest.raises(cabi_cpu.RagnarCudaError, match="Linspace start must be < stop"):
        cabi_cpu.linspace(2.0, 1.0, 5)
    with pytest.raises(cabi_cpu.RagnarCudaError, match="strictly positive"):
        cabi_cpu.logspace(-1.0, 1.0, 5)
    with pytest.raises(cabi_cpu.RagnarCudaError, match="Logspace start must be < stop"):
        cabi_cpu.logspace(3.0, 1.0, 5)


def test_no_cpu_fallback(cabi_cpu):
    """Without a CUDA device the product path must fail loudly, never compute."""
    if cabi_cpu.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(cabi_cpu.RagnarCudaError) as e:
        cabi_cpu.init()
    assert e.value.code == cabi_cpu.ERR_NOT_INITIALIZED and "no CPU fallback" in str(e.value)
    with pytest.raises(cabi_cpu.RagnarCudaError) as e:
        cabi_cpu.Particles(3).allocate(10)
    assert e.value.code == cabi_cpu.ERR_NOT_INITIALIZED
    bins = cabi_cpu.logspace(0.01, 1e5, 20)
    with pytest.raises(cabi_cpu.RagnarCudaError) as e:
        cabi_cpu.sync_spectrum_dist(bins, bins, True, bins, 1, 1)
    assert e.value.code == cabi_cpu.ERR_NOT_INITIALIZED
    import ragnar_b200

    rg = ragnar_b200.load()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rg.Initialize()
    with pytest.raises(RuntimeError, match="not initialized"):
        rg.Logbins(1, 10, 5)


def test_product_never_imports_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/"""
    for path in (ROOT / "ragnar_b200").rglob("*"):
        if path.suffix in (".py", ".cpp", ".cu", ".hpp", ".h") and "_build" not in path.parts:
            text = path.read_text()
            assert "import oracle" not in text and "ragnar_oracle" not in text, path
            assert "orc_" not in text, path


def test_module_api_surface_and_docstrings():
    """same names as the reference module (+ the HDF5 build's), same docstrings"""
    import ragnar_b200

    rg = ragnar_b200.load()
    api = json.loads((G / "api_surface.json").read_text())
    names = {n for n in dir(rg) if not n.startswith("_")}
    assert set(api) <= names, f"missing: {sorted(set(api) - names)}"
    hdf5_build = {f"H5read1DArray_{t}" for t in "ifd"} | {f"H5write1DArray_{t}" for t in "ifd"} | {
        f"TristanV2_{d}D" for d in (1, 2, 3)}
    assert names - set(api) == hdf5_build

    def norm(doc):
        # pybind11 2.13 (reference) and 3.x (here) print float / int arguments differently
        doc = re.sub(r"typing\.SupportsFloat \| typing\.SupportsIndex|typing\.SupportsFloat", "float", doc or "")
        doc = re.sub(r"typing\.SupportsInt \| typing\.SupportsIndex|typing\.SupportsInt", "int", doc)
        doc = re.sub(r"typing\.Annotated\[numpy\.typing\.ArrayLike, (numpy\.\w+)\]", r"numpy.ndarray[\1]", doc)
        doc = doc.replace("collections.abc.Mapping", "dict").replace("typing.Annotated", "")
        return doc

    def body(doc):
        # the hand-written part (after pybind11's generated signature block)
        parts = norm(doc).split("\n\n", 1)
        return parts[1] if len(parts) > 1 else ""

    ref_docs = json.loads((G / "api_docs.json").read_text())
    for name, info in ref_docs.items():
        obj = getattr(rg, name)
        assert body(obj.__doc__).strip() == body(info["doc"]).strip(), name
        for m, mdoc in info.get("members", {}).items():
            assert hasattr(obj, m), f"{name}.{m} missing"
            assert body(getattr(obj, m).__doc__).strip() == body(mdoc).strip(), f"{name}.{m}"
