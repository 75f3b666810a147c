"""TEST INFRASTRUCTURE — CPU oracle for ragnar's radiation hot path.

Two checkers live here (neither is ever on the product path; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs import this):

* ``oracle.port``  — ctypes bindings of ``ragnar_oracle.cpp``, a C++ restatement
  of the reference arithmetic (each function cites the reference file:line).
* ``oracle.ref()`` / ``oracle.ref64()`` — the reference's *own unmodified
  sources* compiled against a Kokkos-subset shim by ``build_ref.sh`` into
  ``oracle/_ref/`` (float ScatterView / double ScatterView).  They exist only
  where ``/root/reference`` was present at build time (they travel to the GPU
  box as prebuilt ``.so`` files).
"""
from __future__ import annotations

import ctypes
import importlib
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
BUILD_DIR = HERE / "_build"
REF_DIR = HERE / "_ref"
PORT_SRC = HERE / "ragnar_oracle.cpp"
PORT_LIB = BUILD_DIR / "libragnar_oracle.so"


def build_port(force: bool = False) -> Path:
    """Compile the C++ restatement (g++, no FMA contraction, OpenMP)."""
    if (
        not force
        and PORT_LIB.exists()
        and PORT_LIB.stat().st_mtime >= PORT_SRC.stat().st_mtime
    ):
        return PORT_LIB
    BUILD_DIR.mkdir(exist_ok=True)
    cmd = [
        "g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-fopenmp",
        "-ffp-contract=off", str(PORT_SRC), "-o", str(PORT_LIB),
    ]
    subprocess.run(cmd, check=True)
    return PORT_LIB


def build_ref() -> bool:
    """Build oracle/_ref from /root/reference when it is present. Returns availability."""
    subprocess.run(["bash", str(HERE / "build_ref.sh")], check=True)
    return ref_available()


def ref_available() -> bool:
    return any(REF_DIR.glob("ragnar_ref.*.so")) and any(REF_DIR.glob("ragnar_ref64.*.so"))


_ref_modules: dict = {}


def _import_ref(name: str):
    """Imports and initialises the module once; the reference's Initialize() banner
    ("Kokkos is already initialized" on a second call) is kept off stdout."""
    if name in _ref_modules:
        return _ref_modules[name]
    import contextlib
    import io

    if str(REF_DIR) not in sys.path:
        sys.path.insert(0, str(REF_DIR))
    mod = importlib.import_module(name)
    with contextlib.redirect_stdout(io.StringIO()):
        mod.Initialize()
    _ref_modules[name] = mod
    return mod


def ref():
    """The reference's own sources, float ScatterView (faithful)."""
    return _import_ref("ragnar_ref")


def ref64():
    """The reference's own sources, per-pair float terms summed in double."""
    return _import_ref("ragnar_ref64")


_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_u64p = ctypes.POINTER(ctypes.c_uint64)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a: np.ndarray, t=_f32p):
    return a.ctypes.data_as(t)


class _Port:
    """ctypes facade over libragnar_oracle.so (lazy build + load)."""

    def __init__(self) -> None:
        self._lib = None

    @property
    def lib(self):
        if self._lib is None:
            lib = ctypes.CDLL(str(build_port()))
            lib.orc_ffunc_integrand.restype = ctypes.c_float
            lib.orc_ffunc_integrand.argtypes = [ctypes.c_float]
            lib.orc_interp.restype = ctypes.c_float
            lib.orc_interp.argtypes = [ctypes.c_int, ctypes.c_float, _f32p, _f32p,
                                       ctypes.c_size_t, ctypes.c_float]
            lib.orc_plaw_norm.restype = ctypes.c_float
            lib.orc_plaw_norm.argtypes = [ctypes.c_float] * 3
            lib.orc_num_threads.restype = ctypes.c_int
            lib.orc_set_num_threads.argtypes = [ctypes.c_int]
            lib.orc_set_num_threads.restype = None
            for name in ("orc_logspace", "orc_linspace"):
                getattr(lib, name).argtypes = [ctypes.c_float, ctypes.c_float,
                                               ctypes.c_size_t, _f32p]
                getattr(lib, name).restype = None
            lib.orc_tabulate_ffunc.argtypes = [ctypes.c_size_t, ctypes.c_float,
                                               ctypes.c_float, _f32p, _f32p]
            lib.orc_plaw_f.argtypes = [ctypes.c_float] * 3 + [_f32p, ctypes.c_size_t, _f32p]
            lib.orc_broken_plaw_f.argtypes = [ctypes.c_float] * 5 + [_f32p, ctypes.c_size_t, _f32p]
            lib.orc_delta_f.argtypes = [ctypes.c_float] * 2 + [_f32p, ctypes.c_size_t, _f32p]
            lib.orc_energy_distribution.argtypes = [
                _f32p, _f32p, _f32p, ctypes.c_size_t, _f32p, ctypes.c_size_t,
                ctypes.c_int, ctypes.c_int, _f32p, _f64p, _u64p]
            lib.orc_sync_spectrum_particles.argtypes = (
                [_f32p] * 9 + [ctypes.c_size_t, _f32p, ctypes.c_size_t, _f32p, _f32p,
                               ctypes.c_size_t, ctypes.c_float, ctypes.c_float,
                               ctypes.c_float, _f32p, _f64p])
            lib.orc_sync_epeak_chir.argtypes = (
                [_f32p] * 9 + [ctypes.c_size_t, ctypes.c_float, ctypes.c_float,
                               ctypes.c_float, _f32p, _f32p])
            lib.orc_sync_spectrum_dist.argtypes = [
                _f32p, _f32p, ctypes.c_size_t, ctypes.c_int, _f32p, ctypes.c_size_t,
                _f32p, _f32p, ctypes.c_size_t, ctypes.c_float, ctypes.c_float,
                _f32p, _f64p]
            lib.orc_ic_spectrum.argtypes = [
                _f32p, _f32p, ctypes.c_size_t, ctypes.c_int, _f32p, _f32p, ctypes.c_size_t,
                _f32p, ctypes.c_size_t, _f32p, _f64p]
            self._lib = lib
        return self._lib

    # -- spaces ------------------------------------------------------------
    def logspace(self, start, stop, num) -> np.ndarray:
        out = np.empty(num, np.float32)
        self.lib.orc_logspace(start, stop, num, _p(out))
        return out

    def linspace(self, start, stop, num) -> np.ndarray:
        out = np.empty(num, np.float32)
        self.lib.orc_linspace(start, stop, num, _p(out))
        return out

    # -- synchrotron function table ------------------------------------------
    def ffunc_integrand(self, x: float) -> float:
        return float(self.lib.orc_ffunc_integrand(x))

    _tab_cache: dict = {}

    def tabulate_ffunc(self, n=200, xmin=1e-6, xmax=100.0):
        key = (n, float(np.float32(xmin)), float(np.float32(xmax)))
        if key not in self._tab_cache:
            xs = np.empty(n, np.float32)
            ys = np.empty(n, np.float32)
            self.lib.orc_tabulate_ffunc(n, xmin, xmax, _p(xs), _p(ys))
            self._tab_cache[key] = (xs, ys)
        xs, ys = self._tab_cache[key]
        return xs.copy(), ys.copy()

    def interp(self, x0, x, y, loggrid=True, yfill=0.0) -> float:
        x, y = _f32(x), _f32(y)
        return float(self.lib.orc_interp(int(loggrid), x0, _p(x), _p(y), len(x), yfill))

    # -- generators ----------------------------------------------------------
    def plaw_f(self, p, emin, emax, energy) -> np.ndarray:
        e = _f32(energy)
        out = np.empty_like(e)
        self.lib.orc_plaw_f(p, emin, emax, _p(e), len(e), _p(out))
        return out

    def broken_plaw_f(self, e_break, p1, p2, emin, emax, energy) -> np.ndarray:
        e = _f32(energy)
        out = np.empty_like(e)
        self.lib.orc_broken_plaw_f(e_break, p1, p2, emin, emax, _p(e), len(e), _p(out))
        return out

    def delta_f(self, energy0, denergy, energy) -> np.ndarray:
        e = _f32(energy)
        out = np.empty_like(e)
        self.lib.orc_delta_f(energy0, denergy, _p(e), len(e), _p(out))
        return out

    # -- hot path --------------------------------------------------------------
    def energy_distribution(self, u1, u2, u3, bins, log_spaced, fourvel=True):
        """-> (hist_f32 faithful serial, hist_f64, counts u64)"""
        u1, u2, u3, bins = _f32(u1), _f32(u2), _f32(u3), _f32(bins)
        n = len(bins)
        h32 = np.zeros(n, np.float32)
        h64 = np.zeros(n, np.float64)
        cnt = np.zeros(n, np.uint64)
        self.lib.orc_energy_distribution(_p(u1), _p(u2), _p(u3), len(u1), _p(bins), n,
                                         int(log_spaced), int(fourvel), _p(h32),
                                         _p(h64, _f64p), _p(cnt, _u64p))
        return h32, h64, cnt

    def sync_spectrum_particles(self, U, E, B, bins_e_syn, B0, g_syn, e_at,
                                table=None, want_f32=False):
        """U, E, B: sequences of three float arrays. -> (spec_f32 | None, spec_f64)"""
        comps = [_f32(c) for q in (U, E, B) for c in q]
        n = len(comps[0])
        bins = _f32(bins_e_syn)
        tx, ty = table if table is not None else self.tabulate_ffunc()
        tx, ty = _f32(tx), _f32(ty)
        s64 = np.zeros(len(bins), np.float64)
        s32 = np.zeros(len(bins), np.float32) if want_f32 else None
        self.lib.orc_sync_spectrum_particles(
            *[_p(c) for c in comps], n, _p(bins), len(bins), _p(tx), _p(ty), len(tx),
            B0, g_syn, e_at, _p(s32) if want_f32 else None, _p(s64, _f64p))
        return s32, s64

    def sync_epeak_chir(self, U, E, B, B0, g_syn, e_at):
        comps = [_f32(c) for q in (U, E, B) for c in q]
        n = len(comps[0])
        ep = np.empty(n, np.float32)
        ch = np.empty(n, np.float32)
        self.lib.orc_sync_epeak_chir(*[_p(c) for c in comps], n, B0, g_syn, e_at,
                                     _p(ep), _p(ch))
        return ep, ch

    def sync_spectrum_dist(self, gbeta, f, islog, bins_e_syn, g_syn, e_at, table=None):
        """-> (spec_f32 faithful serial order, spec_f64)"""
        gbeta, f, bins = _f32(gbeta), _f32(f), _f32(bins_e_syn)
        tx, ty = table if table is not None else self.tabulate_ffunc()
        tx, ty = _f32(tx), _f32(ty)
        s32 = np.zeros(len(bins), np.float32)
        s64 = np.zeros(len(bins), np.float64)
        self.lib.orc_sync_spectrum_dist(_p(gbeta), _p(f), len(gbeta), int(islog), _p(bins),
                                        len(bins), _p(tx), _p(ty), len(tx), g_syn, e_at,
                                        _p(s32), _p(s64, _f64p))
        return s32, s64

    def ic_spectrum(self, g_prtls, f_prtls, islog, e_soft, f_soft, bins_e_ic):
        """-> (spec_f32 faithful serial order, spec_f64)"""
        g, f, es, fs, b = (_f32(a) for a in (g_prtls, f_prtls, e_soft, f_soft, bins_e_ic))
        s32 = np.zeros(len(b), np.float32)
        s64 = np.zeros(len(b), np.float64)
        self.lib.orc_ic_spectrum(_p(g), _p(f), len(g), int(islog), _p(es), _p(fs), len(es),
                                 _p(b), len(b), _p(s32), _p(s64, _f64p))
        return s32, s64

    def num_threads(self) -> int:
        return int(self.lib.orc_num_threads())

    def set_num_threads(self, n: int) -> int:
        """OpenMP width of the oracle AND of oracle/_ref (one libgomp per process)."""
        self.lib.orc_set_num_threads(int(n))
        return self.num_threads()


port = _Port()
