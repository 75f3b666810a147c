"""ctypes binding of libragnar_cuda.so — one Python function per entry point of
include/ragnar_cuda.h, plus small numpy conveniences.  This is what a harness in
another language would bind (INTEGRATION.md shows the C++ side)."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

LIB_PATH = Path(__file__).resolve().parent / "libragnar_cuda.so"

OK, ERR_INVALID, ERR_NOT_INITIALIZED, ERR_CUDA, ERR_NCCL, ERR_IO, ERR_OOM = range(7)
I32, F32, F64 = 0, 1, 2
Q_X, Q_U, Q_E, Q_B = 0, 1, 2, 3
COMM_ID_BYTES = 128

_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_u64p = C.POINTER(C.c_uint64)
_vp = C.c_void_p
_vpp = C.POINTER(C.c_void_p)
_sz = C.c_size_t

# name -> (restype, argtypes): must list every symbol the header declares
SIGNATURES = {
    "rgc_init": (C.c_int, [C.c_int]),
    "rgc_finalize": (C.c_int, []),
    "rgc_is_initialized": (C.c_int, []),
    "rgc_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "rgc_device_info": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(_sz)]),
    "rgc_stream": (C.c_int, [_vpp]),
    "rgc_synchronize": (C.c_int, []),
    "rgc_last_error": (C.c_char_p, []),
    "rgc_launch_count": (C.c_uint64, []),
    "rgc_host_alloc": (C.c_int, [_sz, _vpp]),
    "rgc_host_free": (C.c_int, [_vp]),
    "rgc_comm_get_unique_id": (C.c_int, [C.c_char_p]),
    "rgc_comm_init": (C.c_int, [C.c_char_p, C.c_int, C.c_int]),
    "rgc_comm_destroy": (C.c_int, []),
    "rgc_comm_info": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "rgc_buf_create": (C.c_int, [C.c_int, _sz, _vpp]),
    "rgc_buf_from_host": (C.c_int, [C.c_int, _vp, _sz, _vpp]),
    "rgc_buf_to_host": (C.c_int, [_vp, _sz, _sz, _vp]),
    "rgc_buf_size": (_sz, [_vp]),
    "rgc_buf_dtype": (C.c_int, [_vp]),
    "rgc_buf_device_ptr": (_vp, [_vp]),
    "rgc_buf_retain": (C.c_int, [_vp]),
    "rgc_buf_release": (C.c_int, [_vp]),
    "rgc_particles_create": (C.c_int, [C.c_int, _vpp]),
    "rgc_particles_release": (C.c_int, [_vp]),
    "rgc_particles_allocate": (C.c_int, [_vp, _sz, C.c_int]),
    "rgc_particles_reallocate": (C.c_int, [_vp, _sz]),
    "rgc_particles_enable_coords": (C.c_int, [_vp]),
    "rgc_particles_has_coords": (C.c_int, [_vp]),
    "rgc_particles_nalloc": (_sz, [_vp]),
    "rgc_particles_dim": (C.c_int, [_vp]),
    "rgc_particles_write": (C.c_int, [_vp, C.c_int, C.c_int, _sz, _vp, _sz]),
    "rgc_particles_read": (C.c_int, [_vp, C.c_int, C.c_int, _sz, _sz, _vp]),
    "rgc_particles_column": (C.c_int, [_vp, C.c_int, C.c_int, _sz, _vpp]),
    "rgc_particles_device_ptr": (_vp, [_vp, C.c_int, C.c_int]),
    "rgc_particles_generate": (C.c_int, [_vp, C.c_int, C.c_uint64, C.c_uint64, _sz, _sz,
                                         C.c_float, C.c_float]),
    "rgc_linspace": (C.c_int, [C.c_float, C.c_float, _sz, _f32p]),
    "rgc_logspace": (C.c_int, [C.c_float, C.c_float, _sz, _f32p]),
    "rgc_linspace_device": (C.c_int, [C.c_float, C.c_float, _sz, _vpp]),
    "rgc_logspace_device": (C.c_int, [C.c_float, C.c_float, _sz, _vpp]),
    "rgc_buf_minmax": (C.c_int, [_vp, _f32p, _f32p]),
    "rgc_tabulated_eval": (C.c_int, [C.c_int, _vp, _vp, C.c_float, _vp, _vpp]),
    "rgc_sync_ffunc_integrand": (C.c_int, [C.c_float, _f32p]),
    "rgc_sync_tabulate_ffunc": (C.c_int, [_sz, C.c_float, C.c_float, _f32p, _f32p]),
    "rgc_interpolate": (C.c_int, [C.c_int, C.c_float, _f32p, _f32p, _sz, C.c_float, _f32p]),
    "rgc_generator_eval": (C.c_int, [C.c_int, _f32p, _f32p, _sz, _f32p]),
    "rgc_energy_histogram": (C.c_int, [_vp, _sz, _f32p, _sz, C.c_int, C.c_int, _f32p, _u64p,
                                       _f64p]),
    "rgc_sync_spectrum_particles": (C.c_int, [_vp, _sz, _f32p, _sz, _f32p, _f32p, _sz,
                                              C.c_float, C.c_float, C.c_float, _f32p, _f64p]),
    "rgc_hist_and_spectrum": (C.c_int, [_vp, _sz, _f32p, _sz, C.c_int, C.c_int, _f32p, _f64p, _f32p, _sz,
                                        _f32p, _f32p, _sz, C.c_float, C.c_float, C.c_float, _f32p, _f64p]),
    "rgc_sync_spectrum_dist": (C.c_int, [_f32p, _f32p, _sz, C.c_int, _f32p, _sz, _f32p, _f32p,
                                         _sz, C.c_float, C.c_float, _f32p, _f64p]),
    "rgc_sync_spectrum_dist_batch": (C.c_int, [_f32p, _f32p, _sz, _sz, C.c_int, _f32p, _sz, _f32p, _f32p,
                                               _sz, C.c_float, C.c_float, C.c_int, _f32p, _f64p]),
    "rgc_ic_spectrum": (C.c_int, [_f32p, _f32p, _sz, C.c_int, _f32p, _f32p, _sz, _f32p, _sz,
                                  _f32p, _f64p]),
    "rgc_last_kernel_ms": (C.c_int, [_f32p]),
    "rgc_last_kernel_times": (C.c_int, [_f32p, C.c_int]),
    "rgc_trim_memory": (C.c_int, []),
    "rgc_last_pair_lane_evals": (C.c_int, [_f64p]),
    "rgc_last_pair_ontable_evals": (C.c_int, [_f64p]),
    "rgc_sort_rank_mode": (C.c_int, [C.POINTER(C.c_int)]),
    "rgc_pair_plan_describe": (C.c_int, [_f32p, C.c_size_t, _f32p, _f32p, C.c_size_t,
                                         C.POINTER(C.c_int), C.POINTER(C.c_float),
                                         C.POINTER(C.c_int), C.c_size_t]),
    "rgc_comm_exchange_kind": (C.c_int, [C.POINTER(C.c_int)]),
    "rgc_measure_peak": (C.c_int, [C.c_int, _f64p, _f64p]),
    "rgc_h5_open": (C.c_int, [C.c_char_p, C.c_int, _vpp]),
    "rgc_h5_close": (C.c_int, [_vp]),
    "rgc_h5_list": (C.c_int, [_vp, C.c_char_p, C.c_char_p, _sz, C.POINTER(_sz)]),
    "rgc_h5_dataset_info": (C.c_int, [_vp, C.c_char_p, C.POINTER(C.c_int), _u64p, C.c_int,
                                      C.POINTER(C.c_int), C.POINTER(C.c_int),
                                      C.POINTER(C.c_int)]),
    "rgc_h5_read": (C.c_int, [_vp, C.c_char_p, _sz, _sz, _sz, C.c_int, _vp]),
    "rgc_h5_create_dataset": (C.c_int, [_vp, C.c_char_p, C.c_int, _sz]),
    "rgc_h5_write": (C.c_int, [_vp, C.c_char_p, _sz, _sz, C.c_int, _vp]),
    "rgc_h5_read_array": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int, _sz, _sz, _vpp]),
    "rgc_h5_write_array": (C.c_int, [C.c_char_p, C.c_char_p, _vp]),
    "rgc_tristan_read_particles": (C.c_int, [C.c_char_p, _sz, C.c_uint, _sz, _sz, _sz, C.c_int,
                                             C.c_int, _vpp, C.POINTER(_sz), C.POINTER(_sz)]),
    "rgc_tristan_read_range": (C.c_int, [C.c_char_p, _sz, C.c_uint, _sz, _sz, C.c_int, C.c_int,
                                         _vpp, C.POINTER(_sz)]),
    "rgc_tristan_write_species": (C.c_int, [C.c_char_p, _sz, C.c_uint, _sz, C.c_int,
                                            C.POINTER(_f32p), C.c_int]),
}

_lib = None


class RagnarCudaError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(message or f"ragnar_cuda error {code}")
        self.code = code


def lib() -> C.CDLL:
    """Load libragnar_cuda.so (fails loudly when it has not been built)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(f"{LIB_PATH} not found — run `python -m ragnar_b200.build`; "
                              "there is no CPU fallback")
        handle = C.CDLL(str(LIB_PATH), mode=C.RTLD_GLOBAL)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != OK:
        raise RagnarCudaError(rc, lib().rgc_last_error().decode("utf-8", "replace"))


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a: np.ndarray, t=_f32p):
    return a.ctypes.data_as(t)


# ---------------------------------------------------------------- runtime
def init(device: int = -1) -> None:
    check(lib().rgc_init(device))


def finalize() -> None:
    check(lib().rgc_finalize())


def device_count() -> int:
    n = C.c_int(0)
    check(lib().rgc_device_count(C.byref(n)))
    return n.value


def device_info():
    d, s, b = C.c_int(), C.c_int(), _sz()
    check(lib().rgc_device_info(C.byref(d), C.byref(s), C.byref(b)))
    return d.value, s.value, b.value


def stream_handle() -> int:
    s = _vp()
    check(lib().rgc_stream(C.byref(s)))
    return s.value or 0


def synchronize() -> None:
    check(lib().rgc_synchronize())


def launch_count() -> int:
    return int(lib().rgc_launch_count())


def last_kernel_ms():
    ms = (C.c_float * 2)()
    check(lib().rgc_last_kernel_ms(ms))
    return float(ms[0]), float(ms[1])


def last_kernel_times():
    """(total, dominant kernel, prologue kernel, sort kernels) of the last hot-path call, ms"""
    ms = (C.c_float * 4)()
    check(lib().rgc_last_kernel_times(ms, 4))
    return tuple(float(x) for x in ms)


def comm_exchange_kind() -> str:
    """how result vectors are combined across ranks"""
    k = C.c_int()
    check(lib().rgc_comm_exchange_kind(C.byref(k)))
    return {0: "single rank", 1: "ncclAllReduce", 2: "peer-store exchange over NVLink"}[k.value]


def sort_rank_mode() -> int:
    """1: atomic ranking verified by the on-device probe; 0: ballot ranking; -1: not decided yet"""
    m = C.c_int()
    check(lib().rgc_sort_rank_mode(C.byref(m)))
    return m.value


def pair_plan_describe(bins, tab_x, tab_y) -> dict:
    """host-only: how the hinge path lays the photon bins out in lane groups (no device needed)"""
    bins = np.ascontiguousarray(bins, np.float32)
    tx = np.ascontiguousarray(tab_x, np.float32)
    ty = np.ascontiguousarray(tab_y, np.float32)
    info = (C.c_int * 8)()
    phase = C.c_float()
    cap = 32 * 128
    slot_bin = (C.c_int * cap)()
    check(lib().rgc_pair_plan_describe(bins.ctypes.data_as(_f32p), bins.size, tx.ctypes.data_as(_f32p),
                                       ty.ctypes.data_as(_f32p), tx.size, info, C.byref(phase), slot_bin, cap))
    keys = ("eligible", "groups", "slots", "buckets", "chunks", "buckets_with_extension",
            "most_groups_per_sub_bucket", "sub_buckets")
    out = dict(zip(keys, list(info)))
    out["phase"] = float(phase.value)
    out["slot_bin"] = np.array(slot_bin[:out["slots"]], np.int32)
    return out


def trim_memory() -> None:
    """return the cached (released) particle columns to the device"""
    check(lib().rgc_trim_memory())


def last_pair_ontable_evals() -> float:
    """pairs of the last hinge launch that are on the F table (call right after the spectrum)"""
    v = C.c_double()
    check(lib().rgc_last_pair_ontable_evals(C.byref(v)))
    return float(v.value)


def last_pair_lane_evals() -> float:
    """hinge evaluations the pair kernel issued in the last particle-spectrum call"""
    v = C.c_double()
    check(lib().rgc_last_pair_lane_evals(C.byref(v)))
    return float(v.value)


PEAK_FFMA, PEAK_PAIR, PEAK_LDS64, PEAK_HBM_READ, PEAK_INT = range(5)


def measure_peak(kind: int) -> float:
    """On-device roofline denominators (include/ragnar_cuda.h rgc_measure_peak)."""
    v, clk = C.c_double(), C.c_double()
    check(lib().rgc_measure_peak(kind, C.byref(v), C.byref(clk)))
    return float(v.value)


class PinnedArray:
    """A numpy view over cudaHostAlloc'ed memory (freed on close/GC)."""

    def __init__(self, n: int, dtype=np.float32):
        self._ptr = _vp()
        self.nbytes = int(n) * np.dtype(dtype).itemsize
        check(lib().rgc_host_alloc(self.nbytes, C.byref(self._ptr)))
        buf = (C.c_char * self.nbytes).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=n)

    def close(self):
        if self._ptr is not None and self._ptr.value:
            self.array = None
            lib().rgc_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------- comm
def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(COMM_ID_BYTES)
    check(lib().rgc_comm_get_unique_id(buf))
    return buf.raw


def comm_init(uid: bytes, rank: int, nranks: int) -> None:
    assert len(uid) == COMM_ID_BYTES
    check(lib().rgc_comm_init(uid, rank, nranks))


def comm_destroy() -> None:
    check(lib().rgc_comm_destroy())


def comm_info():
    r, n = C.c_int(), C.c_int()
    check(lib().rgc_comm_info(C.byref(r), C.byref(n)))
    return r.value, n.value


# -------------------------------------------------------------- particles
class Particles:
    """Thin owner of an rgc_particles_t handle (SoA device columns)."""

    def __init__(self, dim: int = 3):
        self.h = _vp()
        check(lib().rgc_particles_create(dim, C.byref(self.h)))
        self.n = 0

    def allocate(self, nalloc: int, with_coords: bool = False):
        check(lib().rgc_particles_allocate(self.h, nalloc, int(with_coords)))
        return self

    def write(self, quantity: int, comp: int, start: int, host: np.ndarray):
        assert host.dtype == np.float32 and host.flags.c_contiguous
        check(lib().rgc_particles_write(self.h, quantity, comp, start,
                                        host.ctypes.data_as(_vp), host.size))

    def read(self, quantity: int, comp: int, start: int, n: int) -> np.ndarray:
        out = np.empty(n, np.float32)
        check(lib().rgc_particles_read(self.h, quantity, comp, start, n, out.ctypes.data_as(_vp)))
        return out

    def from_columns(self, U=None, E=None, B=None):
        """U, E, B: sequences of three float arrays (None = zeros)."""
        n = len(next(c for q in (U, E, B) if q is not None for c in q if c is not None))
        self.allocate(n)
        keep = []
        for qid, q in ((Q_U, U), (Q_E, E), (Q_B, B)):
            if q is None:
                continue
            for d, col in enumerate(q):
                if col is not None:
                    arr = _f32(col)
                    keep.append(arr)
                    self.write(qid, d, 0, arr)
        synchronize()
        self.n = n
        return self

    def generate(self, kind: int, seed: int, global_offset: int, start: int, n: int,
                 umin: float = 1.0, umax: float = 100.0):
        check(lib().rgc_particles_generate(self.h, kind, seed, global_offset, start, n, umin, umax))
        self.n = max(self.n, start + n)
        return self

    def nalloc(self) -> int:
        return int(lib().rgc_particles_nalloc(self.h))

    def release(self):
        if self.h is not None and self.h.value:
            lib().rgc_particles_release(self.h)
            self.h = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


# ----------------------------------------------------------- host-exact bits
def logspace(start, stop, num) -> np.ndarray:
    out = np.empty(num, np.float32)
    check(lib().rgc_logspace(start, stop, num, _ptr(out)))
    return out


def linspace(start, stop, num) -> np.ndarray:
    out = np.empty(num, np.float32)
    check(lib().rgc_linspace(start, stop, num, _ptr(out)))
    return out


class DeviceArray:
    """Owner of an rgc_buf_t handle (a device-resident 1-D float array)."""

    def __init__(self, handle):
        self.h = handle

    @classmethod
    def from_host(cls, a) -> "DeviceArray":
        a = _f32(a)
        h = _vp()
        check(lib().rgc_buf_from_host(1, a.ctypes.data_as(_vp), a.size, C.byref(h)))
        return cls(h)

    def __len__(self) -> int:
        return int(lib().rgc_buf_size(self.h))

    def to_host(self) -> np.ndarray:
        out = np.empty(len(self), np.float32)
        if len(self):
            check(lib().rgc_buf_to_host(self.h, 0, len(self), out.ctypes.data_as(_vp)))
        return out

    def minmax(self):
        mn, mx = C.c_float(), C.c_float()
        check(lib().rgc_buf_minmax(self.h, C.byref(mn), C.byref(mx)))
        return mn.value, mx.value

    def release(self):
        if self.h is not None and self.h.value:
            lib().rgc_buf_release(self.h)
            self.h = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


def linspace_device(start, stop, num) -> DeviceArray:
    h = _vp()
    check(lib().rgc_linspace_device(start, stop, num, C.byref(h)))
    return DeviceArray(h)


def logspace_device(start, stop, num) -> DeviceArray:
    h = _vp()
    check(lib().rgc_logspace_device(start, stop, num, C.byref(h)))
    return DeviceArray(h)


def tabulated_eval(loggrid: bool, tab_x: DeviceArray, tab_y: DeviceArray, x0: DeviceArray,
                   yfill: float = 0.0) -> DeviceArray:
    """InterpolateTabulatedFunction<LG> of a device-resident table at every element of x0"""
    h = _vp()
    check(lib().rgc_tabulated_eval(int(loggrid), tab_x.h, tab_y.h, yfill, x0.h, C.byref(h)))
    return DeviceArray(h)


def ffunc_integrand(x: float) -> float:
    out = C.c_float()
    check(lib().rgc_sync_ffunc_integrand(x, C.byref(out)))
    return out.value


def tabulate_ffunc(n=200, xmin=1e-6, xmax=100.0):
    xs, ys = np.empty(n, np.float32), np.empty(n, np.float32)
    check(lib().rgc_sync_tabulate_ffunc(n, xmin, xmax, _ptr(xs), _ptr(ys)))
    return xs, ys


def interpolate(x0, x, y, loggrid=True, yfill=0.0) -> float:
    x, y = _f32(x), _f32(y)
    out = C.c_float()
    check(lib().rgc_interpolate(int(loggrid), x0, _ptr(x), _ptr(y), len(x), yfill, C.byref(out)))
    return out.value


def generator_eval(kind: int, params, energy) -> np.ndarray:
    prm, e = _f32(params), _f32(energy)
    out = np.empty_like(e)
    check(lib().rgc_generator_eval(kind, _ptr(prm), _ptr(e), len(e), _ptr(out)))
    return out


# ----------------------------------------------------------------- hot path
def energy_histogram(p: Particles, bins, log_spaced: bool, fourvel: bool = True, nactive=None,
                     want_counts=True):
    """-> (hist f32, counts u64 | None, sum f64)"""
    bins = _f32(bins)
    n = len(bins)
    hist = np.zeros(n, np.float32)
    cnt = np.zeros(n, np.uint64) if want_counts else None
    s64 = np.zeros(n, np.float64)
    check(lib().rgc_energy_histogram(p.h, p.n if nactive is None else nactive, _ptr(bins), n,
                                     int(log_spaced), int(fourvel), _ptr(hist),
                                     _ptr(cnt, _u64p) if want_counts else None,
                                     _ptr(s64, _f64p)))
    return hist, cnt, s64


def sync_spectrum_particles(p: Particles, bins_e_syn, B0, g_syn, e_at, table=None, nactive=None):
    """-> (spec f32, spec f64)"""
    bins = _f32(bins_e_syn)
    tx, ty = table if table is not None else tabulate_ffunc()
    tx, ty = _f32(tx), _f32(ty)
    s32 = np.zeros(len(bins), np.float32)
    s64 = np.zeros(len(bins), np.float64)
    check(lib().rgc_sync_spectrum_particles(p.h, p.n if nactive is None else nactive, _ptr(bins),
                                            len(bins), _ptr(tx), _ptr(ty), len(tx), B0, g_syn,
                                            e_at, _ptr(s32), _ptr(s64, _f64p)))
    return s32, s64


def hist_and_spectrum(p: Particles, gamma_bins, log_spaced: bool, fourvel: bool, bins_e_syn, B0, g_syn,
                      e_at, table=None, nactive=None):
    """energy histogram + particle spectrum in one call (kernels back to back on the stream, one
    wait) -> (hist f32, hist f64, spec f32, spec f64)"""
    gb, bins = _f32(gamma_bins), _f32(bins_e_syn)
    tx, ty = table if table is not None else tabulate_ffunc()
    tx, ty = _f32(tx), _f32(ty)
    h32, h64 = np.zeros(len(gb), np.float32), np.zeros(len(gb), np.float64)
    s32, s64 = np.zeros(len(bins), np.float32), np.zeros(len(bins), np.float64)
    check(lib().rgc_hist_and_spectrum(p.h, p.n if nactive is None else nactive, _ptr(gb), len(gb),
                                      int(log_spaced), int(fourvel), _ptr(h32), _ptr(h64, _f64p),
                                      _ptr(bins), len(bins), _ptr(tx), _ptr(ty), len(tx), B0, g_syn, e_at,
                                      _ptr(s32), _ptr(s64, _f64p)))
    return h32, h64, s32, s64


def sync_spectrum_dist(gbeta, f, islog, bins_e_syn, g_syn, e_at, table=None):
    """-> (spec f32, spec f64)"""
    gbeta, f, bins = _f32(gbeta), _f32(f), _f32(bins_e_syn)
    tx, ty = table if table is not None else tabulate_ffunc()
    tx, ty = _f32(tx), _f32(ty)
    s32 = np.zeros(len(bins), np.float32)
    s64 = np.zeros(len(bins), np.float64)
    check(lib().rgc_sync_spectrum_dist(_ptr(gbeta), _ptr(f), len(gbeta), int(islog), _ptr(bins),
                                       len(bins), _ptr(tx), _ptr(ty), len(tx), g_syn, e_at,
                                       _ptr(s32), _ptr(s64, _f64p)))
    return s32, s64


def sync_spectrum_dist_batch(gbeta, f_batch, islog, bins_e_syn, g_syn, e_at, table=None, mode=-1):
    """f_batch: [nbatch, len(gbeta)] distributions on shared bins -> (spec f32, spec f64), each
    [nbatch, len(bins)].  mode 0 literal terms, 1 kernel-matrix contraction, -1 automatic."""
    gbeta, bins = _f32(gbeta), _f32(bins_e_syn)
    fb = np.ascontiguousarray(f_batch, dtype=np.float32)
    if fb.ndim != 2 or fb.shape[1] != len(gbeta):
        raise ValueError("f_batch must be [nbatch, len(gbeta)]")
    tx, ty = table if table is not None else tabulate_ffunc()
    tx, ty = _f32(tx), _f32(ty)
    s32 = np.zeros((fb.shape[0], len(bins)), np.float32)
    s64 = np.zeros((fb.shape[0], len(bins)), np.float64)
    check(lib().rgc_sync_spectrum_dist_batch(_ptr(gbeta), _ptr(fb), fb.shape[0], len(gbeta), int(islog),
                                             _ptr(bins), len(bins), _ptr(tx), _ptr(ty), len(tx), g_syn,
                                             e_at, mode, _ptr(s32), _ptr(s64, _f64p)))
    return s32, s64


def ic_spectrum(g_prtls, f_prtls, islog, e_soft, f_soft, bins_e_ic):
    """ICSpectrum (reference src/physics/ic.cpp:15-46) -> (spec_f32, spec_f64)"""
    g, f, es, fs, b = (_f32(a) for a in (g_prtls, f_prtls, e_soft, f_soft, bins_e_ic))
    if len(g) != len(f) or len(es) != len(fs):
        raise ValueError("distribution arrays of unequal length")
    s32 = np.zeros(len(b), np.float32)
    s64 = np.zeros(len(b), np.float64)
    check(lib().rgc_ic_spectrum(_ptr(g), _ptr(f), len(g), int(islog), _ptr(es), _ptr(fs), len(es),
                                _ptr(b), len(b), _ptr(s32), _ptr(s64, _f64p)))
    return s32, s64


# ------------------------------------------------------------------ plugin
def tristan_write_species(path: str, step: int, sp: int, columns, with_coords: bool,
                          append: bool) -> None:
    """columns: list of float32 arrays (or None) in the order x,y,z (if with_coords),
    u,v,w, ex,ey,ez, bx,by,bz; all of one length."""
    n = len(next(c for c in columns if c is not None))
    arrs = [None if c is None else _f32(c) for c in columns]
    ptrs = (_f32p * len(arrs))(*[_f32p() if a is None else _ptr(a) for a in arrs])
    check(lib().rgc_tristan_write_species(path.encode(), step, sp, n, int(with_coords), ptrs,
                                          int(append)))


def tristan_read_particles(path: str, step: int, sp: int, start=0, size=0, stride=1,
                           ignore_coords=False, dim=3):
    p = Particles.__new__(Particles)
    p.h = _vp()
    ntotal, nread = _sz(), _sz()
    check(lib().rgc_tristan_read_particles(path.encode(), step, sp, start, size, stride,
                                           int(ignore_coords), dim, C.byref(p.h),
                                           C.byref(ntotal), C.byref(nread)))
    p.n = nread.value
    return p, ntotal.value


def tristan_read_range(path: str, step: int, sp: int, start: int, count: int,
                       ignore_coords=False, dim=3):
    """exactly particles [start, start + count) (sharded multi-GPU reads) -> (Particles, ntotal)"""
    p = Particles.__new__(Particles)
    p.h = _vp()
    ntotal = _sz()
    check(lib().rgc_tristan_read_range(path.encode(), step, sp, start, count, int(ignore_coords),
                                       dim, C.byref(p.h), C.byref(ntotal)))
    p.n = count
    return p, ntotal.value


# ------------------------------------------------------------------ HDF5 (host side)
_NP_OF = {I32: np.int32, F32: np.float32, F64: np.float64}
_DTYPE_OF = {np.dtype(np.int32): I32, np.dtype(np.float32): F32,
             np.dtype(np.float64): F64}


class H5File:
    """Host-side view of the library's HDF5 layer (rgc_h5_*): no GPU required.
    mode: "r" (ReadOnly), "a" (ReadWrite | Create), "w" (truncate)."""

    def __init__(self, filename: str, mode: str = "r"):
        self.h = _vp()
        check(lib().rgc_h5_open(str(filename).encode(), {"r": 0, "a": 1, "w": 2}[mode],
                                C.byref(self.h)))

    def close(self) -> None:
        if self.h:
            h, self.h = self.h, _vp()
            check(lib().rgc_h5_close(h))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def list(self, group: str = "/"):
        need = _sz()
        check(lib().rgc_h5_list(self.h, group.encode(), None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        check(lib().rgc_h5_list(self.h, group.encode(), buf, need.value, None))
        return [s for s in buf.value.decode().split("\n") if s]

    def info(self, name: str) -> dict:
        rank, cls, es, layout = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        dims = (C.c_uint64 * 8)()
        check(lib().rgc_h5_dataset_info(self.h, name.encode(), C.byref(rank), dims, 8,
                                        C.byref(cls), C.byref(es), C.byref(layout)))
        return {"dims": [int(dims[k]) for k in range(rank.value)], "class": cls.value,
                "elem_size": es.value, "layout": layout.value}

    def read(self, name: str, start: int = 0, count: int | None = None, stride: int = 1,
             dtype=np.float32) -> np.ndarray:
        if count is None:
            dims = self.info(name)["dims"]
            total = int(np.prod(dims)) if dims else 1
            count = (total - start + stride - 1) // stride if total > start else 0
        out = np.empty(count, dtype)
        check(lib().rgc_h5_read(self.h, name.encode(), start, count, stride,
                                _DTYPE_OF[np.dtype(dtype)], out.ctypes.data_as(_vp)))
        return out

    def create_dataset(self, name: str, dtype, n: int) -> None:
        check(lib().rgc_h5_create_dataset(self.h, name.encode(), _DTYPE_OF[np.dtype(dtype)], n))

    def write(self, name: str, data: np.ndarray, start: int = 0) -> None:
        data = np.ascontiguousarray(data)
        check(lib().rgc_h5_write(self.h, name.encode(), start, data.size,
                                 _DTYPE_OF[data.dtype], data.ctypes.data_as(_vp)))
