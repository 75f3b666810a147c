// Runtime of libragnar_cuda.so: context, errors, Array1D buffers, the SoA
// Particles container, host<->device staging and the NCCL communicator.
//
// Replaces (reference paths relative to haykh/ragnar @ fceb6b08):
//   Kokkos::initialize/finalize            src/pyinterface.cpp:30-49
//   Kokkos::View<T*> + mirror/deep_copy    src/containers/array.{hpp,cpp}
//   Particles<D> storage                   src/containers/particles.cpp:115-187,346-383
// Data layout in HBM: every particle quantity component is its own contiguous
// float column (SoA), 256-byte aligned, zero-initialised, so that a warp reads
// 32 consecutive particles of one component with one 128-byte request and a
// thread reads 4 consecutive particles with one 16-byte load.
#include "rgc_internal.hpp"

#include <dlfcn.h>
#include <immintrin.h>
#include <execinfo.h>
#include <csignal>
#include <nccl.h> // types and enum values only; the library is dlopen'ed

#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <new>
#include <string>
#include <thread>

namespace rgc {

  // ------------------------------------------------------------------ errors
  static thread_local std::string t_last_error;

  int fail(int code, const char* fmt, ...) {
    char    buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    t_last_error = buf;
    return code;
  }

  void clear_error() { t_last_error.clear(); }

  Context& ctx() {
    static Context c;
    return c;
  }

  int ensure_scratch(std::size_t bytes, void** out) {
    auto& c = ctx();
    if (bytes > c.scratch_bytes) {
      if (c.scratch) {
        RGC_CUDA(cudaStreamSynchronize(c.stream));
        RGC_CUDA(cudaFree(c.scratch));
        c.scratch       = nullptr;
        c.scratch_bytes = 0;
      }
      const std::size_t want = (bytes + (std::size_t(1) << 20) - 1) & ~((std::size_t(1) << 20) - 1);
      RGC_CUDA(cudaMalloc(&c.scratch, want));
      c.scratch_bytes = want;
    }
    *out = c.scratch;
    return RGC_OK;
  }

  int ensure_result(std::size_t bytes, void** out) {
    auto& c = ctx();
    if (bytes > c.result_bytes) {
      if (c.result) {
        RGC_CUDA(cudaStreamSynchronize(c.stream));
        RGC_CUDA(cudaFree(c.result));
        c.result       = nullptr;
        c.result_bytes = 0;
      }
      const std::size_t want = (bytes + 65535) & ~std::size_t(65535);
      RGC_CUDA(cudaMalloc(&c.result, want));
      c.result_bytes = want;
    }
    *out = c.result;
    return RGC_OK;
  }

  // ------------------------------------------------------------ H2D staging
  static int ensure_stages() {
    auto& c = ctx();
    for (int s = 0; s < kNumStages; ++s) {
      if (!c.stage[s]) {
        RGC_CUDA(cudaHostAlloc(&c.stage[s], kStageBytes, cudaHostAllocDefault));
        RGC_CUDA(cudaEventCreateWithFlags(&c.stage_free[s], cudaEventDisableTiming));
        RGC_CUDA(cudaEventRecord(c.stage_free[s], c.stream));
      }
    }
    return RGC_OK;
  }

  static bool is_pinned(const void* p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    return attr.type == cudaMemoryTypeHost;
  }

  // memcpy into a pinned stage with non-temporal stores: the destination is read next by the
  // DMA engine, never by this core, so the read-for-ownership of a cached store (a third of
  // the host memory traffic of the copy) is skipped.  dst must be 32-byte aligned.
  __attribute__((target("avx2"))) static void stream_copy_avx2(char* dst, const char* src, std::size_t bytes) {
    std::size_t i = 0;
    for (; i + 128 <= bytes; i += 128) {
      const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
      const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 32));
      const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 64));
      const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 96));
      _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), a);
      _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 32), b);
      _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 64), c);
      _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 96), d);
    }
    _mm_sfence();
    if (i < bytes) {
      std::memcpy(dst + i, src + i, bytes - i);
    }
  }

  static void stage_copy(char* dst, const char* src, std::size_t bytes) {
    // RGC_STAGE_NT=1 selects non-temporal stores: faster in isolation (tools/mb_stage.cpp, 16
    // threads: 83 GB/s against 67 with memcpy) but they push every stage out to DRAM, where the
    // DMA engine then has to fetch it; off by default (see StageGeometry below)
    static const bool avx2 = __builtin_cpu_supports("avx2") && [] {
      const char* e = std::getenv("RGC_STAGE_NT");
      return e && e[0] == '1';
    }();
    if (avx2 && (reinterpret_cast<std::uintptr_t>(dst) & 31u) == 0 && bytes >= 4096) {
      stream_copy_avx2(dst, src, bytes);
    } else {
      std::memcpy(dst, src, bytes);
    }
  }

  // Staging pool for large pageable sources.  Every worker owns two pinned stages and works
  // on its own: claim the next chunk of the source (atomic counter), wait until its stage's
  // previous DMA has finished, copy the chunk in, enqueue the DMA on the caller's stream,
  // record the stage's event.  No barrier between chunks — a worker's copy of chunk k+1
  // overlaps the DMA of its chunk k and everybody else's — one fork / join per copy_h2d call.
  // Created on first use, joined in rgc_finalize.
  // Geometry of the pool: the stages must stay in the last-level cache.  A chunk is written by
  // a core (plain stores) and read once by the DMA engine; while all stages together fit the
  // L3, that read is served from the cache and host DRAM only sees the source being read —
  // one pass instead of three.  Measured (tools/bench_fromarrays.py, bench.py e2e): on one
  // bench box 8 MiB stages written with non-temporal stores gave 71 ms for 3.6 GB, on the next
  // one 110 ms, where 8 workers x 2 x 2 MiB of plain stores gave 73 ms; with two ranks sharing
  // a host 122 ms against 86 ms (12 workers x 2 x 1 MiB each).  So: budget = 0.6 x L3 / ranks
  // on the node, at most 8 workers, chunks of budget / (2 x workers) within [1, 8] MiB (below
  // 1 MiB the per-copy cost of the DMA engine shows), fewer workers when even that does not
  // fit.  RGC_COPY_THREADS / RGC_STAGE_CHUNK_KB override; RGC_STAGE_NT=1 selects non-temporal stores.
  struct StageGeometry {
    int         workers;
    std::size_t chunk;
  };

  static std::size_t l3_bytes() {
    std::size_t total = 0;
    std::vector<std::string> seen;
    for (int cpu = 0; cpu < 1024; ++cpu) {
      const std::string base = "/sys/devices/system/cpu/cpu" + std::to_string(cpu) + "/cache/index3/";
      FILE* f = std::fopen((base + "shared_cpu_list").c_str(), "r");
      if (!f) {
        if (cpu == 0) {
          break;
        }
        continue;
      }
      char who[256] = { 0 };
      const bool got = std::fgets(who, sizeof(who), f) != nullptr;
      std::fclose(f);
      if (!got || std::find(seen.begin(), seen.end(), std::string(who)) != seen.end()) {
        continue; // this L3 instance is counted already
      }
      seen.emplace_back(who);
      if (FILE* g = std::fopen((base + "size").c_str(), "r")) {
        unsigned long v = 0;
        char          unit = 'K';
        if (std::fscanf(g, "%lu%c", &v, &unit) >= 1) {
          total += (std::size_t)v << (unit == 'M' ? 20 : (unit == 'G' ? 30 : 10));
        }
        std::fclose(g);
      }
    }
    return total ? total : (std::size_t(32) << 20);
  }

  static const StageGeometry& stage_geometry() {
    static const StageGeometry g = [] {
      const int   hw   = std::max(1, (int)std::thread::hardware_concurrency());
      const char* lws  = std::getenv("LOCAL_WORLD_SIZE");
      const int   rpn  = std::max(1, lws ? std::atoi(lws) : 1); // ranks on this node (torchrun)
      const std::size_t MiB = std::size_t(1) << 20;
      const std::size_t budget = std::max(4 * MiB, (std::size_t)(0.6 * (double)l3_bytes()) / rpn);
      StageGeometry     sg;
      sg.workers = std::min(8, std::max(2, hw / rpn));
      if (const char* e = std::getenv("RGC_COPY_THREADS")) {
        sg.workers = std::max(1, std::min(std::atoi(e), 64));
      }
      sg.chunk = std::min(8 * MiB, std::max(MiB, budget / (2 * (std::size_t)sg.workers) / (256 << 10) * (256 << 10)));
      if (!std::getenv("RGC_COPY_THREADS") && 2 * (std::size_t)sg.workers * sg.chunk > budget) {
        sg.workers = (int)std::max<std::size_t>(2, budget / (2 * sg.chunk));
      }
      if (const char* e = std::getenv("RGC_STAGE_CHUNK_KB")) {
        const long kb = std::atol(e);
        if (kb >= 64 && kb <= (1 << 16)) {
          sg.chunk = (std::size_t)kb << 10;
        }
      }
      return sg;
    }();
    return g;
  }
#define kPoolChunk (stage_geometry().chunk)
  class CopyPool {
  public:
    explicit CopyPool(int n) : workers_(n) {
      for (int i = 0; i < n; ++i) {
        workers_[i].th = std::thread([this, i] { loop(i); });
      }
    }
    ~CopyPool() {
      {
        std::lock_guard<std::mutex> lk(m_);
        stop_ = true;
      }
      start_.notify_all();
      for (auto& w : workers_) {
        w.th.join();
      }
      for (auto& w : workers_) {
        for (int s = 0; s < 2; ++s) {
          if (w.stage[s]) {
            cudaFreeHost(w.stage[s]);
            cudaEventDestroy(w.free_ev[s]);
          }
        }
      }
    }
    // dst (device) <- src (pageable host), enqueued on `stream`; returns when every chunk has
    // been staged and its DMA enqueued (the source may then be released; the stages are the
    // pool's own).  Returns a cudaError_t as int (0 = ok).
    int copy(void* dst, const void* src, std::size_t bytes, cudaStream_t stream, int device) {
      std::unique_lock<std::mutex> lk(m_);
      dst_     = static_cast<char*>(dst);
      src_     = static_cast<const char*>(src);
      bytes_   = bytes;
      stream_  = stream;
      device_  = device;
      nchunks_ = (bytes + kPoolChunk - 1) / kPoolChunk;
      next_.store(0);
      err_.store(0);
      pending_ = (int)workers_.size();
      ++gen_;
      start_.notify_all();
      done_.wait(lk, [this] { return pending_ == 0; });
      return err_.load();
    }

  private:
    struct Worker {
      std::thread th;
      void*       stage[2] { nullptr, nullptr };
      cudaEvent_t free_ev[2] { nullptr, nullptr };
      int         next { 0 };
    };
    void loop(int id) {
      Worker&       w    = workers_[id];
      std::uint64_t seen = 0;
      bool          ready = false;
      for (;;) {
        std::unique_lock<std::mutex> lk(m_);
        start_.wait(lk, [&] { return stop_ || gen_ != seen; });
        if (stop_) {
          return;
        }
        seen = gen_;
        char*             d = dst_;
        const char*       s = src_;
        const std::size_t b = bytes_, nch = nchunks_;
        cudaStream_t      st = stream_;
        const int         dev = device_;
        lk.unlock();
        cudaError_t e = cudaSuccess;
        if (!ready) { // first job of this thread: its device and its two stages
          e = cudaSetDevice(dev);
          for (int k = 0; k < 2 && e == cudaSuccess; ++k) {
            e = cudaHostAlloc(&w.stage[k], kPoolChunk, cudaHostAllocDefault);
            if (e == cudaSuccess) {
              e = cudaEventCreateWithFlags(&w.free_ev[k], cudaEventDisableTiming);
            }
          }
          ready = e == cudaSuccess;
        }
        while (e == cudaSuccess) {
          const std::size_t c = next_.fetch_add(1);
          if (c >= nch) {
            break;
          }
          const std::size_t off = c * kPoolChunk, len = std::min(kPoolChunk, b - off);
          const int         k   = w.next;
          e = cudaEventSynchronize(w.free_ev[k]); // never recorded yet: returns at once
          if (e != cudaSuccess) {
            break;
          }
          stage_copy(static_cast<char*>(w.stage[k]), s + off, len);
          e = cudaMemcpyAsync(d + off, w.stage[k], len, cudaMemcpyHostToDevice, st);
          if (e == cudaSuccess) {
            e = cudaEventRecord(w.free_ev[k], st);
          }
          w.next ^= 1;
        }
        if (e != cudaSuccess) {
          err_.store((int)e);
          next_.store(nch); // the others stop claiming chunks
        }
        lk.lock();
        if (--pending_ == 0) {
          done_.notify_one();
        }
      }
    }
    std::vector<Worker>      workers_;
    std::mutex               m_;
    std::condition_variable  start_, done_;
    std::uint64_t            gen_ { 0 };
    int                      pending_ { 0 };
    bool                     stop_ { false };
    char*                    dst_ { nullptr };
    const char*              src_ { nullptr };
    std::size_t              bytes_ { 0 }, nchunks_ { 0 };
    cudaStream_t             stream_ { nullptr };
    int                      device_ { 0 };
    std::atomic<std::size_t> next_ { 0 };
    std::atomic<int>         err_ { 0 };
  };

  static CopyPool*  g_copy_pool = nullptr;
  static std::mutex g_stage_mutex; // the pinned ring and the pool serve one copy at a time
  static int        g_next_stage = 0;

  static CopyPool* copy_pool() {
    if (!g_copy_pool) {
      g_copy_pool = new CopyPool(stage_geometry().workers);
    }
    return g_copy_pool;
  }

  void copy_pool_release() {
    std::lock_guard<std::mutex> lk(g_stage_mutex);
    delete g_copy_pool;
    g_copy_pool = nullptr;
  }

  // Pinned sources go out as ONE async DMA (caller keeps them alive until the
  // stream is synchronised); large pageable sources go through the staging pool above,
  // small ones (< 4 MiB: bin edges, tables, plans) through a ring of pinned stages owned by
  // the calling thread's context, whose position persists between calls so that a run of
  // small copies does not wait for the previous one's DMA.
  int copy_h2d(void* dst, const void* src, std::size_t bytes, cudaStream_t stream) {
    if (bytes == 0) {
      return RGC_OK;
    }
    if (is_pinned(src)) {
      RGC_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream));
      return RGC_OK;
    }
    std::lock_guard<std::mutex> lk(g_stage_mutex);
    if (bytes >= (std::size_t(4) << 20)) {
      const int e = copy_pool()->copy(dst, src, bytes, stream, ctx().device);
      if (e != 0) {
        cudaGetLastError();
        return fail(e == (int)cudaErrorMemoryAllocation ? RGC_ERR_OOM : RGC_ERR_CUDA,
                    "staged host-to-device copy failed: %s", cudaGetErrorString((cudaError_t)e));
      }
      return RGC_OK;
    }
    RGC_TRY(ensure_stages());
    auto&       c    = ctx();
    std::size_t done = 0;
    while (done < bytes) {
      const int         s     = g_next_stage;
      const std::size_t chunk = bytes - done < kStageBytes ? bytes - done : kStageBytes;
      RGC_CUDA(cudaEventSynchronize(c.stage_free[s]));
      std::memcpy(c.stage[s], static_cast<const char*>(src) + done, chunk);
      RGC_CUDA(cudaMemcpyAsync(static_cast<char*>(dst) + done, c.stage[s], chunk,
                               cudaMemcpyHostToDevice, stream));
      RGC_CUDA(cudaEventRecord(c.stage_free[s], stream));
      done += chunk;
      g_next_stage = (s + 1) % kNumStages;
    }
    return RGC_OK;
  }

  // -------------------------------------------------------------------- NCCL
  struct NcclApi {
    void* handle { nullptr };
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) { nullptr };
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) { nullptr };
    ncclResult_t (*CommDestroy)(ncclComm_t) { nullptr };
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t,
                              ncclComm_t, cudaStream_t) { nullptr };
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) {
      nullptr
    };
    ncclResult_t (*GroupStart)() { nullptr };
    ncclResult_t (*GroupEnd)() { nullptr };
    const char* (*GetErrorString)(ncclResult_t) { nullptr };
  };

  static NcclApi& nccl() {
    static NcclApi api;
    return api;
  }

  static int load_nccl() {
    auto& api = nccl();
    if (api.handle) {
      return RGC_OK;
    }
    // A process that already loaded NCCL (e.g. through torch) gets that copy.
    const char* names[] = { "libnccl.so.2", "libnccl.so" };
    for (const char* nm : names) {
      api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) {
        break;
      }
    }
    if (!api.handle) {
      return fail(RGC_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
    }
#define RGC_NCCL_SYM(field, sym)                                                    \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, sym));        \
  if (!api.field) {                                                                 \
    return fail(RGC_ERR_NCCL, "libnccl lacks symbol %s", sym);                      \
  }
    RGC_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
    RGC_NCCL_SYM(CommInitRank, "ncclCommInitRank");
    RGC_NCCL_SYM(CommDestroy, "ncclCommDestroy");
    RGC_NCCL_SYM(AllReduce, "ncclAllReduce");
    RGC_NCCL_SYM(AllGather, "ncclAllGather");
    RGC_NCCL_SYM(GroupStart, "ncclGroupStart");
    RGC_NCCL_SYM(GroupEnd, "ncclGroupEnd");
    RGC_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef RGC_NCCL_SYM
    return RGC_OK;
  }

#define RGC_NCCL(expr)                                                              \
  do {                                                                              \
    ncclResult_t rgc_nr__ = (expr);                                                 \
    if (rgc_nr__ != ncclSuccess) {                                                  \
      return ::rgc::fail(RGC_ERR_NCCL, "%s failed: %s", #expr,                      \
                         ::rgc::nccl().GetErrorString(rgc_nr__));                   \
    }                                                                               \
  } while (0)

  // ------------------------------------------------- peer-store all-reduce (NVLink)
  // Layout of every rank's exchange buffer: two sets (call parity) of
  //   [kXchgMaxRanks][kXchgSlot] 8-byte elements   — slot r is written by rank r
  //   [kXchgMaxRanks] flags, 128 B apart           — flag r = sequence number of rank r's data
  // A call stores this rank's vector into slot `rank` of EVERY rank (its own included),
  // fences, raises its flag everywhere, waits for all flags of its own buffer and sums
  // the slots in rank order: the same order on every rank, so all ranks get bit-identical
  // sums.  The other parity's set is only rewritten two calls later, when every peer has
  // provably finished reading it (a rank raises flag k+1 after its call-k sum, in stream
  // order).  One CTA; payloads <= kXchgSlot elements (larger ones go through NCCL).
  constexpr int         kXchgMaxRanks = 8;
  constexpr std::size_t kXchgSlot     = 8192; // elements
  constexpr std::size_t kXchgSetBytes = kXchgMaxRanks * kXchgSlot * 8 + kXchgMaxRanks * 128;
  constexpr std::size_t kXchgBytes    = 2 * kXchgSetBytes;

  struct XchgParams {
    unsigned long long* peer[kXchgMaxRanks];
    int                 rank, nranks;
    unsigned long long  seq;
    unsigned long long* data; // in / out, n elements: [0, n_u64) summed as u64, the rest as f64
    int                 n, n_u64;
    unsigned long long  timeout_ns;
    int*                status; // host-mapped word: 1 + rank of a peer that never arrived
  };

  __device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
  }

  __global__ void __launch_bounds__(256) xchg_allreduce_kernel(const XchgParams P) {
    __shared__ int    s_late;
    const std::size_t set_off  = (P.seq & 1ull) * (kXchgSetBytes / 8);
    const std::size_t flag_off = set_off + (std::size_t)kXchgMaxRanks * kXchgSlot;
    if (threadIdx.x == 0) {
      s_late = 0;
    }
    // 1. my vector into slot `rank` of every rank
    for (int p = 0; p < P.nranks; ++p) {
      unsigned long long* dst = P.peer[p] + set_off + (std::size_t)P.rank * kXchgSlot;
      for (int i = threadIdx.x; i < P.n; i += blockDim.x) {
        dst[i] = P.data[i];
      }
    }
    __threadfence_system();
    __syncthreads();
    // 2. raise my flag everywhere, wait for everyone's flag here.  The wait is bounded in
    // TIME (RGC_XCHG_TIMEOUT_MS, default 10 min — a peer may legitimately sit in a cold
    // disk read): on expiry nothing is summed, the result is poisoned and the host-visible
    // status word makes the calling entry point fail with RGC_ERR_NCCL.
    if (threadIdx.x < P.nranks) {
      unsigned long long* f = P.peer[threadIdx.x] + flag_off + (std::size_t)P.rank * 16;
      asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(P.seq) : "memory");
      const unsigned long long* mine = P.peer[P.rank] + flag_off + (std::size_t)threadIdx.x * 16;
      unsigned long long        v    = 0;
      const unsigned long long  t0   = global_timer_ns();
      for (unsigned spin = 0;; ++spin) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
        if (v >= P.seq) {
          break;
        }
        if ((spin & 255u) == 255u && global_timer_ns() - t0 > P.timeout_ns) {
          atomicMax(&s_late, 1 + (int)threadIdx.x);
          break;
        }
      }
    }
    __syncthreads();
    if (s_late != 0) {
      if (threadIdx.x == 0) {
        *reinterpret_cast<volatile int*>(P.status) = s_late;
        __threadfence_system();
      }
      for (int i = threadIdx.x; i < P.n; i += blockDim.x) {
        P.data[i] = i < P.n_u64 ? ~0ull : 0x7ff8000000000000ull; // never a plausible partial sum
      }
      return;
    }
    // 3. sum the slots in rank order
    const unsigned long long* src = P.peer[P.rank] + set_off;
    for (int i = threadIdx.x; i < P.n; i += blockDim.x) {
      if (i >= P.n_u64) {
        double s = 0.0;
        for (int r = 0; r < P.nranks; ++r) {
          s += __longlong_as_double((long long)src[(std::size_t)r * kXchgSlot + i]);
        }
        P.data[i] = (unsigned long long)__double_as_longlong(s);
      } else {
        unsigned long long s = 0;
        for (int r = 0; r < P.nranks; ++r) {
          s += src[(std::size_t)r * kXchgSlot + i];
        }
        P.data[i] = s;
      }
    }
  }

  static int* g_xchg_status_host = nullptr; // cudaHostAlloc'ed (mapped), written by the kernel on timeout
  static int* g_xchg_status_dev  = nullptr;
  static bool g_xchg_broken      = false;

  static unsigned long long xchg_timeout_ns() {
    if (const char* s = std::getenv("RGC_XCHG_TIMEOUT_MS")) {
      const long long ms = std::atoll(s);
      if (ms > 0) {
        return (unsigned long long)ms * 1000000ull;
      }
    }
    return 600ull * 1000000000ull;
  }

  static int xchg_allreduce(void* dev, std::size_t n_u64, std::size_t n_f64) {
    auto& c = ctx();
    if (g_xchg_broken) {
      return fail(RGC_ERR_NCCL, "the peer-store exchange timed out earlier; destroy and re-create the "
                                "communicator (rgc_comm_destroy / rgc_comm_init)");
    }
    XchgParams P {};
    for (int r = 0; r < c.nranks; ++r) {
      P.peer[r] = static_cast<unsigned long long*>(c.xchg_peer[r]);
    }
    P.rank       = c.rank;
    P.nranks     = c.nranks;
    P.seq        = ++c.xchg_seq;
    P.data       = static_cast<unsigned long long*>(dev);
    P.n          = (int)(n_u64 + n_f64);
    P.n_u64      = (int)n_u64;
    P.timeout_ns = xchg_timeout_ns();
    P.status     = g_xchg_status_dev;
    xchg_allreduce_kernel<<<1, 256, 0, c.stream>>>(P);
    RGC_CUDA(cudaGetLastError());
    count_launch(1);
    return RGC_OK;
  }

  // after the caller's stream synchronisation: did an exchange of this call time out?
  int exchange_check() {
    if (!g_xchg_status_host) {
      return RGC_OK;
    }
    const int late = *reinterpret_cast<volatile int*>(g_xchg_status_host);
    if (late == 0) {
      return RGC_OK;
    }
    *g_xchg_status_host = 0;
    g_xchg_broken       = true;
    return fail(RGC_ERR_NCCL,
                "all-reduce over the peer-store exchange timed out: rank %d never delivered its "
                "partial result (RGC_XCHG_TIMEOUT_MS); the result of this call is invalid",
                late - 1);
  }

  static void xchg_release(void* local) {
    auto& c = ctx();
    for (int r = 0; r < kXchgMaxRanks; ++r) {
      if (c.xchg_peer[r] && c.xchg_peer[r] != local) {
        cudaIpcCloseMemHandle(c.xchg_peer[r]);
      }
      c.xchg_peer[r] = nullptr;
    }
    if (local) {
      cudaFree(local);
    }
    if (g_xchg_status_host) {
      cudaFreeHost(g_xchg_status_host);
      g_xchg_status_host = nullptr;
      g_xchg_status_dev  = nullptr;
    }
    cudaGetLastError();
  }

  // maps every rank's exchange buffer into this process (CUDA IPC; handles travel through
  // one ncclAllGather).  EVERY rank takes part in both collectives below whatever happened
  // locally (RGC_XCHG=0, a failed allocation, a handle that does not open): a local failure
  // travels as ok = 0 in the gathered record / the reduced verdict, and all ranks then agree
  // on NCCL.  Nothing is leaked on any path.
  static int xchg_setup() {
    auto& c = ctx();
    c.xchg_ready  = false;
    g_xchg_broken = false;
    if (c.nranks < 2 || c.nranks > kXchgMaxRanks) { // the same on every rank
      return RGC_OK;
    }
    const char* env = std::getenv("RGC_XCHG"); // "0": always NCCL (may differ between ranks)
    bool        ok  = !(env && env[0] == '0');
    void*       local = nullptr;
    cudaIpcMemHandle_t mine;
    std::memset(&mine, 0, sizeof(mine));
    if (ok) {
      ok = cudaMalloc(&local, kXchgBytes) == cudaSuccess;
      if (!ok) {
        local = nullptr;
      }
    }
    if (ok) {
      ok = cudaMemset(local, 0, kXchgBytes) == cudaSuccess && cudaDeviceSynchronize() == cudaSuccess &&
           cudaIpcGetMemHandle(&mine, local) == cudaSuccess;
    }
    if (ok) {
      ok = cudaHostAlloc(reinterpret_cast<void**>(&g_xchg_status_host), 64, cudaHostAllocMapped) == cudaSuccess;
      if (ok) {
        *g_xchg_status_host = 0;
        ok = cudaHostGetDevicePointer(reinterpret_cast<void**>(&g_xchg_status_dev), g_xchg_status_host, 0) ==
             cudaSuccess;
      } else {
        g_xchg_status_host = nullptr;
      }
    }
    cudaGetLastError();
    // ---- collective 1: all-gather {handle, ok}; the buffer comes from the context's
    // scratch (no rank-local early exit between here and the verdict)
    const std::size_t rec     = sizeof(cudaIpcMemHandle_t) + 8;
    void*             scratch = nullptr;
    int               rc      = ensure_scratch(rec * c.nranks + 64, &scratch);
    if (rc != RGC_OK) {
      xchg_release(local);
      return rc;
    }
    unsigned char* dbuf = static_cast<unsigned char*>(scratch);
    int*           dflag = reinterpret_cast<int*>(dbuf + ((rec * c.nranks + 15) & ~std::size_t(15)));
    std::vector<unsigned char> hbuf(rec * c.nranks, 0);
    std::memcpy(hbuf.data() + rec * c.rank, &mine, sizeof(mine));
    hbuf[rec * c.rank + sizeof(mine)] = ok ? 1 : 0;
    auto bail = [&](int code) {
      xchg_release(local);
      return code;
    };
    if (cudaMemcpy(dbuf + rec * c.rank, hbuf.data() + rec * c.rank, rec, cudaMemcpyHostToDevice) != cudaSuccess) {
      return bail(fail(RGC_ERR_CUDA, "exchange setup: cudaMemcpy failed: %s", cudaGetErrorString(cudaGetLastError())));
    }
    ncclResult_t nr = nccl().AllGather(dbuf + rec * c.rank, dbuf, rec, ncclChar,
                                       static_cast<ncclComm_t>(c.nccl_comm), c.stream);
    if (nr != ncclSuccess) {
      return bail(fail(RGC_ERR_NCCL, "exchange setup: ncclAllGather failed: %s", nccl().GetErrorString(nr)));
    }
    if (cudaStreamSynchronize(c.stream) != cudaSuccess ||
        cudaMemcpy(hbuf.data(), dbuf, rec * c.nranks, cudaMemcpyDeviceToHost) != cudaSuccess) {
      return bail(fail(RGC_ERR_CUDA, "exchange setup: %s", cudaGetErrorString(cudaGetLastError())));
    }
    for (int r = 0; r < c.nranks; ++r) {
      ok = ok && hbuf[rec * r + sizeof(mine)] == 1;
    }
    if (ok) {
      for (int r = 0; r < c.nranks; ++r) {
        if (r == c.rank) {
          c.xchg_peer[r] = local;
          continue;
        }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, hbuf.data() + rec * r, sizeof(h));
        void* ptr = nullptr;
        if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          cudaGetLastError();
          ok = false;
          break;
        }
        c.xchg_peer[r] = ptr;
      }
    }
    // ---- collective 2: every rank must agree before anyone stores into a peer (this is
    // also the barrier behind the memsets above)
    const int mine_bad = ok ? 0 : 1;
    if (cudaMemcpy(dflag, &mine_bad, sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) {
      return bail(fail(RGC_ERR_CUDA, "exchange setup: cudaMemcpy failed: %s", cudaGetErrorString(cudaGetLastError())));
    }
    nr = nccl().AllReduce(dflag, dflag, 1, ncclInt32, ncclSum, static_cast<ncclComm_t>(c.nccl_comm), c.stream);
    if (nr != ncclSuccess) {
      return bail(fail(RGC_ERR_NCCL, "exchange setup: ncclAllReduce failed: %s", nccl().GetErrorString(nr)));
    }
    int failed = 1;
    if (cudaStreamSynchronize(c.stream) != cudaSuccess ||
        cudaMemcpy(&failed, dflag, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) {
      return bail(fail(RGC_ERR_CUDA, "exchange setup: %s", cudaGetErrorString(cudaGetLastError())));
    }
    if (failed != 0) {
      xchg_release(local); // NCCL carries the exchange on every rank
      return RGC_OK;
    }
    c.xchg_seq   = 0;
    c.xchg_ready = true;
    return RGC_OK;
  }

  static void xchg_teardown() {
    auto& c = ctx();
    if (!c.xchg_ready) {
      return;
    }
    cudaStreamSynchronize(c.stream);
    xchg_release(c.xchg_peer[c.rank]);
    c.xchg_ready = false;
  }

  // in-place sum over ranks of [n_u64 unsigned 64-bit | n_f64 doubles], contiguous
  int allreduce_sum_mixed(void* dev, std::size_t n_u64, std::size_t n_f64) {
    auto& c = ctx();
    if (!c.nccl_comm || c.nranks <= 1 || n_u64 + n_f64 == 0) {
      return RGC_OK;
    }
    if (c.xchg_ready && n_u64 + n_f64 <= kXchgSlot) {
      return xchg_allreduce(dev, n_u64, n_f64);
    }
    auto* base = static_cast<unsigned long long*>(dev);
    RGC_NCCL(nccl().GroupStart());
    if (n_u64) {
      RGC_NCCL(nccl().AllReduce(base, base, n_u64, ncclUint64, ncclSum, static_cast<ncclComm_t>(c.nccl_comm),
                                c.stream));
    }
    if (n_f64) {
      RGC_NCCL(nccl().AllReduce(base + n_u64, base + n_u64, n_f64, ncclFloat64, ncclSum,
                                static_cast<ncclComm_t>(c.nccl_comm), c.stream));
    }
    RGC_NCCL(nccl().GroupEnd());
    return RGC_OK;
  }

  int allreduce_sum_f64(double* dev, std::size_t n) { return allreduce_sum_mixed(dev, 0, n); }

  int allreduce_sum_u64(unsigned long long* dev, std::size_t n) { return allreduce_sum_mixed(dev, n, 0); }

} // namespace rgc

using namespace rgc;

static_assert(sizeof(ncclUniqueId) == RGC_COMM_ID_BYTES, "ncclUniqueId size");

extern "C" {

  // ------------------------------------------------------------------ runtime
  const char* rgc_last_error(void) { return t_last_error.c_str(); }

  int rgc_device_count(int* count) {
    int         n   = 0;
    cudaError_t err = cudaGetDeviceCount(&n);
    if (err != cudaSuccess) {
      cudaGetLastError();
      n = 0;
    }
    if (count) {
      *count = n;
    }
    return RGC_OK;
  }

  int rgc_init(int device) {
    auto& c = ctx();
    if (c.initialized) {
      return RGC_OK;
    }
    if (std::getenv("RGC_DEBUG_SIGNALS")) { // native backtrace on SIGFPE / SIGSEGV (debug aid)
      auto handler = +[](int sig) {
        void* frames[64];
        const int nf = backtrace(frames, 64);
        backtrace_symbols_fd(frames, nf, 2);
        signal(sig, SIG_DFL);
        raise(sig);
      };
      signal(SIGFPE, handler);
      signal(SIGSEGV, handler);
    }
    int ndev = 0;
    rgc_device_count(&ndev);
    if (ndev == 0) {
      return fail(RGC_ERR_NOT_INITIALIZED,
                  "no CUDA device visible: ragnar_cuda has no CPU fallback");
    }
    if (device < 0) {
      const char* lr = std::getenv("LOCAL_RANK");
      device         = lr ? std::atoi(lr) % ndev : 0;
    }
    if (device >= ndev) {
      return fail(RGC_ERR_INVALID, "device %d out of range (%d visible)", device, ndev);
    }
    RGC_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    RGC_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
      return fail(RGC_ERR_NOT_INITIALIZED,
                  "device %d (%s, sm_%d%d) is not a Blackwell sm_100 part; this library "
                  "ships sm_100a code only",
                  device, prop.name, prop.major, prop.minor);
    }
    c.device    = device;
    c.sm_count  = prop.multiProcessorCount;
    c.hbm_bytes = prop.totalGlobalMem;
    RGC_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    RGC_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
    for (auto& e : c.ev) {
      RGC_CUDA(cudaEventCreate(&e));
    }
    {
      cudaMemPoolProps props {};
      props.allocType     = cudaMemAllocationTypePinned;
      props.handleTypes   = cudaMemHandleTypeNone;
      props.location.type = cudaMemLocationTypeDevice;
      props.location.id   = device;
      RGC_CUDA(cudaMemPoolCreate(&c.column_pool, &props));
      std::uint64_t keep = ~std::uint64_t(0); // freed columns stay cached for the next container
      RGC_CUDA(cudaMemPoolSetAttribute(c.column_pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    c.launches    = 0;
    c.initialized = true;
    return RGC_OK;
  }

  int rgc_trim_memory(void) {
    RGC_REQUIRE_INIT();
    auto& c = ctx();
    RGC_CUDA(cudaStreamSynchronize(c.stream));
    RGC_CUDA(cudaMemPoolTrimTo(c.column_pool, 0));
    return RGC_OK;
  }

  int rgc_finalize(void) {
    auto& c = ctx();
    if (!c.initialized) {
      return RGC_OK;
    }
    cudaSetDevice(c.device);
    cudaDeviceSynchronize();
    rgc_comm_destroy();
    io_release_lanes();
    pair_release_plans();
    copy_pool_release();
    for (int s = 0; s < kNumStages; ++s) {
      if (c.stage[s]) {
        cudaFreeHost(c.stage[s]);
        cudaEventDestroy(c.stage_free[s]);
        c.stage[s]      = nullptr;
        c.stage_free[s] = nullptr;
      }
    }
    if (c.scratch) {
      cudaFree(c.scratch);
      c.scratch       = nullptr;
      c.scratch_bytes = 0;
    }
    if (c.result) {
      cudaFree(c.result);
      c.result       = nullptr;
      c.result_bytes = 0;
    }
    for (auto& e : c.ev) {
      if (e) {
        cudaEventDestroy(e);
        e = nullptr;
      }
    }
    if (c.column_pool) { // containers still alive keep their (now orphaned) columns until process exit
      cudaMemPoolDestroy(c.column_pool);
      c.column_pool = nullptr;
    }
    cudaStreamDestroy(c.stream);
    cudaStreamDestroy(c.copy_stream);
    c.stream      = nullptr;
    c.copy_stream = nullptr;
    c.initialized = false;
    return RGC_OK;
  }

  int rgc_is_initialized(void) { return ctx().initialized ? 1 : 0; }

  int rgc_device_info(int* device, int* sm_count, size_t* hbm_bytes) {
    RGC_REQUIRE_INIT();
    if (device) {
      *device = ctx().device;
    }
    if (sm_count) {
      *sm_count = ctx().sm_count;
    }
    if (hbm_bytes) {
      *hbm_bytes = ctx().hbm_bytes;
    }
    return RGC_OK;
  }

  int rgc_stream(void** stream) {
    RGC_REQUIRE_INIT();
    *stream = static_cast<void*>(ctx().stream);
    return RGC_OK;
  }

  int rgc_synchronize(void) {
    RGC_REQUIRE_INIT();
    RGC_CUDA(cudaStreamSynchronize(ctx().copy_stream));
    RGC_CUDA(cudaStreamSynchronize(ctx().stream));
    return RGC_OK;
  }

  uint64_t rgc_launch_count(void) { return ctx().launches.load(); }

  int rgc_last_kernel_ms(float ms[2]) {
    ms[0] = ctx().last_ms[0];
    ms[1] = ctx().last_ms[1];
    return RGC_OK;
  }

  int rgc_last_kernel_times(float* ms, int n) {
    for (int i = 0; i < n; ++i) {
      ms[i] = i < 4 ? ctx().last_ms[i] : 0.f;
    }
    return RGC_OK;
  }

  int rgc_host_alloc(size_t bytes, void** ptr) {
    RGC_REQUIRE_INIT();
    cudaError_t err = cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault);
    if (err != cudaSuccess) {
      cudaGetLastError();
      return fail(RGC_ERR_OOM, "cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(err));
    }
    return RGC_OK;
  }

  int rgc_host_free(void* ptr) {
    if (ptr) {
      RGC_CUDA(cudaFreeHost(ptr));
    }
    return RGC_OK;
  }

  // --------------------------------------------------------------------- comm
  int rgc_comm_get_unique_id(unsigned char id[RGC_COMM_ID_BYTES]) {
    RGC_TRY(load_nccl());
    ncclUniqueId uid;
    RGC_NCCL(nccl().GetUniqueId(&uid));
    std::memcpy(id, &uid, RGC_COMM_ID_BYTES);
    return RGC_OK;
  }

  int rgc_comm_init(const unsigned char id[RGC_COMM_ID_BYTES], int rank, int nranks) {
    RGC_REQUIRE_INIT();
    if (nranks < 1 || rank < 0 || rank >= nranks) {
      return fail(RGC_ERR_INVALID, "bad rank %d / nranks %d", rank, nranks);
    }
    RGC_TRY(load_nccl());
    auto& c = ctx();
    if (c.nccl_comm) {
      return fail(RGC_ERR_INVALID, "communicator already initialised");
    }
    ncclUniqueId uid;
    std::memcpy(&uid, id, RGC_COMM_ID_BYTES);
    ncclComm_t comm = nullptr;
    RGC_CUDA(cudaSetDevice(c.device));
    RGC_NCCL(nccl().CommInitRank(&comm, nranks, uid, rank));
    c.nccl_comm = comm;
    c.rank      = rank;
    c.nranks    = nranks;
    RGC_TRY(xchg_setup());
    return RGC_OK;
  }

  int rgc_comm_destroy(void) {
    auto& c = ctx();
    xchg_teardown();
    if (c.nccl_comm) {
      nccl().CommDestroy(static_cast<ncclComm_t>(c.nccl_comm));
      c.nccl_comm = nullptr;
    }
    c.rank   = 0;
    c.nranks = 1;
    return RGC_OK;
  }

  int rgc_comm_exchange_kind(int* kind) {
    if (kind) {
      auto& c = ctx();
      *kind   = (c.nccl_comm && c.nranks > 1) ? (c.xchg_ready ? 2 : 1) : 0;
    }
    return RGC_OK;
  }

  int rgc_comm_info(int* rank, int* nranks) {
    if (rank) {
      *rank = ctx().rank;
    }
    if (nranks) {
      *nranks = ctx().nranks;
    }
    return RGC_OK;
  }

  // ------------------------------------------------------------------ buffers
  int rgc_buf_create(int dtype, size_t n, rgc_buf_t** out) {
    RGC_REQUIRE_INIT();
    if (dtype != RGC_I32 && dtype != RGC_F32 && dtype != RGC_F64) {
      return fail(RGC_ERR_INVALID, "unknown dtype %d", dtype);
    }
    auto* b = new (std::nothrow) rgc_buf;
    if (!b) {
      return fail(RGC_ERR_OOM, "out of host memory");
    }
    b->dtype = dtype;
    b->n     = n;
    if (n > 0) {
      const std::size_t bytes = n * dtype_size(dtype);
      cudaError_t       err   = cudaMalloc(&b->dev, bytes);
      if (err != cudaSuccess) {
        cudaGetLastError();
        delete b;
        return fail(RGC_ERR_OOM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(err));
      }
      err = cudaMemsetAsync(b->dev, 0, bytes, ctx().stream);
      if (err != cudaSuccess) {
        cudaFree(b->dev);
        delete b;
        return fail(RGC_ERR_CUDA, "cudaMemsetAsync failed: %s", cudaGetErrorString(err));
      }
    }
    *out = b;
    return RGC_OK;
  }

  int rgc_buf_from_host(int dtype, const void* host, size_t n, rgc_buf_t** out) {
    rgc_buf_t* b = nullptr;
    RGC_TRY(rgc_buf_create(dtype, n, &b));
    if (n > 0) {
      int rc = copy_h2d(b->dev, host, n * dtype_size(dtype), ctx().stream);
      if (rc == RGC_OK) {
        cudaError_t err = cudaStreamSynchronize(ctx().stream);
        if (err != cudaSuccess) {
          rc = fail(RGC_ERR_CUDA, "cudaStreamSynchronize failed: %s", cudaGetErrorString(err));
        }
      }
      if (rc != RGC_OK) {
        rgc_buf_release(b);
        return rc;
      }
    }
    *out = b;
    return RGC_OK;
  }

  int rgc_buf_to_host(const rgc_buf_t* buf, size_t start, size_t n, void* host) {
    RGC_REQUIRE_INIT();
    if (!buf || start + n > buf->n) {
      return fail(RGC_ERR_INVALID, "rgc_buf_to_host: range [%zu, %zu) exceeds extent %zu",
                  start, start + n, buf ? buf->n : (size_t)0);
    }
    if (n == 0) {
      return RGC_OK;
    }
    const std::size_t es = dtype_size(buf->dtype);
    RGC_CUDA(cudaMemcpyAsync(host, static_cast<const char*>(buf->dev) + start * es, n * es,
                             cudaMemcpyDeviceToHost, ctx().stream));
    RGC_CUDA(cudaStreamSynchronize(ctx().stream));
    return RGC_OK;
  }

  size_t rgc_buf_size(const rgc_buf_t* buf) { return buf ? buf->n : 0; }
  int    rgc_buf_dtype(const rgc_buf_t* buf) { return buf ? buf->dtype : -1; }
  void*  rgc_buf_device_ptr(const rgc_buf_t* buf) { return buf ? buf->dev : nullptr; }

  int rgc_buf_retain(rgc_buf_t* buf) {
    if (buf) {
      buf->refcount.fetch_add(1);
    }
    return RGC_OK;
  }

  int rgc_buf_release(rgc_buf_t* buf) {
    if (buf && buf->refcount.fetch_sub(1) == 1) {
      if (buf->dev && ctx().initialized) {
        cudaStreamSynchronize(ctx().stream);
        cudaFree(buf->dev);
      }
      delete buf;
    }
    return RGC_OK;
  }

  // ---------------------------------------------------------------- particles
  int rgc_particles_create(int dim, rgc_particles_t** out) {
    if (dim < 1 || dim > 3) {
      return fail(RGC_ERR_INVALID, "dim must be 1, 2 or 3 (got %d)", dim);
    }
    auto* p = new (std::nothrow) rgc_particles;
    if (!p) {
      return fail(RGC_ERR_OOM, "out of host memory");
    }
    p->dim = dim;
    *out   = p;
    return RGC_OK;
  }

  // Particle columns come from the library's own stream-ordered memory pool (release threshold
  // unlimited, created in rgc_init): a container that is dropped and re-created — what a script
  // does per species and step — gets its memory back without a device-wide cudaFree /
  // cudaMalloc round trip (tens of ms for GB-sized columns).  rgc_trim_memory() returns the
  // cached blocks to the device.
  static void free_columns(rgc_particles_t* p) {
    for (auto& q : p->col) {
      for (auto& c : q) {
        if (c) {
          cudaFreeAsync(c, ctx().stream);
          c = nullptr;
        }
      }
    }
  }

  int rgc_particles_release(rgc_particles_t* p) {
    if (p) {
      if (ctx().initialized) {
        cudaStreamSynchronize(ctx().copy_stream);
        free_columns(p); // ordered behind everything already enqueued on the compute stream
      }
      delete p;
    }
    return RGC_OK;
  }

  static std::size_t pitch_for(std::size_t nalloc) {
    const std::size_t n = nalloc < 64 ? 64 : nalloc;
    return (n + 63) & ~std::size_t(63);
  }

  static int alloc_column(float** col, std::size_t pitch) {
    cudaError_t err = cudaMallocFromPoolAsync(reinterpret_cast<void**>(col), pitch * sizeof(float),
                                              ctx().column_pool, ctx().stream);
    if (err != cudaSuccess) {
      cudaGetLastError();
      *col = nullptr;
      return fail(RGC_ERR_OOM, "cudaMallocAsync of a %zu-particle column failed: %s", pitch,
                  cudaGetErrorString(err));
    }
    RGC_CUDA(cudaMemsetAsync(*col, 0, pitch * sizeof(float), ctx().stream));
    return RGC_OK;
  }

  int rgc_particles_allocate(rgc_particles_t* p, size_t nalloc, int with_coords) {
    RGC_REQUIRE_INIT();
    if (!p) {
      return fail(RGC_ERR_INVALID, "null particles handle");
    }
    if (p->allocated) {
      return fail(RGC_ERR_INVALID, "Particles already allocated");
    }
    p->pitch = pitch_for(nalloc);
    for (int q = with_coords ? RGC_Q_X : RGC_Q_U; q <= RGC_Q_B; ++q) {
      const int ncomp = q == RGC_Q_X ? p->dim : 3;
      for (int c = 0; c < ncomp; ++c) {
        int rc = alloc_column(&p->col[q][c], p->pitch);
        if (rc != RGC_OK) {
          free_columns(p);
          return rc;
        }
      }
    }
    p->nalloc      = nalloc;
    p->with_coords = with_coords != 0;
    p->allocated   = true;
    return RGC_OK;
  }

  int rgc_particles_enable_coords(rgc_particles_t* p) {
    RGC_REQUIRE_INIT();
    if (!p || !p->allocated) {
      return fail(RGC_ERR_INVALID, "Particles not allocated");
    }
    if (p->with_coords) {
      return RGC_OK;
    }
    for (int c = 0; c < p->dim; ++c) {
      RGC_TRY(alloc_column(&p->col[RGC_Q_X][c], p->pitch));
    }
    p->with_coords = true;
    return RGC_OK;
  }

  int rgc_particles_reallocate(rgc_particles_t* p, size_t nalloc) {
    RGC_REQUIRE_INIT();
    if (!p || !p->allocated) {
      return fail(RGC_ERR_INVALID,
                  "Particles not allocated, if you want to allocate, call `allocate` instead");
    }
    if (nalloc <= p->nalloc) {
      return fail(RGC_ERR_INVALID, "New allocation size must be greater than the current one");
    }
    const std::size_t new_pitch = pitch_for(nalloc);
    if (new_pitch > p->pitch) {
      RGC_CUDA(cudaStreamSynchronize(ctx().copy_stream));
      for (auto& q : p->col) {
        for (auto& c : q) {
          if (!c) {
            continue;
          }
          float* bigger = nullptr;
          RGC_TRY(alloc_column(&bigger, new_pitch));
          RGC_CUDA(cudaMemcpyAsync(bigger, c, p->nalloc * sizeof(float),
                                   cudaMemcpyDeviceToDevice, ctx().stream));
          RGC_CUDA(cudaFreeAsync(c, ctx().stream));
          c = bigger;
        }
      }
      p->pitch = new_pitch;
    }
    p->nalloc = nalloc;
    return RGC_OK;
  }

  int    rgc_particles_has_coords(const rgc_particles_t* p) { return p && p->with_coords; }
  size_t rgc_particles_nalloc(const rgc_particles_t* p) { return p ? p->nalloc : 0; }
  int    rgc_particles_dim(const rgc_particles_t* p) { return p ? p->dim : 0; }

  static int check_column(const rgc_particles_t* p, int quantity, int comp) {
    if (!p || !p->allocated) {
      return fail(RGC_ERR_INVALID, "Particles not allocated");
    }
    if (quantity < RGC_Q_X || quantity > RGC_Q_B) {
      return fail(RGC_ERR_INVALID, "Invalid quantity");
    }
    const int ncomp = quantity == RGC_Q_X ? p->dim : 3;
    if (comp < 0 || comp >= ncomp) {
      return fail(RGC_ERR_INVALID, "Invalid component");
    }
    if (quantity == RGC_Q_X && !p->with_coords) {
      return fail(RGC_ERR_INVALID, "Particle coordinates ignored");
    }
    return RGC_OK;
  }

  int rgc_particles_write(rgc_particles_t* p, int quantity, int comp, size_t start,
                          const float* host, size_t n) {
    RGC_REQUIRE_INIT();
    RGC_TRY(check_column(p, quantity, comp));
    if (start + n > p->nalloc) {
      return fail(RGC_ERR_INVALID, "write [%zu, %zu) exceeds allocation %zu", start, start + n,
                  p->nalloc);
    }
    // ordered after the zero-fill issued on the compute stream
    return copy_h2d(p->col[quantity][comp] + start, host, n * sizeof(float), ctx().stream);
  }

  int rgc_particles_read(const rgc_particles_t* p, int quantity, int comp, size_t start,
                         size_t n, float* host) {
    RGC_REQUIRE_INIT();
    RGC_TRY(check_column(p, quantity, comp));
    if (start + n > p->nalloc) {
      return fail(RGC_ERR_INVALID, "read [%zu, %zu) exceeds allocation %zu", start, start + n,
                  p->nalloc);
    }
    if (n == 0) {
      return RGC_OK;
    }
    RGC_CUDA(cudaMemcpyAsync(host, p->col[quantity][comp] + start, n * sizeof(float),
                             cudaMemcpyDeviceToHost, ctx().stream));
    RGC_CUDA(cudaStreamSynchronize(ctx().stream));
    return RGC_OK;
  }

  int rgc_particles_column(const rgc_particles_t* p, int quantity, int comp, size_t n,
                           rgc_buf_t** out) {
    RGC_REQUIRE_INIT();
    RGC_TRY(check_column(p, quantity, comp));
    if (n > p->nalloc) {
      return fail(RGC_ERR_INVALID, "column length %zu exceeds allocation %zu", n, p->nalloc);
    }
    rgc_buf_t* b = nullptr;
    RGC_TRY(rgc_buf_create(RGC_F32, n, &b));
    if (n > 0) {
      cudaError_t err = cudaMemcpyAsync(b->dev, p->col[quantity][comp], n * sizeof(float),
                                        cudaMemcpyDeviceToDevice, ctx().stream);
      if (err != cudaSuccess) {
        rgc_buf_release(b);
        return fail(RGC_ERR_CUDA, "cudaMemcpyAsync D2D failed: %s", cudaGetErrorString(err));
      }
    }
    *out = b;
    return RGC_OK;
  }

  void* rgc_particles_device_ptr(const rgc_particles_t* p, int quantity, int comp) {
    if (check_column(p, quantity, comp) != RGC_OK) {
      return nullptr;
    }
    return p->col[quantity][comp];
  }

} // extern "C"
