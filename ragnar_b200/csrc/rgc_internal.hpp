// Internal declarations shared by the translation units of libragnar_cuda.so.
// Not part of the C-ABI (include/ragnar_cuda.h is).
#ifndef RGC_INTERNAL_HPP
#define RGC_INTERNAL_HPP

#include "ragnar_cuda.h"

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h> // header-only; a no-op unless a profiler is attached

#include <atomic>
#include <cstdarg>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace rgc {

  // ------------------------------------------------------------------ errors
  int  fail(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
  void clear_error();

#define RGC_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t rgc_err__ = (expr);                                                 \
    if (rgc_err__ != cudaSuccess) {                                                 \
      return ::rgc::fail(rgc_err__ == cudaErrorMemoryAllocation ? RGC_ERR_OOM       \
                                                                : RGC_ERR_CUDA,     \
                         "%s failed: %s (%s:%d)", #expr,                            \
                         cudaGetErrorString(rgc_err__), __FILE__, __LINE__);        \
    }                                                                               \
  } while (0)

#define RGC_TRY(expr)               \
  do {                              \
    int rgc_rc__ = (expr);          \
    if (rgc_rc__ != RGC_OK) {       \
      return rgc_rc__;              \
    }                               \
  } while (0)

#define RGC_REQUIRE_INIT()                                                          \
  do {                                                                              \
    if (!::rgc::ctx().initialized) {                                                \
      return ::rgc::fail(RGC_ERR_NOT_INITIALIZED,                                   \
                         "ragnar_cuda is not initialized (call rgc_init; a CUDA "   \
                         "device is required, there is no CPU fallback)");          \
    }                                                                               \
  } while (0)

  // NVTX range over an entry point, named after the Kokkos kernel label it replaces
  // (reference: "SynchrotronSpectrum" src/physics/synchrotron.cpp:92,131,
  // "ComputeEnergyDistribution" src/containers/particles.cpp:225, "ICSpectrum"
  // src/physics/ic.cpp:38, "Linspace"/"Logspace" src/utils/snippets.cpp:27,49, "XMinMax"
  // src/containers/tabulation.cpp:89), so a timeline of this library reads like one of
  // the reference's Kokkos-tools traces
  struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&)            = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
  };
#define RGC_NVTX(name) ::rgc::NvtxRange rgc_nvtx_range__ { name }

  // ----------------------------------------------------------------- context
  constexpr std::size_t kStageBytes = std::size_t(4) << 20; // per pinned stage of the small-copy ring
  constexpr int         kNumStages  = 4;

  struct Context {
    bool         initialized { false };
    int          device { 0 };
    int          sm_count { 0 };
    std::size_t  hbm_bytes { 0 };
    cudaStream_t stream { nullptr };      // compute + ordered copies
    cudaStream_t copy_stream { nullptr }; // bulk H2D of particle columns
    cudaMemPool_t column_pool { nullptr }; // stream-ordered pool of the particle columns (own pool:
                                           // its unlimited release threshold touches nobody else)
    // pinned staging ring for pageable host sources
    void*       stage[kNumStages] {};
    cudaEvent_t stage_free[kNumStages] {};
    // event pair for rgc_last_kernel_ms
    cudaEvent_t ev[6] { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
    float       last_ms[4] { 0.f, 0.f, 0.f, 0.f }; // total, dominant kernel, prologue kernel, sort
    double      last_lane_evals { 0.0 }; // hinge evaluations the pair kernel issued in the last call
    double      last_ontable_evals { 0.0 }; // pairs of that call on a non-zero cell pair of the table
    // NCCL (dlopen'ed lazily)
    void* nccl_comm { nullptr };
    int   rank { 0 };
    int   nranks { 1 };
    // peer-store exchange over NVLink (rgc_runtime.cu: xchg_*): every rank's buffer is
    // mapped into every other rank through CUDA IPC; small result vectors are all-reduced
    // by one kernel (stores into all peers, a flag, sum in rank order) instead of NCCL
    bool               xchg_ready { false };
    void*              xchg_peer[8] { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
    unsigned long long xchg_seq { 0 };
    // reusable device scratch (grown on demand, freed in rgc_finalize)
    void*       scratch { nullptr };
    std::size_t scratch_bytes { 0 };
    // small device buffer for per-call result vectors (survives scratch re-layouts)
    void*       result { nullptr };
    std::size_t result_bytes { 0 };
    std::atomic<std::uint64_t> launches { 0 };
  };

  Context& ctx();
  int      ensure_scratch(std::size_t bytes, void** out);
  int      ensure_result(std::size_t bytes, void** out);
  inline void count_launch(int n = 1) { ctx().launches.fetch_add((std::uint64_t)n); }

  // all-reduce (sum) in place on the compute stream; no-op without a communicator
  int allreduce_sum_f64(double* dev, std::size_t n);
  int allreduce_sum_u64(unsigned long long* dev, std::size_t n);
  // [n_u64 unsigned 64-bit | n_f64 doubles], contiguous: ONE exchange kernel (or one NCCL group)
  int allreduce_sum_mixed(void* dev, std::size_t n_u64, std::size_t n_f64);
  // after the stream synchronisation that follows an all-reduce: RGC_ERR_NCCL when a peer
  // never delivered its partial result within RGC_XCHG_TIMEOUT_MS (the result is poisoned)
  int exchange_check();

  // frees the pinned I/O lanes of the HDF5 streaming reader (rgc_tristan.cpp)
  void io_release_lanes();

  // host -> device copy that accepts pageable or pinned sources (see rgc_runtime.cu)
  int copy_h2d(void* dst, const void* src, std::size_t bytes, cudaStream_t stream);

  // ------------------------------------------------------------ host math
  // (rgc_hostmath.cpp, compiled by g++ with -ffp-contract=off: these reproduce
  // the reference's float/double promotions on the host)
  void  host_linspace(float start, float stop, std::size_t num, float* out);
  void  host_logspace(float start, float stop, std::size_t num, float* out);
  float host_ffunc_integrand(float x);
  void  host_tabulate_ffunc(std::size_t n, float xmin, float xmax, float* xs, float* ys);
  float host_interpolate(bool loggrid, float x0, const float* x, const float* y,
                         std::size_t n, float yfill);
  int   host_generator_eval(int kind, const float* params, const float* energy,
                            std::size_t n, float* out);
  // reference bin index of the energy histogram as a function of Usqr (float)
  std::size_t host_energy_bin_index(float Usqr, bool fourvel, float energy_min,
                                    float energy_max, std::size_t n);
  double      host_energy_from_usqr(float Usqr, bool fourvel);

  // ------------------------------------------------- synchrotron table plan
  // The F(x) table as the kernels see it: node positions in cell units of the
  // uniform log grid the reference indexes with (tabulation.hpp:33-35).
  struct TablePlan {
    double              L0 { 0 }, dL { 0 };
    std::vector<double> tx; // actual node positions in cell units
    std::vector<double> y;
    std::size_t         T { 0 };
  };

} // namespace rgc

struct rgc_particles;

namespace rgc {
  // bucketed hinge path (rgc_sync_pair.cu)
  bool pair_path_eligible(const TablePlan& tp, const float* bins_e_syn,
                          const std::vector<int>& bins);
  // adds the chunk's per-bin sums into d_acc[bins[s]] on the device (stream-ordered)
  int  run_spectrum_pair(const rgc_particles* prtls, std::size_t n, float B0, float g_syn,
                         float e_at, const TablePlan& tp, const float* bins_e_syn,
                         const std::vector<int>& bins, double* d_acc, int* d_poison,
                         float* main_ms, bool defer_sync);
  // after the caller's own stream synchronisation: kernel times of a deferred pass
  int  collect_pair_times(float* main_ms);
  void pair_plan_describe(const TablePlan& tp, const float* bins_e_syn, const std::vector<int>& bins,
                          int info[8], float* phase, int* slot_bin, std::size_t cap);
  bool pair_single_pass(std::size_t n); // n particles fit one pipeline pass
  void pair_release_plans();            // frees the cached device plans (rgc_finalize)
  int  pair_rank_mode(); // verdict of the rank-order probe: -1 not run, 0 ballots, 1 atomics
  // d_acc[binmap[s]] += src[s] for every slot with binmap[s] >= 0
  int  launch_scatter_add(const double* src, const int* binmap, int nslots, double* d_acc);
  // literal evaluation (rgc_sync_literal.cu): the reference's float term per pair, summed
  // in fp64; sources = the first n particles of `prtls`, or n host triples when src_ep != 0
  // (then src_w1 holds nbatch rows of n weights and d_out nbatch rows of nbins sums;
  // contract = true builds the kernel matrix once and contracts the batch in fp64)
  std::size_t literal_max_n(); // RGC_LITERAL_MAX_N, default 2^19
  int  run_spectrum_literal(const rgc_particles* prtls, std::size_t n, float B0, float g_syn,
                            float e_at, const float* src_ep, const float* src_w1,
                            const float* src_w2, const float* bins_e_syn, std::size_t nbins,
                            const float* tab_x, const float* tab_y, std::size_t T, double* d_out,
                            std::size_t nbatch = 1, bool contract = false);
} // namespace rgc

// ------------------------------------------------------------ opaque handles
struct rgc_buf {
  std::atomic<int> refcount { 1 };
  int              dtype { RGC_F32 };
  std::size_t      n { 0 };
  void*            dev { nullptr };
};

struct rgc_particles {
  int         dim { 3 };
  std::size_t nalloc { 0 };
  std::size_t pitch { 0 }; // floats per column (nalloc rounded up to 64)
  bool        allocated { false };
  bool        with_coords { false };
  float*      col[4][3] { { nullptr, nullptr, nullptr },
                          { nullptr, nullptr, nullptr },
                          { nullptr, nullptr, nullptr },
                          { nullptr, nullptr, nullptr } };
};

namespace rgc {
  inline std::size_t dtype_size(int dtype) { return dtype == RGC_F64 ? 8 : 4; }
} // namespace rgc

#endif // RGC_INTERNAL_HPP
