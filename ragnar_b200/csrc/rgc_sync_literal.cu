// Literal evaluation of the synchrotron pair term for sm_100a: every (source, photon
// bin) term is formed with the reference's own float arithmetic, bit for bit, and the
// terms are summed in fp64 in a fixed order.
//
// Replaces (reference paths relative to haykh/ragnar @ fceb6b08):
//   sync::Kernel<D>::operator() / OmegaSync_ChiR   src/physics/synchrotron.hpp:145-232
//   sync::KernelFromDist::operator()               src/physics/synchrotron.hpp:72-95
//   InterpolateTabulatedFunction<true>             src/containers/tabulation.hpp:19-42
//
// Why it exists: the bucketed hinge pipeline (rgc_sync_pair.cu) evaluates the exact
// interpolant, while the reference evaluates it through ~5 log10f and 5 float divisions
// per pair; the two agree to ~1e-6 once a few thousand particles with different energies
// average the reference's rounding, but not for a handful of particles, for identical
// particles (the rounding is then the same in every term) or for the 200-term sums of
// SynchrotronSpectrumFromDist.  Those calls cost nothing, so they are evaluated the
// reference's way: x0 = e_syn / e_peak, the index from log10f(x0 / xmin), the log-log
// blend of the two nodes — all IEEE float operations (-fmad=false, -prec-div=true) plus
// glibc's log10f (rgc_glibc_log10f.cuh, bit-identical to the image's libm over all 2^32
// arguments).  Per-term parity with the reference is exact; the sums differ from
// `ragnar_ref64` (the reference's float terms summed in double) only by fp64
// summation order (~1e-16 n).
//
// Used for SynchrotronSpectrum_<D>D when nactive <= RGC_LITERAL_MAX_N (default 2^19;
// 0 disables) and always for SynchrotronSpectrumFromDist.
//
// Layout: a thread owns one photon bin, a CTA 128 consecutive bins; sources are cut into
// slices (gridDim.y) that a CTA walks in chunks of 128 staged in shared memory (the
// per-source prologue runs once per CTA and chunk); per-slice partial sums are added in
// slice order by reduce_partials_kernel: deterministic.  ~250 instructions per pair
// against 2 in the hinge kernel — by design only for calls of <= ~1e9 pairs.
#include "rgc_glibc_log10f.cuh"
#include "rgc_internal.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

namespace rgc {

  constexpr int kLitThreads = 128;

  struct LiteralParams {
    // sources: particle columns (src_ep == nullptr) or per-source arrays (FromDist form)
    const float* u[3];
    const float* e[3];
    const float* b[3];
    const float* src_ep; // e_peak of every source
    const float* src_w1; // term = ((w1 * e_syn) [* w2]) * F; [nbatch][nsrc]: one row per batch item
    const float* src_w2; // nullptr: no second factor
    std::size_t  nsrc;
    int          nbatch; // blockIdx.z: distributions sharing (e_peak, w2) and the photon bins
    std::size_t  src_per_slice;
    float        B0, g_syn, e_at;
    const float* bins;
    int          nbins;
    const float* tab_x;
    const float* tab_y;
    const float* tab_den; // [T - 1] log10f(x[k + 1] / x[k])
    int          T;
    float        xmin, xmax, Lspan; // Lspan = log10f(xmax / xmin)
    double*      partials;          // [nbatch][slices][nbins]
  };

  __device__ const LogfEntry g_logf_tab[16] = RGC_LOGF_TAB_INIT;

  // den[k] = log10f(x[k + 1] / x[k])  (tabulation.hpp:41), once per call
  __global__ void literal_den_kernel(const float* __restrict__ x, int T, float* __restrict__ den) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k + 1 < T) {
      den[k] = glibc_log10f(x[k + 1] / x[k], g_logf_tab);
    }
  }

  // reference src/physics/synchrotron.hpp:193-231 with its double promotions; the dead
  // terms (beta_Sqr, eperp_*, eprime_*) never reach the outputs
  __device__ __forceinline__ void literal_prologue(const LiteralParams& P, std::size_t i,
                                                   float& e_peak, float& chiR) {
    const float  ux = P.u[0][i], uy = P.u[1][i], uz = P.u[2][i];
    const float  ex = P.e[0][i], ey = P.e[1][i], ez = P.e[2][i];
    const float  bx = P.b[0][i], by = P.b[1][i], bz = P.b[2][i];
    const double gamma  = sqrt(((1.0 + (double)(ux * ux)) + (double)(uy * uy)) + (double)(uz * uz));
    const double beta_x = (double)ux / gamma;
    const double beta_y = (double)uy / gamma;
    const double beta_z = (double)uz / gamma;
    const double bde    = (beta_x * ex + beta_y * ey) + beta_z * ez;
    const double cx     = beta_y * bz - beta_z * by;
    const double cy     = beta_z * bx - beta_x * bz;
    const double cz     = beta_x * by - beta_y * bx;
    const double sx = ex + cx, sy = ey + cy, sz = ez + cz;
    const double ssq = (sx * sx + sy * sy) + sz * sz;
    chiR   = (float)(sqrt(ssq - bde * bde) / (double)P.B0);
    e_peak = (float)((((double)P.e_at * gamma) * gamma) * (double)chiR / (double)(P.g_syn * P.g_syn));
  }

  // InterpolateTabulatedFunction<true> (tabulation.hpp:19-42), all float
  __device__ __forceinline__ float literal_interp(const LiteralParams& P, const LogfEntry* lt,
                                                  float x0) {
    if (x0 < P.xmin || x0 >= P.xmax) {
      return 0.0f; // yfill
    }
    const float v = ((float)(P.T - 1) * fabsf(glibc_log10f(x0 / P.xmin, lt))) / P.Lspan;
    // static_cast<std::size_t>(v): NaN (x0 = NaN passes both comparisons) and values
    // beyond 2^63 convert to an index >= n - 1 on x86-64
    const bool big = !(v < 9.2e18f);
    const unsigned long long xi = big ? ~0ull : (unsigned long long)v;
    if (big || xi >= (unsigned long long)(P.T - 1)) {
      return __ldg(P.tab_y + (P.T - 1));
    }
    const float xk = __ldg(P.tab_x + xi), xk1 = __ldg(P.tab_x + xi + 1);
    const float yk = __ldg(P.tab_y + xi), yk1 = __ldg(P.tab_y + xi + 1);
    const float la = glibc_log10f(x0 / xk, lt);
    const float lb = glibc_log10f(xk1 / x0, lt);
    return (yk1 * la + yk * lb) / __ldg(P.tab_den + xi);
  }

  __global__ void __launch_bounds__(kLitThreads)
    sync_literal_kernel(const __grid_constant__ LiteralParams P) {
    __shared__ LogfEntry lt[16];
    __shared__ float     s_ep[kLitThreads], s_w1[kLitThreads], s_w2[kLitThreads];
    const int tid = threadIdx.x;
    if (tid < 16) {
      lt[tid] = g_logf_tab[tid];
    }
    const int   j     = blockIdx.x * kLitThreads + tid;
    const float e_syn = j < P.nbins ? P.bins[j] : 0.0f;
    const bool  has_w2 = P.src_w2 != nullptr;
    double      acc   = 0.0;
    const std::size_t s0 = (std::size_t)blockIdx.y * P.src_per_slice;
    const std::size_t s1 = min(s0 + P.src_per_slice, P.nsrc);
    for (std::size_t base = s0; base < s1; base += kLitThreads) {
      __syncthreads();
      {
        const std::size_t i = base + tid;
        float ep = 0.0f, w1 = 0.0f, w2 = 1.0f;
        if (i < s1) {
          if (P.src_ep) {
            ep = P.src_ep[i];
            w1 = P.src_w1[(std::size_t)blockIdx.z * P.nsrc + i];
            w2 = has_w2 ? P.src_w2[i] : 1.0f;
          } else {
            literal_prologue(P, i, ep, w1);
          }
        }
        s_ep[tid] = ep;
        s_w1[tid] = w1;
        s_w2[tid] = w2;
      }
      __syncthreads();
      if (j < P.nbins) {
        const int cnt = (int)min((std::size_t)kLitThreads, s1 - base);
        for (int k = 0; k < cnt; ++k) {
          const float ep = s_ep[k];
          if (ep > 0.0f) { // `if (e_peak > 0.0)`: NaN fails, +inf passes (x0 = 0 -> yfill)
            const float F = literal_interp(P, lt, e_syn / ep);
            float       t = s_w1[k] * e_syn;
            if (has_w2) {
              t = t * s_w2[k];
            }
            acc += (double)(t * F);
          }
        }
      }
    }
    if (j < P.nbins) {
      P.partials[((std::size_t)blockIdx.z * gridDim.y + blockIdx.y) * P.nbins + j] = acc;
    }
  }

  // out[b][j] = sum over slices, in slice order (b = blockIdx.y)
  __global__ void literal_reduce_kernel(const double* __restrict__ partials, int nslices, int nbins,
                                        double* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nbins) {
      return;
    }
    const double* part = partials + (std::size_t)blockIdx.y * nslices * nbins;
    double        s    = 0.0;
    for (int c = 0; c < nslices; ++c) {
      s += part[(std::size_t)c * nbins + j];
    }
    out[(std::size_t)blockIdx.y * nbins + j] = s;
  }

  // ---- the contraction form of a batch of distributions:
  //   out[b][j] = sum_g f[b][g] * K[g][j],   K[g][j] = e_syn[j] * w2[g] * F(e_syn[j] / e_peak[g])
  // K is built once per (distribution bins, photon bins) by the literal machinery (F is the
  // reference's float value), the products are formed in double: every term equals the
  // reference's float term ((f * e_syn) * w2) * F up to its three float roundings (~1e-7).
  __global__ void __launch_bounds__(kLitThreads)
    kernel_matrix_kernel(const __grid_constant__ LiteralParams P, double* __restrict__ K) {
    __shared__ LogfEntry lt[16];
    if (threadIdx.x < 16) {
      lt[threadIdx.x] = g_logf_tab[threadIdx.x];
    }
    __syncthreads();
    const int j = blockIdx.x * kLitThreads + threadIdx.x;
    const int g = blockIdx.y;
    if (j >= P.nbins) {
      return;
    }
    const float e_syn = P.bins[j];
    const float ep    = P.src_ep[g];
    double      k     = 0.0;
    if (ep > 0.0f) {
      const float F = literal_interp(P, lt, e_syn / ep);
      k = (double)e_syn * (P.src_w2 ? (double)P.src_w2[g] : 1.0) * (double)F;
    }
    K[(std::size_t)g * P.nbins + j] = k;
  }

  // C[b][j] = sum_g A[b][g] K[g][j] in fp64, g ascending (a fixed order): 64 x 64 output tile
  // per CTA, 4 x 4 per thread, K and A staged through shared memory 16 rows at a time.
  // Rows of A whose e_peak fails `e_peak > 0` are skipped like the reference skips them.
  constexpr int kCtTile = 64, kCtK = 16;
  __global__ void __launch_bounds__(256)
    dist_contract_kernel(const float* __restrict__ A, const float* __restrict__ ep,
                         const double* __restrict__ K, int nbatch, int G, int M,
                         double* __restrict__ C) {
    __shared__ double sA[kCtK][kCtTile + 1];
    __shared__ double sK[kCtK][kCtTile];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int b0 = blockIdx.y * kCtTile, j0 = blockIdx.x * kCtTile;
    double    acc[4][4] = {};
    for (int g0 = 0; g0 < G; g0 += kCtK) {
      for (int i = threadIdx.x; i < kCtK * kCtTile; i += 256) {
        const int gg = i / kCtTile, c = i % kCtTile;
        const int g  = g0 + gg;
        // A transposed into [g][b]; a skipped source contributes nothing whatever f holds
        const bool live = g < G && b0 + c < nbatch && ep[g] > 0.0f;
        sA[gg][c] = live ? (double)A[(std::size_t)(b0 + c) * G + g] : 0.0;
        sK[gg][c] = (g < G && j0 + c < M) ? K[(std::size_t)g * M + j0 + c] : 0.0;
      }
      __syncthreads();
#pragma unroll
      for (int gg = 0; gg < kCtK; ++gg) {
        double a[4], k[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          a[r] = sA[gg][ty * 4 + r];
          k[r] = sK[gg][tx * 4 + r];
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            acc[r][q] = fma(a[r], k[q], acc[r][q]);
          }
        }
      }
      __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int b = b0 + ty * 4 + r;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = j0 + tx * 4 + q;
        if (b < nbatch && j < M) {
          C[(std::size_t)b * M + j] = acc[r][q];
        }
      }
    }
  }

  // read on every call: tests switch paths with the environment variable
  std::size_t literal_max_n() {
    if (const char* s = std::getenv("RGC_LITERAL_MAX_N")) {
      return (std::size_t)std::strtoull(s, nullptr, 10);
    }
    return std::size_t(1) << 19;
  }

  // d_out[j] = sum over sources of the reference's float term (e_syn factor included),
  // on the compute stream.  Sources are the particle columns of `prtls` (first n) when
  // src_ep == nullptr, else the n host triples (src_ep, src_w1, src_w2 or nullptr).
  // Records c.ev[2] / c.ev[3] around the pair kernel.
  int run_spectrum_literal(const rgc_particles* prtls, std::size_t n, float B0, float g_syn,
                           float e_at, const float* src_ep, const float* src_w1,
                           const float* src_w2, const float* bins_e_syn, std::size_t nbins,
                           const float* tab_x, const float* tab_y, std::size_t T, double* d_out,
                           std::size_t nbatch, bool contract) {
    auto& c = ctx();
    if (nbins == 0 || nbatch == 0) {
      return RGC_OK;
    }
    if (nbatch > 65535) {
      return fail(RGC_ERR_INVALID, "at most 65535 distributions per batch (got %zu)", nbatch);
    }
    // the reference's findMinMax over the x array (tabulation.cpp:84-102)
    float xmin = tab_x[0], xmax = tab_x[0];
    for (std::size_t k = 1; k < T; ++k) {
      xmin = std::min(xmin, tab_x[k]);
      xmax = std::max(xmax, tab_x[k]);
    }
    const int nbx = (int)((nbins + kLitThreads - 1) / kLitThreads);
    // slices of sources: ~8 CTAs per SM over the whole grid, at least 8 sources each
    const std::size_t want_slices =
      (std::size_t)std::max<std::size_t>(1, 8 * (std::size_t)c.sm_count / ((std::size_t)nbx * nbatch));
    const std::size_t per_slice   = std::max<std::size_t>(8, (n + want_slices - 1) / want_slices);
    const int         nslices     = (int)std::max<std::size_t>(1, (n + per_slice - 1) / per_slice);
    // one packed upload: [bins | tab_x | tab_y | e_peak, w1, w2 of the sources]
    const std::size_t nsrc_f = src_ep ? n : 0;
    std::vector<float> pack(nbins + 2 * T + (2 + nbatch) * nsrc_f);
    std::copy(bins_e_syn, bins_e_syn + nbins, pack.begin());
    std::copy(tab_x, tab_x + T, pack.begin() + nbins);
    std::copy(tab_y, tab_y + T, pack.begin() + nbins + T);
    if (src_ep) {
      float* ps = pack.data() + nbins + 2 * T; // [e_peak | w2 | w1 rows]
      std::copy(src_ep, src_ep + n, ps);
      if (src_w2) {
        std::copy(src_w2, src_w2 + n, ps + n);
      }
      std::copy(src_w1, src_w1 + nbatch * n, ps + 2 * n);
    }
    auto align = [](std::size_t x) { return (x + 255) & ~std::size_t(255); };
    const std::size_t o_pack = 0;
    const std::size_t o_den  = align(o_pack + pack.size() * sizeof(float));
    const std::size_t o_part = align(o_den + T * sizeof(float));
    const std::size_t part_bytes =
      contract ? n * nbins * sizeof(double) : nbatch * (std::size_t)nslices * nbins * sizeof(double);
    const std::size_t total = o_part + part_bytes;
    void* scratch = nullptr;
    RGC_TRY(ensure_scratch(total, &scratch));
    char*  sb     = static_cast<char*>(scratch);
    float* d_pack = reinterpret_cast<float*>(sb + o_pack);
    RGC_TRY(copy_h2d(d_pack, pack.data(), pack.size() * sizeof(float), c.stream));
    LiteralParams P {};
    if (src_ep) {
      const float* d_src = d_pack + nbins + 2 * T;
      P.src_ep = d_src;
      P.src_w2 = src_w2 ? d_src + n : nullptr;
      P.src_w1 = d_src + 2 * n;
    } else {
      for (int d = 0; d < 3; ++d) {
        P.u[d] = prtls->col[RGC_Q_U][d];
        P.e[d] = prtls->col[RGC_Q_E][d];
        P.b[d] = prtls->col[RGC_Q_B][d];
      }
    }
    P.nsrc          = n;
    P.nbatch        = (int)nbatch;
    P.src_per_slice = per_slice;
    P.B0            = B0;
    P.g_syn         = g_syn;
    P.e_at          = e_at;
    P.bins          = d_pack;
    P.nbins         = (int)nbins;
    P.tab_x         = d_pack + nbins;
    P.tab_y         = d_pack + nbins + T;
    P.tab_den       = reinterpret_cast<const float*>(sb + o_den);
    P.T             = (int)T;
    P.xmin          = xmin;
    P.xmax          = xmax;
    static const LogfEntry host_tab[16] = RGC_LOGF_TAB_INIT;
    P.Lspan    = glibc_log10f(xmax / xmin, host_tab);
    P.partials = reinterpret_cast<double*>(sb + o_part);
    literal_den_kernel<<<(unsigned)((T + 127) / 128), 128, 0, c.stream>>>(
      P.tab_x, P.T, reinterpret_cast<float*>(sb + o_den));
    RGC_CUDA(cudaGetLastError());
    RGC_CUDA(cudaEventRecord(c.ev[2], c.stream));
    if (contract) {
      // K[g][j] once, then the fp64 contraction over the batch
      kernel_matrix_kernel<<<dim3((unsigned)nbx, (unsigned)n), kLitThreads, 0, c.stream>>>(P, P.partials);
      RGC_CUDA(cudaGetLastError());
      dist_contract_kernel<<<dim3((unsigned)((nbins + kCtTile - 1) / kCtTile),
                                  (unsigned)((nbatch + kCtTile - 1) / kCtTile)),
                             256, 0, c.stream>>>(P.src_w1, P.src_ep, P.partials, (int)nbatch, (int)n,
                                                 (int)nbins, d_out);
      RGC_CUDA(cudaGetLastError());
      RGC_CUDA(cudaEventRecord(c.ev[3], c.stream));
      count_launch(3);
      return RGC_OK;
    }
    sync_literal_kernel<<<dim3((unsigned)nbx, (unsigned)nslices, (unsigned)nbatch), kLitThreads, 0, c.stream>>>(P);
    RGC_CUDA(cudaGetLastError());
    RGC_CUDA(cudaEventRecord(c.ev[3], c.stream));
    literal_reduce_kernel<<<dim3((unsigned)((nbins + 127) / 128), (unsigned)nbatch), 128, 0, c.stream>>>(
      P.partials, nslices, (int)nbins, d_out);
    RGC_CUDA(cudaGetLastError());
    count_launch(3);
    return RGC_OK;
  }

} // namespace rgc
