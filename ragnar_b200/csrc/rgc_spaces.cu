// Device-side grids and tabulated functions for sm_100a: Linspace /
// Logspace for grids too large to build on the host, the MinMax reduction and the verify
// step of TabulatedFunction on device-resident tables, and a batched evaluator of
// InterpolateTabulatedFunction.  All results are bit-identical to the reference's host
// arithmetic.
//
// Replaces (reference paths relative to haykh/ragnar @ fceb6b08):
//   Linspace / Logspace                      src/utils/snippets.cpp:21-62
//   TabulatedFunction<LG>::findMinMax/verify src/containers/tabulation.cpp:84-117
//   InterpolateTabulatedFunction<LG>         src/containers/tabulation.hpp:19-53
//
// How bit-exactness is kept:
//   Linspace   start + i * (stop - start) / (num - 1): IEEE float operations only
//              (size_t -> float conversions round to nearest on both sides).
//   Logspace   the exponent log10f(start) + float(i) * log10f(stop / start) / float(num - 1)
//              is float arithmetic around glibc's log10f (host, two calls); the element is
//              float(pow(10.0, double(exponent))) with glibc's double pow.  The device
//              evaluates CUDA's pow (<= 2 ulp of double); the float rounding of that agrees
//              with the float rounding of glibc's result (<= 1 ulp) unless the value lies
//              within a few double ulps of a float rounding boundary (probability ~1e-8 per
//              element).  The kernel reports exactly those indices and the host re-evaluates
//              them with glibc and patches them: exact by construction, whatever num is.
//   Interpolation  the literal float sequence of tabulation.hpp:29-51 with glibc's log10f
//              restated on the device (rgc_glibc_log10f.cuh).
// HBM-bound streaming kernels: 4 B written per element (spaces), 4 B read (min / max),
// 8 B per evaluated point plus table gathers from L1/L2 (interpolation).
#include "rgc_glibc_log10f.cuh"
#include "rgc_internal.hpp"

#include <algorithm>
#include <cmath>
#include <limits>
#include <vector>

namespace rgc {

  constexpr int kSpThreads  = 256;
  constexpr int kMaxHardIdx = 4096;

  __device__ const LogfEntry g_sp_logf_tab[16] = RGC_LOGF_TAB_INIT;

  __global__ void __launch_bounds__(kSpThreads)
    linspace_kernel(float start, float stop, unsigned long long num, float* __restrict__ out) {
    const unsigned long long i = (unsigned long long)blockIdx.x * kSpThreads + threadIdx.x;
    if (i < num) {
      // `start + i * (stop - start) / (num - 1)` with i, num - 1 converted to float
      out[i] = num == 1 ? start : start + ((float)i * (stop - start)) / (float)(num - 1);
    }
  }

  // out[i] = float(pow(10.0, double(lg_start + float(i) * lg_ratio / denom))); indices whose
  // double result sits within 16 ulp of a float rounding boundary are appended to `hard`
  __global__ void __launch_bounds__(kSpThreads)
    logspace_kernel(float start, float lg_start, float lg_ratio, unsigned long long num,
                    float* __restrict__ out, unsigned long long* __restrict__ hard,
                    int* __restrict__ nhard) {
    const unsigned long long i = (unsigned long long)blockIdx.x * kSpThreads + threadIdx.x;
    if (i >= num) {
      return;
    }
    if (num == 1) {
      out[i] = start;
      return;
    }
    const float  expo = lg_start + ((float)i * lg_ratio) / (float)(num - 1);
    const double y    = pow(10.0, (double)expo);
    const float  f    = (float)y;
    out[i]            = f;
    if (isfinite(y) && y > 0.0 && isfinite(f)) {
      // the two rounding boundaries around f: midpoints to its float neighbours
      const float  dn  = __uint_as_float(__float_as_uint(f) - 1u); // f > 0
      const float  up  = __uint_as_float(__float_as_uint(f) + 1u);
      const double mlo = 0.5 * ((double)f + (double)dn);
      const double mhi = 0.5 * ((double)f + (double)up);
      const double tol = 16.0 * (y * 2.220446049250313e-16);
      if (fabs(y - mlo) <= tol || fabs(y - mhi) <= tol) {
        const int k = atomicAdd(nhard, 1);
        if (k < kMaxHardIdx) {
          hard[k] = i;
        }
      }
    }
  }

  // Kokkos MinMax reducer semantics (tabulation.cpp:84-102): `v < min` / `v > max`
  // comparisons from (max float, lowest float): NaNs never win
  __global__ void __launch_bounds__(kSpThreads)
    minmax_kernel(const float* __restrict__ x, unsigned long long n, float* __restrict__ part) {
    float mn = 3.402823466e+38f, mx = -3.402823466e+38f;
    for (unsigned long long i = (unsigned long long)blockIdx.x * kSpThreads + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * kSpThreads) {
      const float v = x[i];
      mn = v < mn ? v : mn;
      mx = v > mx ? v : mx;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float a = __shfl_xor_sync(0xffffffffu, mn, off);
      const float b = __shfl_xor_sync(0xffffffffu, mx, off);
      mn = a < mn ? a : mn;
      mx = b > mx ? b : mx;
    }
    __shared__ float smn[kSpThreads / 32], smx[kSpThreads / 32];
    if ((threadIdx.x & 31) == 0) {
      smn[threadIdx.x >> 5] = mn;
      smx[threadIdx.x >> 5] = mx;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < kSpThreads / 32; ++w) {
        mn = smn[w] < mn ? smn[w] : mn;
        mx = smx[w] > mx ? smx[w] : mx;
      }
      part[2 * blockIdx.x]     = mn;
      part[2 * blockIdx.x + 1] = mx;
    }
  }

  // InterpolateTabulatedFunction<LG> (tabulation.hpp:19-53) at n points; den (log grid) =
  // log10f(xmax / xmin) computed once
  template <bool LG>
  __global__ void __launch_bounds__(kSpThreads)
    tabulated_eval_kernel(const float* __restrict__ tx, const float* __restrict__ ty,
                          unsigned long long T, float xmin, float xmax, float span, float yfill,
                          const float* __restrict__ x0s, unsigned long long n,
                          float* __restrict__ out) {
    __shared__ LogfEntry lt[16];
    if (threadIdx.x < 16) {
      lt[threadIdx.x] = g_sp_logf_tab[threadIdx.x];
    }
    __syncthreads();
    const unsigned long long i = (unsigned long long)blockIdx.x * kSpThreads + threadIdx.x;
    if (i >= n) {
      return;
    }
    const float x0 = x0s[i];
    if (x0 < xmin || x0 >= xmax) {
      out[i] = yfill;
      return;
    }
    float v;
    if (LG) {
      v = ((float)(T - 1) * fabsf(glibc_log10f(x0 / xmin, lt))) / span;
    } else {
      v = ((float)(T - 1) * fabsf(x0 - xmin)) / span;
    }
    const bool               big = !(v < 9.2e18f); // NaN / beyond 2^63 -> index >= n - 1 on x86-64
    const unsigned long long xi  = big ? ~0ull : (unsigned long long)v;
    if (big || xi >= T - 1) {
      out[i] = ty[T - 1];
      return;
    }
    const float xk = tx[xi], xk1 = tx[xi + 1], yk = ty[xi], yk1 = ty[xi + 1];
    if (LG) {
      out[i] = (yk1 * glibc_log10f(x0 / xk, lt) + yk * glibc_log10f(xk1 / x0, lt)) /
               glibc_log10f(xk1 / xk, lt);
    } else {
      out[i] = (yk1 * (x0 - xk) + yk * (xk1 - x0)) / (xk1 - xk);
    }
  }

  static int device_minmax(const float* d, std::size_t n, float* mn, float* mx) {
    auto& c = ctx();
    *mn = std::numeric_limits<float>::max();
    *mx = std::numeric_limits<float>::lowest();
    if (n == 0) {
      return RGC_OK;
    }
    const int grid = (int)std::min<std::size_t>((std::size_t)c.sm_count * 8, (n + kSpThreads - 1) / kSpThreads);
    void*     scratch = nullptr;
    RGC_TRY(ensure_scratch((std::size_t)grid * 2 * sizeof(float), &scratch));
    minmax_kernel<<<grid, kSpThreads, 0, c.stream>>>(d, n, static_cast<float*>(scratch));
    RGC_CUDA(cudaGetLastError());
    count_launch(1);
    std::vector<float> part((std::size_t)grid * 2);
    RGC_CUDA(cudaMemcpyAsync(part.data(), scratch, part.size() * sizeof(float), cudaMemcpyDeviceToHost, c.stream));
    RGC_CUDA(cudaStreamSynchronize(c.stream));
    for (int b = 0; b < grid; ++b) {
      *mn = part[2 * b] < *mn ? part[2 * b] : *mn;
      *mx = part[2 * b + 1] > *mx ? part[2 * b + 1] : *mx;
    }
    return RGC_OK;
  }

} // namespace rgc

using namespace rgc;

extern "C" {

  int rgc_linspace_device(float start, float stop, size_t num, rgc_buf_t** out) {
    RGC_REQUIRE_INIT();
    RGC_NVTX("Linspace");
    if (start >= stop) {
      return fail(RGC_ERR_INVALID, "Linspace start must be < stop");
    }
    RGC_TRY(rgc_buf_create(RGC_F32, num, out));
    if (num == 0) {
      return RGC_OK;
    }
    auto& c = ctx();
    linspace_kernel<<<(unsigned)((num + kSpThreads - 1) / kSpThreads), kSpThreads, 0, c.stream>>>(
      start, stop, num, static_cast<float*>(rgc_buf_device_ptr(*out)));
    RGC_CUDA(cudaGetLastError());
    count_launch(1);
    RGC_CUDA(cudaStreamSynchronize(c.stream));
    return RGC_OK;
  }

  int rgc_logspace_device(float start, float stop, size_t num, rgc_buf_t** out) {
    RGC_REQUIRE_INIT();
    RGC_NVTX("Logspace");
    if (start <= 0.0 or stop <= 0.0) {
      return fail(RGC_ERR_INVALID, "Logspace start and stop must be strictly positive");
    }
    if (start >= stop) {
      return fail(RGC_ERR_INVALID, "Logspace start must be < stop");
    }
    RGC_TRY(rgc_buf_create(RGC_F32, num, out));
    if (num == 0) {
      return RGC_OK;
    }
    auto&  c     = ctx();
    float* d_out = static_cast<float*>(rgc_buf_device_ptr(*out));
    void*  scratch = nullptr;
    RGC_TRY(ensure_scratch(kMaxHardIdx * sizeof(unsigned long long) + 64, &scratch));
    auto* d_hard  = static_cast<unsigned long long*>(scratch);
    int*  d_nhard = reinterpret_cast<int*>(d_hard + kMaxHardIdx);
    RGC_CUDA(cudaMemsetAsync(d_nhard, 0, sizeof(int), c.stream));
    // the two log10f of the exponent, as the reference's host build evaluates them
    const float lg_start = std::log10(start);
    const float lg_ratio = std::log10(stop / start);
    logspace_kernel<<<(unsigned)((num + kSpThreads - 1) / kSpThreads), kSpThreads, 0, c.stream>>>(
      start, lg_start, lg_ratio, num, d_out, d_hard, d_nhard);
    RGC_CUDA(cudaGetLastError());
    count_launch(1);
    int nhard = 0;
    RGC_CUDA(cudaMemcpyAsync(&nhard, d_nhard, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    RGC_CUDA(cudaStreamSynchronize(c.stream));
    if (nhard > kMaxHardIdx) {
      // cannot happen for a sane grid (1e-8 of the elements); stay exact anyway
      std::vector<float> all(num);
      host_logspace(start, stop, num, all.data());
      RGC_TRY(copy_h2d(d_out, all.data(), num * sizeof(float), c.stream));
      RGC_CUDA(cudaStreamSynchronize(c.stream));
      return RGC_OK;
    }
    if (nhard > 0) {
      std::vector<unsigned long long> idx(nhard);
      RGC_CUDA(cudaMemcpy(idx.data(), d_hard, nhard * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
      const float denom = static_cast<float>(num - 1);
      for (unsigned long long i : idx) {
        const float expo = lg_start + static_cast<float>(i) * lg_ratio / denom;
        const float v    = static_cast<float>(std::pow(10.0, static_cast<double>(expo)));
        RGC_CUDA(cudaMemcpy(d_out + i, &v, sizeof(float), cudaMemcpyHostToDevice));
      }
    }
    return RGC_OK;
  }

  int rgc_buf_minmax(const rgc_buf_t* buf, float* min_out, float* max_out) {
    RGC_REQUIRE_INIT();
    RGC_NVTX("XMinMax");
    if (!buf || rgc_buf_dtype(buf) != RGC_F32) {
      return fail(RGC_ERR_INVALID, "rgc_buf_minmax needs a float buffer");
    }
    float mn, mx;
    RGC_TRY(device_minmax(static_cast<const float*>(rgc_buf_device_ptr(buf)), rgc_buf_size(buf), &mn, &mx));
    if (min_out) {
      *min_out = mn;
    }
    if (max_out) {
      *max_out = mx;
    }
    return RGC_OK;
  }

  int rgc_tabulated_eval(int loggrid, const rgc_buf_t* tab_x, const rgc_buf_t* tab_y, float yfill,
                         const rgc_buf_t* x0, rgc_buf_t** out) {
    RGC_REQUIRE_INIT();
    RGC_NVTX("InterpolateTabulatedFunction");
    if (!tab_x || !tab_y || !x0 || rgc_buf_dtype(tab_x) != RGC_F32 || rgc_buf_dtype(tab_y) != RGC_F32 ||
        rgc_buf_dtype(x0) != RGC_F32) {
      return fail(RGC_ERR_INVALID, "rgc_tabulated_eval needs float buffers");
    }
    const std::size_t T = rgc_buf_size(tab_x), n = rgc_buf_size(x0);
    if (rgc_buf_size(tab_y) != T) {
      return fail(RGC_ERR_INVALID, "y.size != x.size in TabulatedFunction");
    }
    const float* tx = static_cast<const float*>(rgc_buf_device_ptr(tab_x));
    float        xmin, xmax;
    RGC_TRY(device_minmax(tx, T, &xmin, &xmax));
    if (xmin >= xmax) {
      return fail(RGC_ERR_INVALID, "xmin >= xmax in TabulatedFunction");
    }
    if (loggrid && xmin <= 0.0f) {
      return fail(RGC_ERR_INVALID, "xmin <= 0.0 in Logspace TabulatedFunction");
    }
    RGC_TRY(rgc_buf_create(RGC_F32, n, out));
    if (n == 0) {
      return RGC_OK;
    }
    auto& c = ctx();
    static const LogfEntry host_tab[16] = RGC_LOGF_TAB_INIT;
    const float  span = loggrid ? glibc_log10f(xmax / xmin, host_tab) : xmax - xmin;
    const float* ty   = static_cast<const float*>(rgc_buf_device_ptr(tab_y));
    const float* xs   = static_cast<const float*>(rgc_buf_device_ptr(x0));
    float*       o    = static_cast<float*>(rgc_buf_device_ptr(*out));
    const unsigned grid = (unsigned)((n + kSpThreads - 1) / kSpThreads);
    if (loggrid) {
      tabulated_eval_kernel<true><<<grid, kSpThreads, 0, c.stream>>>(tx, ty, T, xmin, xmax, span, yfill, xs, n, o);
    } else {
      tabulated_eval_kernel<false><<<grid, kSpThreads, 0, c.stream>>>(tx, ty, T, xmin, xmax, span, yfill, xs, n, o);
    }
    RGC_CUDA(cudaGetLastError());
    count_launch(1);
    RGC_CUDA(cudaStreamSynchronize(c.stream));
    return RGC_OK;
  }

} // extern "C"
