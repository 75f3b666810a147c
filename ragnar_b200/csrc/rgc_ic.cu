// Inverse-Compton spectrum from two tabulated distributions for sm_100a.
//
// Replaces (reference paths relative to haykh/ragnar @ fceb6b08):
//   ic::KNfunc                     src/physics/ic.hpp:20-25
//   ic::Kernel::operator()         src/physics/ic.hpp:58-85
//   ICSpectrum (driver)            src/physics/ic.cpp:15-46
//
// The reference runs a rank-3 MDRange over (particle bin g, IC bin, soft-photon
// bin) and scatters float terms into spec[IC bin].  Here one CTA owns one IC bin
// (and one slice of the particle bins): lanes run over soft-photon bins, the
// CTA's particle-bin slice is staged in shared memory, every term is formed with
// the reference's exact float / double promotions (KNfunc's `0.5 * ...` makes the
// bracket a double; `2 * q * log(q)` stays float with logf) and summed in fp64 in a
// fixed order: per thread, then warp shuffles, then warps, then slices — bitwise
// reproducible.  No atomics.
//
// Index order.  The reference's MDRange is {nprtls, nsoft, nic} but the functor is
// declared (gidx, eidx, esidx) (ic.cpp:31-34 vs ic.hpp:58), so it indexes the IC
// bins with the soft-photon extent and vice versa: in bounds only when
// nsoft == nic.  This kernel iterates the documented intent (IC bins over nic,
// soft photons over nsoft), identical whenever the reference is well defined.
//
// Latency-bound at the reference's sizes (200^3 = 8e6 terms, ~10 us); replicas
// only, never all-reduced (like FromDist).
//
// Compiled with -fmad=false -prec-div=true: every float operation below rounds
// exactly as the reference's unfused host arithmetic does.
#include "rgc_internal.hpp"

#include <algorithm>
#include <vector>

namespace rgc {

  constexpr int kICThreads = 256;
  constexpr int kICSlice   = 512; // particle bins staged per CTA pass

  struct ICParams {
    const float* g;  // particle gamma bins   [ng]
    const float* f;  // particle distribution [ng]
    const float* es; // soft-photon energies  [ns]
    const float* fs; // soft-photon distribution [ns]
    const float* eic; // IC energy bins       [nic]
    int          ng, ns, nic, nsplit, islog;
    double*      partial; // [nic][nsplit]
  };

  // reference ic.hpp:20-25 — returns real_t; logf(q) is taken as the correctly
  // rounded float of the fp64 logarithm (glibc's logf is correctly rounded in all
  // but a vanishing fraction of cases)
  __device__ __forceinline__ float kn_func(float Gamma, float q) {
    const float  Gq   = Gamma * q;
    const float  lq   = (float)log((double)q);
    const float  t1   = (2.0f * q) * lq;                                   // float
    const double brk  = (double)(1.0f + 2.0f * q) +
                       ((0.5 * (double)Gq) * (double)Gq) / (double)(1.0f + Gq); // double
    const double val  = (double)t1 + (double)(1.0f - q) * brk;
    return (float)val;
  }

  __global__ void __launch_bounds__(kICThreads) ic_spectrum_kernel(const ICParams P) {
    __shared__ float  sg[kICSlice], sf[kICSlice];
    __shared__ double wsum[kICThreads / 32];
    const int   eidx  = blockIdx.x;
    const int   split = blockIdx.y;
    const float e_ic  = P.eic[eidx];
    // this CTA's contiguous slice of the particle bins
    const int per = (P.ng + P.nsplit - 1) / P.nsplit;
    const int g0  = split * per;
    const int g1  = min(P.ng, g0 + per);
    double    acc = 0.0;
    for (int gb = g0; gb < g1; gb += kICSlice) {
      const int cnt = min(kICSlice, g1 - gb);
      __syncthreads();
      for (int i = threadIdx.x; i < cnt; i += kICThreads) {
        sg[i] = P.g[gb + i];
        sf[i] = P.f[gb + i];
      }
      __syncthreads();
      for (int s = threadIdx.x; s < P.ns; s += kICThreads) {
        const float e_soft = P.es[s];
        const float f_soft = P.fs[s];
        const float ratio  = f_soft / e_soft; // (f_soft_photons / e_soft_photons)
        for (int i = 0; i < cnt; ++i) {
          const float g_prtls = sg[i];
          const float f_prtls = sf[i];
          const float Gamma   = (4.0f * g_prtls) * e_soft;
          // reference: `if (e_ic > g * Gamma / (1 + Gamma)) return;`  (NaN compares false)
          if (e_ic > (g_prtls * Gamma) / (1.0f + Gamma)) {
            continue;
          }
          const float eg    = e_ic / g_prtls;
          const float q     = eg / (Gamma * (1.0f - eg));
          const float KNval = kn_func(Gamma, q);
          float       term  = (((f_prtls * ratio) * e_ic) * e_ic) * KNval;
          term              = P.islog ? term / g_prtls : term / (g_prtls * g_prtls);
          acc += (double)term;
        }
      }
    }
    // fixed-order reduction: lanes (xor tree), then warps
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      acc += __shfl_xor_sync(0xffffffffu, acc, off);
    }
    if ((threadIdx.x & 31) == 0) {
      wsum[threadIdx.x >> 5] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < kICThreads / 32; ++w) {
        s += wsum[w];
      }
      P.partial[(std::size_t)eidx * P.nsplit + split] = s;
    }
  }

  __global__ void ic_final_kernel(const double* __restrict__ partial, int nic, int nsplit,
                                  double* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nic) {
      return;
    }
    double s = 0.0;
    for (int k = 0; k < nsplit; ++k) {
      s += partial[(std::size_t)j * nsplit + k];
    }
    out[j] = s;
  }

} // namespace rgc

using namespace rgc;

extern "C" {

  int rgc_ic_spectrum(const float* g_prtls, const float* f_prtls, size_t nprtls,
                      int islog_bins_prtls, const float* e_soft, const float* f_soft,
                      size_t nsoft, const float* bins_e_ic, size_t nic, float* out_spec,
                      double* out_spec64) {
    RGC_REQUIRE_INIT();
    RGC_NVTX("ICSpectrum");
    if (nic == 0) {
      return RGC_OK;
    }
    if (nprtls > (std::size_t)1 << 30 || nsoft > (std::size_t)1 << 30 || nic > 65535u * 32u) {
      return fail(RGC_ERR_INVALID, "ICSpectrum: grid too large (%zu x %zu x %zu)", nprtls, nsoft,
                  nic);
    }
    auto& c = ctx();
    // enough CTAs to fill the device: nic * nsplit >= 4 per SM, slices of >= 8 bins
    int nsplit = 1;
    if (nprtls > 0) {
      const std::size_t want = ((std::size_t)c.sm_count * 4 + nic - 1) / nic;
      nsplit = (int)std::max<std::size_t>(1, std::min<std::size_t>(want, (nprtls + 7) / 8));
      nsplit = std::min(nsplit, 65535);
    }
    auto align = [](std::size_t x) { return (x + 255) & ~std::size_t(255); };
    const std::size_t o_g   = 0;
    const std::size_t o_f   = align(o_g + nprtls * 4);
    const std::size_t o_es  = align(o_f + nprtls * 4);
    const std::size_t o_fs  = align(o_es + nsoft * 4);
    const std::size_t o_eic = align(o_fs + nsoft * 4);
    const std::size_t o_out = align(o_eic + nic * 4);
    const std::size_t o_par = align(o_out + nic * 8);
    const std::size_t total = o_par + nic * (std::size_t)nsplit * 8;
    void*             scratch = nullptr;
    RGC_TRY(ensure_scratch(total, &scratch));
    char* sb = static_cast<char*>(scratch);
    RGC_CUDA(cudaEventRecord(c.ev[0], c.stream));
    if (nprtls) {
      RGC_TRY(copy_h2d(sb + o_g, g_prtls, nprtls * 4, c.stream));
      RGC_TRY(copy_h2d(sb + o_f, f_prtls, nprtls * 4, c.stream));
    }
    if (nsoft) {
      RGC_TRY(copy_h2d(sb + o_es, e_soft, nsoft * 4, c.stream));
      RGC_TRY(copy_h2d(sb + o_fs, f_soft, nsoft * 4, c.stream));
    }
    RGC_TRY(copy_h2d(sb + o_eic, bins_e_ic, nic * 4, c.stream));
    ICParams P {};
    P.g       = reinterpret_cast<const float*>(sb + o_g);
    P.f       = reinterpret_cast<const float*>(sb + o_f);
    P.es      = reinterpret_cast<const float*>(sb + o_es);
    P.fs      = reinterpret_cast<const float*>(sb + o_fs);
    P.eic     = reinterpret_cast<const float*>(sb + o_eic);
    P.ng      = (int)nprtls;
    P.ns      = (int)nsoft;
    P.nic     = (int)nic;
    P.nsplit  = nsplit;
    P.islog   = islog_bins_prtls ? 1 : 0;
    P.partial = reinterpret_cast<double*>(sb + o_par);
    double* d_out = reinterpret_cast<double*>(sb + o_out);
    RGC_CUDA(cudaEventRecord(c.ev[2], c.stream));
    ic_spectrum_kernel<<<dim3((unsigned)nic, (unsigned)nsplit), kICThreads, 0, c.stream>>>(P);
    RGC_CUDA(cudaGetLastError());
    RGC_CUDA(cudaEventRecord(c.ev[3], c.stream));
    ic_final_kernel<<<(unsigned)((nic + 127) / 128), 128, 0, c.stream>>>(P.partial, (int)nic,
                                                                         nsplit, d_out);
    RGC_CUDA(cudaGetLastError());
    count_launch(2);
    std::vector<double> host(nic);
    RGC_CUDA(cudaMemcpyAsync(host.data(), d_out, nic * 8, cudaMemcpyDeviceToHost, c.stream));
    RGC_CUDA(cudaEventRecord(c.ev[1], c.stream));
    RGC_CUDA(cudaStreamSynchronize(c.stream));
    RGC_CUDA(cudaEventElapsedTime(&c.last_ms[0], c.ev[0], c.ev[1]));
    RGC_CUDA(cudaEventElapsedTime(&c.last_ms[1], c.ev[2], c.ev[3]));
    for (std::size_t j = 0; j < nic; ++j) {
      if (out_spec64) {
        out_spec64[j] = host[j];
      }
      if (out_spec) {
        out_spec[j] = (float)host[j]; // rounded once (the reference accumulates in float)
      }
    }
    return RGC_OK;
  }

} // extern "C"
