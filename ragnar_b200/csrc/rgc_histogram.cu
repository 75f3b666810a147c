// Energy (gamma*beta or gamma) histogram kernel for sm_100a.
//
// Replaces Particles<D>::energyDistribution, reference
// src/containers/particles.cpp:189-260 (Kokkos RangePolicy + ScatterView<float>).
//
// Bit-exact binning without fp64 transcendentals on the device: the reference's
// index  size_t(float(n-1) * |log10(energy/emin)| / log10f(emax/emin))  (evaluated
// in double) is a monotone step function of the float
// Usqr = (ux*ux + uy*uy) + uz*uz.  The host bisects, with the reference's exact
// formula and the same libm, the n-1 float values of Usqr at which the index
// steps (rgc_hostmath.cpp: host_energy_bin_index) and uploads them; the device
// forms Usqr with the reference's unfused float arithmetic, guesses the bin from
// one MUFU.LG2, and verifies the guess against the two neighbouring thresholds
// (a binary search over the thresholds is the always-correct fallback).  The
// clamps (energy < emin -> 0, energy >= emax -> n-1) are part of the same step
// function.
//
// Accumulation: block-private shared-memory bins, one private copy per warp.
//   counts   u32 shared atomics (native ATOMS.ADD), summed to u64 at the end
//   weights  1/energy is quantised per bin to unsigned fixed point (scale chosen
//            from the bin's largest possible weight, 20 significant bits) so that
//            it too is a native u32 shared atomic; u32 partials are flushed to
//            u64 before they can overflow.  Integer sums are exact and
//            order-independent -> deterministic, G-independent results.
//   the two clamp bins (0 and n-1) hold all out-of-range particles: unbounded
//            weights and the worst contention, so they live in per-thread
//            registers instead; bins too wide for 20-bit fixed point fall back to
//            fp64 shared atomics.
// HBM traffic: 12 B per particle (three float columns, streamed once with
// 16-byte evict-first loads); everything else stays on chip.
//
// Compiled with -fmad=false (Usqr must not be contracted into FMAs).
#include "rgc_internal.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

namespace rgc {

  constexpr int kHThreads       = 256;
  constexpr int kHWarps         = kHThreads / 32;
  constexpr int kPerThread      = 8;                     // particles per thread per tile
  constexpr int kHTile          = kHThreads * kPerThread; // 2048
  constexpr int kWeightBits     = 20;
  constexpr unsigned kWeightCap = 1u << 21; // larger quantised weights take the slow path

  struct HistParams {
    const float*  u[3];
    std::size_t   nprtl;
    int           n;          // bins
    int           ncopies;    // private shared-memory copies (divides kHWarps)
    int           flush_every; // tiles between u32 -> u64 weight flushes
    const float4* binfo;      // per bin: thr[b], thr[b+1], weight scale, 1/scale (0 = slow path)
    float         estA, estB; // bin guess = estA * log2(X) + estB
    int           est_ok;
    // outputs
    unsigned long long* counts;  // [n]
    unsigned long long* wfx;     // [n] fixed-point weight sums
    double*             wslow;   // [n] fp64 slow-path weight sums (global atomics)
    double*             clamp_part; // [gridDim][2] per-CTA weight sums of bins 0 and n-1
  };

  __device__ __forceinline__ int search_bin(const float4* binfo, int n, float U) {
    if (U != U) {
      return n - 1; // x86 NaN -> size_t conversion as in the reference build
    }
    int lo = 0, hi = n - 1; // largest b with U >= thr[b]; thr[0] = 0
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (U >= binfo[mid].x) {
        lo = mid;
      } else {
        hi = mid - 1;
      }
    }
    return lo;
  }

  template <bool FOURVEL, bool WEIGHTED, bool COUNTS>
  __global__ void __launch_bounds__(kHThreads)
    energy_hist_kernel(const __grid_constant__ HistParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = P.n;
    float4*   binfo = reinterpret_cast<float4*>(smem_raw);
    unsigned long long* bfx = reinterpret_cast<unsigned long long*>(binfo + n);
    double*   bslow = reinterpret_cast<double*>(bfx + n);
    unsigned* wcnt  = reinterpret_cast<unsigned*>(bslow + n);          // [ncopies][n]
    unsigned* wfx   = wcnt + (COUNTS ? (std::size_t)P.ncopies * n : 0); // [ncopies][n]

    const int tid  = threadIdx.x;
    const int warp = tid >> 5;
    const int copy = warp % P.ncopies;

    for (int i = tid; i < n; i += kHThreads) {
      binfo[i] = P.binfo[i];
      bfx[i]   = 0ull;
      bslow[i] = 0.0;
    }
    for (int i = tid; i < P.ncopies * n; i += kHThreads) {
      if (COUNTS) {
        wcnt[i] = 0u;
      }
      if (WEIGHTED) {
        wfx[i] = 0u;
      }
    }
    __syncthreads();

    unsigned* my_cnt = wcnt + (std::size_t)copy * n;
    unsigned* my_fx  = wfx + (std::size_t)copy * n;

    unsigned long long lo_cnt = 0, hi_cnt = 0;
    double             lo_sum = 0.0, hi_sum = 0.0;
    const float        nm1f   = (float)(n - 1);

    const std::size_t ntiles = (P.nprtl + kHTile - 1) / kHTile;
    int               since_flush = 0;
    for (std::size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const std::size_t base = tile * kHTile;
      float4            ux[kPerThread / 4], uy[kPerThread / 4], uz[kPerThread / 4];
#pragma unroll
      for (int h = 0; h < kPerThread / 4; ++h) {
        const std::size_t i0 = base + (std::size_t)h * (kHThreads * 4) + (std::size_t)tid * 4;
        if (i0 < P.nprtl) {
          ux[h] = __ldcs(reinterpret_cast<const float4*>(P.u[0] + i0));
          uy[h] = __ldcs(reinterpret_cast<const float4*>(P.u[1] + i0));
          uz[h] = __ldcs(reinterpret_cast<const float4*>(P.u[2] + i0));
        } else {
          ux[h] = uy[h] = uz[h] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      float    lo_f = 0.0f, hi_f = 0.0f;
      unsigned lo_c = 0, hi_c = 0;
      // full tiles (all but the last) carry no per-particle bounds checks
      const bool full = base + kHTile <= P.nprtl;
#pragma unroll
      for (int h = 0; h < kPerThread / 4; ++h) {
        const std::size_t i0 = base + (std::size_t)h * (kHThreads * 4) + (std::size_t)tid * 4;
        const float*      px = reinterpret_cast<const float*>(&ux[h]);
        const float*      py = reinterpret_cast<const float*>(&uy[h]);
        const float*      pz = reinterpret_cast<const float*>(&uz[h]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (!full && i0 + k >= P.nprtl) {
            continue;
          }
          // reference particles.cpp:228-230, float, left to right, unfused
          const float U = (px[k] * px[k] + py[k] * py[k]) + pz[k] * pz[k];
          const float X = FOURVEL ? U : 1.0f + U;
          int         idx;
          float4      info;
          bool        ok = false;
          if (P.est_ok) {
            float g = fmaf(P.estA, __log2f(X), P.estB);
            g       = fminf(fmaxf(g, 0.0f), nm1f);
            idx     = (int)g;
            info    = binfo[idx];
            ok      = (U >= info.x) && (!(U >= info.y) || idx == n - 1);
          }
          if (!ok) {
            idx  = search_bin(binfo, n, U);
            info = binfo[idx];
          }
          float w = 0.0f;
          if (WEIGHTED) {
            w = rsqrtf(X); // 1/energy; the reference rounds 1.0/energy to float
          }
          if (idx == 0) {
            lo_c += 1u;
            lo_f += w;
          } else if (idx == n - 1) {
            hi_c += 1u;
            hi_f += w;
          } else {
            if (COUNTS) {
              atomicAdd(&my_cnt[idx], 1u);
            }
            if (WEIGHTED) {
              const float scaled = w * info.z;
              if (info.z > 0.0f && scaled < (float)kWeightCap) {
                atomicAdd(&my_fx[idx], __float2uint_rn(scaled));
              } else {
                atomicAdd(&bslow[idx], (double)w);
              }
            }
          }
        }
      }
      lo_cnt += lo_c;
      hi_cnt += hi_c;
      if (WEIGHTED) {
        lo_sum += (double)lo_f;
        hi_sum += (double)hi_f;
        if (++since_flush == P.flush_every) {
          since_flush = 0;
          __syncthreads();
          for (int i = tid; i < n; i += kHThreads) {
            unsigned long long s = 0;
            for (int cpy = 0; cpy < P.ncopies; ++cpy) {
              s += wfx[(std::size_t)cpy * n + i];
              wfx[(std::size_t)cpy * n + i] = 0u;
            }
            bfx[i] += s;
          }
          __syncthreads();
        }
      }
    }

    // ---- CTA epilogue: private copies -> global u64 (exact, order independent)
    __syncthreads();
    for (int i = tid; i < n; i += kHThreads) {
      unsigned long long c = 0, f = bfx[i];
      for (int cpy = 0; cpy < P.ncopies; ++cpy) {
        if (COUNTS) {
          c += wcnt[(std::size_t)cpy * n + i];
        }
        if (WEIGHTED) {
          f += wfx[(std::size_t)cpy * n + i];
        }
      }
      if (COUNTS && c) {
        atomicAdd(&P.counts[i], c);
      }
      if (WEIGHTED && f) {
        atomicAdd(&P.wfx[i], f);
      }
      if (WEIGHTED && bslow[i] != 0.0) {
        atomicAdd(&P.wslow[i], bslow[i]);
      }
    }
    // clamp bins: block reduction of the per-thread registers
    __syncthreads();
    unsigned long long* rc = reinterpret_cast<unsigned long long*>(smem_raw); // reuse
    double*             rs = reinterpret_cast<double*>(rc + 2 * kHWarps);
    for (int off = 16; off > 0; off >>= 1) {
      lo_cnt += __shfl_down_sync(0xffffffffu, lo_cnt, off);
      hi_cnt += __shfl_down_sync(0xffffffffu, hi_cnt, off);
      lo_sum += __shfl_down_sync(0xffffffffu, lo_sum, off);
      hi_sum += __shfl_down_sync(0xffffffffu, hi_sum, off);
    }
    if ((tid & 31) == 0) {
      rc[warp * 2 + 0] = lo_cnt;
      rc[warp * 2 + 1] = hi_cnt;
      rs[warp * 2 + 0] = lo_sum;
      rs[warp * 2 + 1] = hi_sum;
    }
    __syncthreads();
    if (tid == 0) {
      unsigned long long c0 = 0, c1 = 0;
      double             s0 = 0.0, s1 = 0.0;
      for (int wq = 0; wq < kHWarps; ++wq) {
        c0 += rc[wq * 2 + 0];
        c1 += rc[wq * 2 + 1];
        s0 += rs[wq * 2 + 0];
        s1 += rs[wq * 2 + 1];
      }
      // counts of the clamp bins are needed even when COUNTS is off only for the
      // unweighted histogram value, which is requested with COUNTS on
      if (c0) {
        atomicAdd(&P.counts[0], c0);
      }
      if (c1) {
        atomicAdd(&P.counts[n - 1], c1);
      }
      P.clamp_part[(std::size_t)blockIdx.x * 2 + 0] = s0;
      P.clamp_part[(std::size_t)blockIdx.x * 2 + 1] = s1;
    }
  }

  // wslow[0] += sum_cta clamp[cta][0], wslow[n-1] += sum_cta clamp[cta][1]: lane-strided
  // partial sums and a fixed shuffle tree (one warp)
  __global__ void hist_fold_clamp_kernel(const double* __restrict__ clamp, int nctas,
                                         double* __restrict__ wslow, int n) {
    const int lane = threadIdx.x;
    double    s0 = 0.0, s1 = 0.0;
    for (int b = lane; b < nctas; b += 32) {
      s0 += clamp[(std::size_t)b * 2 + 0];
      s1 += clamp[(std::size_t)b * 2 + 1];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, off);
      s1 += __shfl_xor_sync(0xffffffffu, s1, off);
    }
    if (lane == 0) {
      wslow[0] += s0;
      wslow[n - 1] += s1;
    }
  }

  // ------------------------------------------------------------------ host side
  static float u2f(std::uint32_t u) {
    float f;
    std::memcpy(&f, &u, 4);
    return f;
  }

  // smallest non-negative float Usqr whose reference bin index is >= b
  // (bit pattern 0x7f800001, a NaN, when no finite or infinite value reaches b:
  // `U >= NaN` is false on the device)
  static float threshold_for(std::size_t b, bool fourvel, float emin, float emax, std::size_t n) {
    const std::uint32_t inf_bits = 0x7f800000u;
    if (host_energy_bin_index(0.0f, fourvel, emin, emax, n) >= b) {
      return 0.0f;
    }
    if (host_energy_bin_index(u2f(inf_bits), fourvel, emin, emax, n) < b) {
      return u2f(0x7f800001u);
    }
    std::uint32_t lo = 0, hi = inf_bits; // index(lo) < b <= index(hi)
    while (hi - lo > 1) {
      const std::uint32_t mid = lo + (hi - lo) / 2;
      if (host_energy_bin_index(u2f(mid), fourvel, emin, emax, n) >= b) {
        hi = mid;
      } else {
        lo = mid;
      }
    }
    return u2f(hi);
  }


  // Host-side plan of one histogram configuration (reference particles.cpp:231-245
  // evaluated through rgc_hostmath.cpp): thresholds, per-bin fixed-point scales,
  // bin-guess coefficients.  A few recent plans are kept.
  struct HistPlan {
    std::vector<float>  bins;
    bool                fv { true }, weighted { true };
    std::vector<float4> binfo;
    std::vector<double> inv_scale;
    int                 est_ok { 0 };
    float               estA { 0 }, estB { 0 };
  };

  static const HistPlan& hist_plan(const float* bins, std::size_t n, bool fv, bool weighted,
                                   float emin, float emax) {
    static thread_local std::vector<HistPlan> cache;
    for (const auto& hp : cache) {
      if (hp.fv == fv && hp.weighted == weighted && hp.bins.size() == n &&
          std::memcmp(hp.bins.data(), bins, n * sizeof(float)) == 0) {
        return hp;
      }
    }
    if (cache.size() >= 8) {
      cache.erase(cache.begin());
    }
    cache.emplace_back();
    HistPlan& hp = cache.back();
    hp.bins.assign(bins, bins + n);
    hp.fv       = fv;
    hp.weighted = weighted;
    // ---- threshold table
    std::vector<float> thr(n + 1);
    thr[0] = 0.0f;
    for (std::size_t b = 1; b < n; ++b) {
      thr[b] = threshold_for(b, fv, emin, emax, n);
    }
    thr[n] = u2f(0x7f800001u);
    std::vector<float4>& binfo = hp.binfo;
    std::vector<double>& inv_scale = hp.inv_scale;
    binfo.assign(n, make_float4(0.f, 0.f, 0.f, 0.f));
    inv_scale.assign(n, 0.0);
    for (std::size_t b = 0; b < n; ++b) {
      float scale = 0.0f;
      if (weighted && b > 0 && b + 1 < n && thr[b] == thr[b] && thr[b + 1] == thr[b + 1] &&
          thr[b + 1] > thr[b]) {
        // all energies of the bin lie in [E(thr[b]), E(thr[b+1])): weights within
        // (1/E_hi, 1/E_lo]; fixed point is used when that range is narrow
        const double e_lo = host_energy_from_usqr(thr[b], fv);
        const double e_hi = host_energy_from_usqr(thr[b + 1], fv);
        if (e_lo > 0.0 && std::isfinite(e_hi) && e_hi / e_lo <= 4.0) {
          const double s = std::ldexp(0.98, kWeightBits) * e_lo; // w_max * s = 0.98 * 2^20
          if (s > 1e-30 && s < 1e30) {
            scale        = (float)s;
            inv_scale[b] = 1.0 / (double)scale;
          }
        }
      }
      binfo[b] = make_float4(thr[b], thr[b + 1], scale, 0.0f);
    }
    // ---- bin guess coefficients: index ~ (n-1) * (0.5*log10(X) - log10(emin)) / log10f(emax/emin)
    {
      const double den = (double)std::log10(emax / emin);
      const double A   = (double)(n - 1) * 0.5 * std::log10(2.0) / den;
      const double B   = -(double)(n - 1) * std::log10((double)emin) / den;
      hp.est_ok = (emin > 0.0f && std::isfinite(A) && std::isfinite(B) && den > 0.0) ? 1 : 0;
      hp.estA   = (float)A;
      hp.estB   = (float)B;
    }
    return hp;
  }

} // namespace rgc

using namespace rgc;

extern "C" {

  int rgc_energy_histogram(const rgc_particles_t* p, size_t nactive, const float* bins, size_t n,
                           int log_spaced, int fourvel, float* out_hist, uint64_t* out_counts,
                           double* out_sum64) {
    RGC_REQUIRE_INIT();
    if (!p || !p->allocated) {
      return fail(RGC_ERR_INVALID, "Particles not allocated");
    }
    if (nactive > p->nalloc) {
      return fail(RGC_ERR_INVALID, "nactive %zu exceeds allocation %zu", nactive, p->nalloc);
    }
    if (n == 0) {
      return RGC_OK;
    }
    if (n > 5000) {
      return fail(RGC_ERR_INVALID, "energy histogram supports at most 5000 bins (got %zu)", n);
    }
    auto& c = ctx();
    // reference particles.cpp:200-216: bins' min / max, not first / last
    float emin = std::numeric_limits<float>::max(), emax = std::numeric_limits<float>::lowest();
    for (std::size_t i = 0; i < n; ++i) {
      emin = bins[i] < emin ? bins[i] : emin;
      emax = bins[i] > emax ? bins[i] : emax;
    }
    const bool fv       = fourvel != 0;
    const bool weighted = log_spaced != 0;
    // ---- threshold table, bin info and the bin-guess coefficients depend only on
    // (bins, fourvel, log_spaced): built once (199 bisections through libm) and kept
    const HistPlan& plan = hist_plan(bins, n, fv, weighted, emin, emax);
    const std::vector<float4>& binfo     = plan.binfo;
    const std::vector<double>& inv_scale = plan.inv_scale;
    HistParams P {};
    P.est_ok = plan.est_ok;
    P.estA   = plan.estA;
    P.estB   = plan.estB;
    for (int d = 0; d < 3; ++d) {
      P.u[d] = p->col[RGC_Q_U][d];
    }
    P.nprtl = nactive;
    P.n     = (int)n;
    // private copies: as many warps' worth as fit comfortably
    const bool want_counts = !weighted || out_counts != nullptr;
    const int  arrays      = (want_counts ? 1 : 0) + (weighted ? 1 : 0);
    int        ncopies     = kHWarps;
    auto smem_for = [&](int copies) {
      return n * (sizeof(float4) + 8 + 8) + (std::size_t)copies * arrays * n * 4;
    };
    while (ncopies > 1 && smem_for(ncopies) > 48 * 1024) {
      ncopies /= 2;
    }
    P.ncopies = ncopies;
    // a copy receives (kHWarps/ncopies) * 32 * kPerThread quantised weights < 2^21 per tile
    P.flush_every = std::max(1, (int)((1ull << 32) / ((unsigned long long)kWeightCap *
                                                      (kHWarps / ncopies) * 32 * kPerThread)) - 1);
    const std::size_t smem = std::max<std::size_t>(smem_for(ncopies), 1024);

    // ---- device scratch
    const std::size_t ntiles = (nactive + kHTile - 1) / kHTile;
    auto align = [](std::size_t x) { return (x + 255) & ~std::size_t(255); };

    using kern_t = void (*)(HistParams);
    kern_t kern  = nullptr;
#define RGC_PICK(FV, W, C)                               \
  if (fv == FV && weighted == W && want_counts == C) {   \
    kern = energy_hist_kernel<FV, W, C>;                 \
  }
    RGC_PICK(true, true, true)
    RGC_PICK(true, true, false)
    RGC_PICK(true, false, true)
    RGC_PICK(false, true, true)
    RGC_PICK(false, true, false)
    RGC_PICK(false, false, true)
#undef RGC_PICK
    if (smem > 48 * 1024) {
      RGC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    int per_sm = 0;
    RGC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kHThreads, smem));
    per_sm    = std::max(1, std::min(per_sm, 8));
    int nctas = (int)std::min<std::size_t>((std::size_t)c.sm_count * per_sm,
                                           std::max<std::size_t>(ntiles, 1));

    const std::size_t off_binfo  = 0;
    const std::size_t off_counts = align(off_binfo + n * sizeof(float4));
    const std::size_t off_wfx    = off_counts + n * 8; // adjacent: one u64 all-reduce of 2n
    const std::size_t off_wslow  = off_wfx + n * 8;     // adjacent too: [counts | wfx | wslow] is one exchange
    const std::size_t off_clamp  = align(off_wslow + n * 8);
    const std::size_t total      = off_clamp + (std::size_t)nctas * 2 * sizeof(double);
    void*             scratch    = nullptr;
    RGC_TRY(ensure_scratch(total, &scratch));
    char* sbase  = static_cast<char*>(scratch);
    P.binfo      = reinterpret_cast<const float4*>(sbase + off_binfo);
    P.counts     = reinterpret_cast<unsigned long long*>(sbase + off_counts);
    P.wfx        = reinterpret_cast<unsigned long long*>(sbase + off_wfx);
    P.wslow      = reinterpret_cast<double*>(sbase + off_wslow);
    P.clamp_part = reinterpret_cast<double*>(sbase + off_clamp);

    RGC_CUDA(cudaMemcpyAsync(sbase + off_binfo, binfo.data(), n * sizeof(float4),
                             cudaMemcpyHostToDevice, c.stream));
    RGC_CUDA(cudaMemsetAsync(sbase + off_counts, 0, total - off_counts, c.stream));
    RGC_CUDA(cudaEventRecord(c.ev[0], c.stream));
    kern<<<nctas, kHThreads, smem, c.stream>>>(P);
    RGC_CUDA(cudaGetLastError());
    count_launch(1);
    RGC_CUDA(cudaEventRecord(c.ev[1], c.stream));

    // ---- combine on the device: the clamp bins' per-CTA partial sums are folded (fixed
    // order) into the fp64 slow-path array; with a communicator, one fused group of two
    // all-reduces (u64 counts + fixed-point sums, f64 slow path); then ONE D2H of
    // [counts | fixed-point sums | fp64 slow path].  Integer sums are exact.
    hist_fold_clamp_kernel<<<1, 32, 0, c.stream>>>(P.clamp_part, nctas, P.wslow, (int)n);
    RGC_CUDA(cudaGetLastError());
    count_launch(1);
    RGC_TRY(allreduce_sum_mixed(P.counts, 2 * n, n));
    std::vector<unsigned char> raw(off_clamp - off_counts);
    RGC_CUDA(cudaMemcpyAsync(raw.data(), sbase + off_counts, raw.size(), cudaMemcpyDeviceToHost,
                             c.stream));
    RGC_CUDA(cudaStreamSynchronize(c.stream));
    RGC_TRY(exchange_check());
    const auto* counts = reinterpret_cast<const unsigned long long*>(raw.data());
    const auto* wfx    = reinterpret_cast<const unsigned long long*>(raw.data() + (off_wfx - off_counts));
    const auto* wslow  = reinterpret_cast<const double*>(raw.data() + (off_wslow - off_counts));
    float ms = 0.f;
    RGC_CUDA(cudaEventElapsedTime(&ms, c.ev[0], c.ev[1]));
    c.last_ms[0] = ms;
    c.last_ms[1] = ms;

    for (std::size_t b = 0; b < n; ++b) {
      double sum;
      if (weighted) {
        sum = (double)wfx[b] * inv_scale[b] + wslow[b];
      } else {
        sum = (double)counts[b];
      }
      if (out_sum64) {
        out_sum64[b] = sum;
      }
      if (out_hist) {
        out_hist[b] = (float)sum;
      }
      if (out_counts) {
        out_counts[b] = counts[b];
      }
    }
    return RGC_OK;
  }

} // extern "C"
