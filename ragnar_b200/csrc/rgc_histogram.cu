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
#include "rgc_hist_device.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

namespace rgc {

  template <bool FOURVEL, bool WEIGHTED, bool COUNTS>
  __global__ void __launch_bounds__(kHThreads)
    energy_hist_kernel(const __grid_constant__ HistParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    HistAccum<FOURVEL, WEIGHTED, COUNTS> acc;
    acc.init(smem_raw, P);
    const int         tid    = threadIdx.x;
    const std::size_t ntiles = (P.nprtl + kHTile - 1) / kHTile;
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
          acc.add(P, px[k], py[k], pz[k]);
        }
      }
      acc.end_tile(P);
    }
    acc.finish(P, smem_raw);
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

  static std::size_t align256(std::size_t x) { return (x + 255) & ~std::size_t(255); }

  // device bytes of one job: [counts | wfx | wslow] (one exchange), clamp partials, bin info
  std::size_t hist_job_bytes(std::size_t n, int sm_count) {
    return align256(3 * n * 8) + align256((std::size_t)sm_count * 8 * 2 * sizeof(double)) +
           align256(n * sizeof(float4));
  }

  int hist_enqueue(const rgc_particles* p, std::size_t nactive, const float* bins, std::size_t n,
                   bool weighted, bool fv, bool want_counts, char* dev, HistJob& job) {
    auto& c = ctx();
    // reference particles.cpp:200-216: bins' min / max, not first / last
    float emin = std::numeric_limits<float>::max(), emax = std::numeric_limits<float>::lowest();
    for (std::size_t i = 0; i < n; ++i) {
      emin = bins[i] < emin ? bins[i] : emin;
      emax = bins[i] > emax ? bins[i] : emax;
    }
    // ---- threshold table, bin info and the bin-guess coefficients depend only on
    // (bins, fourvel, log_spaced): built once (199 bisections through libm) and kept
    const HistPlan& plan = hist_plan(bins, n, fv, weighted, emin, emax);
    job.n         = n;
    job.weighted  = weighted;
    job.inv_scale = plan.inv_scale;
    job.dev       = dev;
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
    want_counts      = want_counts || !weighted;
    const int arrays = (want_counts ? 1 : 0) + (weighted ? 1 : 0);
    int       ncopies = kHWarps;
    while (ncopies > 1 && hist_smem_bytes((int)n, ncopies, arrays) > 48 * 1024) {
      ncopies /= 2;
    }
    P.ncopies = ncopies;
    // a copy receives (kHWarps/ncopies) * 32 * kPerThread quantised weights < 2^21 per tile
    P.flush_every = std::max(1, (int)((1ull << 32) / ((unsigned long long)kWeightCap *
                                                      (kHWarps / ncopies) * 32 * kPerThread)) - 1);
    const std::size_t smem = std::max<std::size_t>(hist_smem_bytes((int)n, ncopies, arrays), 1024);

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
    per_sm = std::max(1, std::min(per_sm, 8));
    const std::size_t ntiles = (nactive + kHTile - 1) / kHTile;
    const int nctas = (int)std::min<std::size_t>((std::size_t)c.sm_count * per_sm,
                                                 std::max<std::size_t>(ntiles, 1));
    const std::size_t off_clamp = align256(3 * n * 8);
    const std::size_t off_binfo = off_clamp + align256((std::size_t)c.sm_count * 8 * 2 * sizeof(double));
    P.counts     = reinterpret_cast<unsigned long long*>(dev);
    P.wfx        = reinterpret_cast<unsigned long long*>(dev + n * 8); // adjacent: [counts | wfx | wslow]
    P.wslow      = reinterpret_cast<double*>(dev + 2 * n * 8);         // is ONE exchange
    P.clamp_part = reinterpret_cast<double*>(dev + off_clamp);
    P.binfo      = reinterpret_cast<const float4*>(dev + off_binfo);
    RGC_TRY(copy_h2d(dev + off_binfo, plan.binfo.data(), n * sizeof(float4), c.stream));
    RGC_CUDA(cudaMemsetAsync(dev, 0, off_binfo, c.stream));
    RGC_CUDA(cudaEventRecord(c.ev[0], c.stream));
    kern<<<nctas, kHThreads, smem, c.stream>>>(P);
    RGC_CUDA(cudaGetLastError());
    count_launch(1);
    RGC_CUDA(cudaEventRecord(c.ev[1], c.stream));
    // ---- combine on the device: the clamp bins' per-CTA partial sums are folded (fixed
    // order) into the fp64 slow-path array; with a communicator, one exchange of
    // [u64 counts | fixed-point sums | f64 slow path].  Integer sums are exact.
    hist_fold_clamp_kernel<<<1, 32, 0, c.stream>>>(P.clamp_part, nctas, P.wslow, (int)n);
    RGC_CUDA(cudaGetLastError());
    count_launch(1);
    return allreduce_sum_mixed(P.counts, 2 * n, n);
  }

  int hist_collect(const HistJob& job, float* out_hist, std::uint64_t* out_counts, double* out_sum64) {
    auto&             c = ctx();
    const std::size_t n = job.n;
    std::vector<unsigned long long> raw(3 * n);
    RGC_CUDA(cudaMemcpyAsync(raw.data(), job.dev, 3 * n * 8, cudaMemcpyDeviceToHost, c.stream));
    RGC_CUDA(cudaStreamSynchronize(c.stream));
    RGC_TRY(exchange_check());
    const auto* counts = raw.data();
    const auto* wfx    = raw.data() + n;
    const auto* wslow  = reinterpret_cast<const double*>(raw.data() + 2 * n);
    for (std::size_t b = 0; b < n; ++b) {
      const double sum = job.weighted ? (double)wfx[b] * job.inv_scale[b] + wslow[b] : (double)counts[b];
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

} // namespace rgc

using namespace rgc;

extern "C" {

  int rgc_energy_histogram(const rgc_particles_t* p, size_t nactive, const float* bins, size_t n,
                           int log_spaced, int fourvel, float* out_hist, uint64_t* out_counts,
                           double* out_sum64) {
    RGC_REQUIRE_INIT();
    RGC_NVTX("ComputeEnergyDistribution");
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
    auto& c       = ctx();
    void* scratch = nullptr;
    RGC_TRY(ensure_scratch(hist_job_bytes(n, c.sm_count), &scratch));
    HistJob job;
    RGC_TRY(hist_enqueue(p, nactive, bins, n, log_spaced != 0, fourvel != 0, out_counts != nullptr,
                         static_cast<char*>(scratch), job));
    RGC_TRY(hist_collect(job, out_hist, out_counts, out_sum64));
    float ms = 0.f;
    RGC_CUDA(cudaEventElapsedTime(&ms, c.ev[0], c.ev[1]));
    c.last_ms[0] = ms;
    c.last_ms[1] = ms;
    return RGC_OK;
  }

} // extern "C"
