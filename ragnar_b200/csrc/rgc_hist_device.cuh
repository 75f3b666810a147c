// Device-side accumulation of the energy histogram (energy_hist_kernel, rgc_histogram.cu) and
// the enqueue / collect split of its host side.  See rgc_histogram.cu for the method;
// reference src/containers/particles.cpp:189-260.
// (Accumulating the histogram inside the hinge pipeline's prologue kernel, which holds U in
// registers anyway, was built and measured in round 2: 1.07 ms against 0.81 + 0.29 ms for the
// two kernels — both are bound by instruction latency, not by HBM, so sharing the pass saves
// nothing; profiles/r2_mb_pair.txt.)
#ifndef RGC_HIST_DEVICE_CUH
#define RGC_HIST_DEVICE_CUH

#include "rgc_internal.hpp"

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
    int           flush_every; // tiles of kHTile particles between u32 -> u64 weight flushes
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

  // shared memory of one CTA of kHThreads threads
  __host__ __device__ inline std::size_t hist_smem_bytes(int n, int ncopies, int arrays) {
    return (std::size_t)n * (sizeof(float4) + 8 + 8) + (std::size_t)ncopies * arrays * n * 4;
  }

  // Block-private bins of one CTA (kHThreads threads).  Usage: init(); add() per particle;
  // end_tile() after every kHTile particles the CTA has binned (all threads); finish().
  template <bool FOURVEL, bool WEIGHTED, bool COUNTS>
  struct HistAccum {
    float4*             binfo;
    unsigned long long* bfx;
    double*             bslow;
    unsigned*           wcnt; // [ncopies][n]
    unsigned*           wfx;  // [ncopies][n]
    unsigned*           my_cnt;
    unsigned*           my_fx;
    int                 n;
    float               nm1f;
    unsigned long long  lo_cnt, hi_cnt;
    double              lo_sum, hi_sum;
    float               lo_f, hi_f;
    unsigned            lo_c, hi_c;
    int                 since_flush;

    __device__ __forceinline__ void init(unsigned char* smem_raw, const HistParams& P) {
      n     = P.n;
      binfo = reinterpret_cast<float4*>(smem_raw);
      bfx   = reinterpret_cast<unsigned long long*>(binfo + n);
      bslow = reinterpret_cast<double*>(bfx + n);
      wcnt  = reinterpret_cast<unsigned*>(bslow + n);
      wfx   = wcnt + (COUNTS ? (std::size_t)P.ncopies * n : 0);
      const int tid  = threadIdx.x;
      const int copy = (tid >> 5) % P.ncopies;
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
      my_cnt = wcnt + (std::size_t)copy * n;
      my_fx  = wfx + (std::size_t)copy * n;
      lo_cnt = hi_cnt = 0ull;
      lo_sum = hi_sum = 0.0;
      lo_f = hi_f = 0.0f;
      lo_c = hi_c = 0u;
      nm1f        = (float)(n - 1);
      since_flush = 0;
      __syncthreads();
    }

    // one particle; reference particles.cpp:228-252
    __device__ __forceinline__ void add(const HistParams& P, float ux, float uy, float uz) {
      // float, left to right, unfused (particles.cpp:228-230)
      const float U = (ux * ux + uy * uy) + uz * uz;
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

    // after every kHTile particles of the CTA (each thread added at most kPerThread)
    __device__ __forceinline__ void end_tile(const HistParams& P) {
      lo_cnt += lo_c;
      hi_cnt += hi_c;
      lo_c = hi_c = 0u;
      if (WEIGHTED) {
        lo_sum += (double)lo_f;
        hi_sum += (double)hi_f;
        lo_f = hi_f = 0.0f;
        if (++since_flush == P.flush_every) {
          since_flush = 0;
          __syncthreads();
          for (int i = threadIdx.x; i < n; i += kHThreads) {
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

    // CTA epilogue: private copies -> global u64 (exact, order independent); the clamp bins'
    // weight sums go to clamp_part[blockIdx.x] (folded in a fixed order by hist_fold_clamp_kernel)
    __device__ __forceinline__ void finish(const HistParams& P, unsigned char* smem_raw) {
      const int tid = threadIdx.x, warp = tid >> 5;
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
  };

  // One histogram in flight on the compute stream (rgc_histogram.cu): enqueue = plan, upload,
  // kernel, clamp fold, all-reduce — no synchronisation; collect = D2H + synchronise + convert.
  // rgc_hist_and_spectrum enqueues the histogram, runs the spectrum pipeline behind it and
  // collects both after ONE wait.
  struct HistJob {
    std::size_t n { 0 };
    bool        weighted { false };
    std::vector<double> inv_scale;
    char*       dev { nullptr }; // [counts | wfx | wslow] contiguous, then clamp partials, binfo
  };
  std::size_t hist_job_bytes(std::size_t n, int sm_count);
  int hist_enqueue(const rgc_particles* p, std::size_t nactive, const float* bins, std::size_t n,
                   bool log_spaced, bool fourvel, bool want_counts, char* dev, HistJob& job);
  int hist_collect(const HistJob& job, float* out_hist, std::uint64_t* out_counts, double* out_sum64);

} // namespace rgc

#endif // RGC_HIST_DEVICE_CUH
