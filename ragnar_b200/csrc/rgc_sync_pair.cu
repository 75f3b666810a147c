// Bucketed hinge form of the particle synchrotron spectrum for sm_100a — the
// FP32-pipe-bound main kernel of SynchrotronSpectrum_<D>D.
//
// Replaces (reference paths relative to haykh/ragnar @ fceb6b08):
//   sync::Kernel<D>::operator() / OmegaSync_ChiR   src/physics/synchrotron.hpp:145-232
//   InterpolateTabulatedFunction<true>             src/containers/tabulation.hpp:19-42
//
// Same sum as the gather kernel (rgc_synchrotron.cu), regrouped.  With the table
// coordinate t = a_j + c_i split into integer and fractional parts,
//     a_j = A_j + fa_j (photon bin j),   c_i = K_i + fc_i (particle i),
// every pair of one bucket (all particles with the same K) and one bin looks at
// the fixed pair of adjacent table cells q = A_j + K, q + 1, where the
// reference's piecewise-linear interpolant is, with u = fa_j + fc_i in [0, 2),
//     F = v_q + s_q * u + (s_{q+1} - s_q) * max(0, u - h_q)
// (h_q ~ 1 is the position of the table node between the two cells).  Summed
// over the bucket's particles with weights w_i = chiR_i:
//     sum_i w_i F_ij = v_q S0 + s_q (fa_j S0 + S1) + ds_q * sum_i w_i max(0, fa_j - h_q + fc_i)
// S0 = sum w_i and S1 = sum w_i fc_i do not depend on the bin; only the hinge
// needs per-pair work:  r = sat(fa'_j + fc_i);  S2_j += w_i * r   — one FADD.SAT
// and one FFMA per (particle, bin) evaluation, nothing else in the inner loop.
// (The upper clamp of .SAT never binds for real bins: fa' + fc < 1 + |h - 1|.)
// Two spare lanes per warp column run the same instructions with fa' = 1 and
// fa' = 0 and so deliver S0 and S1 for free.
//
// One CTA owns a tile of 4096 particles at a time:
//   pass 1  prologue per particle (gamma, beta, chiR, e_peak with the reference's
//           fp64 promotions) -> (bucket, fc, w) staged in shared memory, per-warp
//           bucket counts
//   scan    bucket offsets (each bucket padded to an even length with one
//           zero-weight entry), per-warp cursors
//   pass 2  warp-synchronous stable ranking (MATCH.ANY) -> bucket-sorted (fc, w)
//   pairs   the sorted range is split evenly over the warp rows; a warp walks its
//           range bucket by bucket: loads (ds, h) of its lanes' cells once per
//           segment, then runs the 2-instruction pair loop from broadcast LDS.128
// Hinge sums are float per segment, folded into fp64 per tile; bucket moments
// are fp64 per CTA; all reductions run in a fixed order: bitwise reproducible.
// The linear part  v_q S0 + s_q (fa_j S0 + S1)  is added once, in fp64, by the
// final reduction kernel from the moments summed over CTAs.
//
// Requires a table with F = 0 at both ends (true for sync::TabulateFfunc, whose
// first node is forced to 0 and whose nodes beyond x = 20 are 0); any other table,
// very wide bin ranges and the FromDist form use the gather kernel.
//
// Roofline: 2 FP32-pipe instructions per evaluation against the measured FFMA issue
// rate (rgc_measure_peak kind 0/1); HBM: 36 B per particle amortised over nbins.
//
// Compiled with -fmad=false: every FMA below is an explicit fmaf()/fma().
#include "rgc_internal.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace rgc {

  constexpr int kPThreads  = 256;
  constexpr int kPWarps    = kPThreads / 32;
  constexpr int kPTile     = 4096; // particles per CTA tile (16 per thread)
  constexpr int kPSteps    = kPTile / kPThreads;
  constexpr int kPMaxGPW   = 8;
  constexpr int kPMaxBins  = 8 * (kPMaxGPW * 32 - 2); // 2032 per launch
  constexpr int kPMaxBuckets = 1024;
  constexpr unsigned kInvalidKey = 0xffffu;
  constexpr int kSegCostDefault = 16; // per-segment overhead of the pair phase, in sorted entries

  struct PairParams {
    const float* u[3];
    const float* e[3];
    const float* b[3];
    std::size_t  nprtl;
    const int2*   slot_i;  // per slot {Aoff, -}: cell q = max(Aoff + bucket, 0); spare slots Aoff << 0
    const float2* slot_f;  // per slot {fa0, -}:  fa' = (fa0 - h_q) * sign_q
    const float4* coef_dh; // per padded cell {hinge coefficient, hinge position, sign, 0}
    int n_pad, nb, nbp, ncols;
    int kmin;              // bucket = floor(c) - kmin
    double kmin_d;
    double inv_B0;           // 1 / B0
    double e_scale;          // e_syn_at_g_syn / (g_syn * g_syn), the float product promoted
    double cells_per_octave; // log10(2) / dL
    double c0, inv_dL, c_lo, c_hi;
    // staged per-particle results of the prologue kernel, padded to whole tiles
    float2*         cw;   // (fc, w)
    unsigned short* keys; // bucket, kInvalidKey = not on the table
    std::size_t     npad; // ntiles * kPTile
    double* partials; // [cta][nslots] hinge sums
    double* moments;  // [cta][2 * nb]  S0 then S1 per bucket
    int     nslots;
    int     seg_cost; // cost-model weight of one bucket segment, in sorted entries
    // shared-memory layout (byte offsets, computed once on the host)
    int o_coef, o_s0tot, o_s1tot, o_start, o_cstart, o_hw, o_stage_cw, o_stage_k, o_sorted, o_edge, o_scan, o_mbar, o_seg_s0, o_seg_s1;
  };

  struct PairEdge {
    int   b;
    float s0, s1;
  };

  __host__ __device__ inline std::size_t pair_align16(std::size_t x) { return (x + 15) & ~std::size_t(15); }

  struct PairSmem {
    std::size_t coef, s0tot, s1tot, start, cstart, seg_s0, seg_s1, hw, stage_cw, stage_k, sorted, edge, scan, mbar, total;
  };

  __host__ __device__ inline PairSmem pair_smem_layout(int n_pad, int nb, int nbp) {
    PairSmem L;
    std::size_t o = 0;
    L.coef = o;      o = pair_align16(o + (std::size_t)n_pad * sizeof(float4));
    L.s0tot = o;     o = pair_align16(o + (std::size_t)nb * sizeof(double));
    L.s1tot = o;     o = pair_align16(o + (std::size_t)nb * sizeof(double));
    L.start = o;     o = pair_align16(o + (std::size_t)(nb + 2) * sizeof(int));
    L.cstart = o;    o = pair_align16(o + (std::size_t)(nb + 2) * sizeof(int));
    L.seg_s0 = o;    o = pair_align16(o + (std::size_t)(nb + 2) * sizeof(float));
    L.seg_s1 = o;    o = pair_align16(o + (std::size_t)(nb + 2) * sizeof(float));
    L.hw = o;        o = pair_align16(o + (std::size_t)kPWarps * nbp * sizeof(unsigned short));
    L.stage_cw = o;  o = pair_align16(o + (std::size_t)kPTile * sizeof(float2));
    L.stage_k = o;   o = pair_align16(o + (std::size_t)kPTile * sizeof(unsigned short));
    L.sorted = o;    o = pair_align16(o + (std::size_t)(kPTile + nb + 10) * sizeof(float2));
    L.edge = o;      o = pair_align16(o + (std::size_t)kPWarps * 2 * sizeof(PairEdge));
    L.scan = o;      o = pair_align16(o + (std::size_t)(2 * kPWarps + 1) * sizeof(int));
    L.mbar = o;      o = pair_align16(o + 16);
    L.total = o;
    return L;
  }

  // The prologue is bound by the XU pipe (16 lanes/clk/SM: MUFU and every
  // float<->double conversion), so it spends as few of those as it can: 9 input
  // conversions, two MUFU.RSQ64H, one MUFU.RCP64H, one I2F and two results -> float.

  // 1/sqrt(x), x a positive normal double: MUFU.RSQ64H seed (~2^-20) and one
  // second-order Newton step -> relative error < 1e-12
  __device__ __forceinline__ double rsqrt_nr(double x) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double t = x * r;
    const double e = fma(-t, r, 1.0);
    return fma(r * 0.5, e, r);
  }

  // log2 of a positive normal double, |error| < 1e-12: exponent + atanh series of
  // the mantissa folded into [sqrt(1/2), sqrt(2)); the quotient (m-1)/(m+1) comes
  // from a MUFU.RCP64H seed refined by one Newton step
  __device__ __forceinline__ double log2_pos(double x) {
    const int hi = __double2hiint(x);
    int       ex = ((hi >> 20) & 0x7ff) - 1023;
    double    m  = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
    if (m > 1.4142135623730951) {
      m *= 0.5;
      ex += 1;
    }
    const double a = m - 1.0, b = m + 1.0;
    double       rc;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rc) : "d"(b));
    rc              = fma(rc, fma(-b, rc, 1.0), rc);
    const double s  = a * rc;
    const double s2 = s * s; // <= 0.0295
    double       t  = fma(s2, 1.0 / 15.0, 1.0 / 13.0);
    t               = fma(s2, t, 1.0 / 11.0);
    t               = fma(s2, t, 1.0 / 9.0);
    t               = fma(s2, t, 1.0 / 7.0);
    t               = fma(s2, t, 1.0 / 5.0);
    t               = fma(s2, t, 1.0 / 3.0);
    t               = t * s2;
    const double sc = s * 2.8853900817779268; // 2 / ln 2
    return (double)ex + fma(sc, t, sc);
  }

  // reference src/physics/synchrotron.hpp:193-231 — gamma, beta, beta.E, beta x B,
  // chiR, e_peak — in fp64 like the reference's promoted arithmetic, then the table
  // coordinate of e_peak split into bucket and fraction.  Two deliberate, bounded
  // departures from the reference's rounding sequence (the gather kernel keeps the
  // exact one; tests pin the two paths against each other):
  //   * ux*ux etc. enter gamma^2 unrounded (the reference rounds each square to float
  //     before promoting it), and e_peak is formed from the unrounded chiR and is not
  //     itself rounded to float: e_peak moves by < 2 float ulp, i.e. the table
  //     coordinate by ~1e-6 cell;
  //   * sqrt, quotients and log10 go through rsqrt_nr / log2_pos (error < 1e-12).
  // The weight chiR is rounded to float exactly as in the reference.
  __device__ __forceinline__ bool pair_prologue(const PairParams& P, float ux, float uy, float uz,
                                                float ex, float ey, float ez, float bx, float by,
                                                float bz, unsigned& bucket, float& fc, float& w) {
    const double dux = (double)ux, duy = (double)uy, duz = (double)uz;
    const double dex = (double)ex, dey = (double)ey, dez = (double)ez;
    const double dbx = (double)bx, dby = (double)by, dbz = (double)bz;
    const double g2  = fma(duz, duz, fma(duy, duy, fma(dux, dux, 1.0)));
    const double rg  = rsqrt_nr(g2);
    const double beta_x = dux * rg, beta_y = duy * rg, beta_z = duz * rg;
    const double bde = fma(beta_z, dez, fma(beta_y, dey, beta_x * dex));
    const double sx  = dex + fma(beta_y, dbz, -(beta_z * dby));
    const double sy  = dey + fma(beta_z, dbx, -(beta_x * dbz));
    const double sz  = dez + fma(beta_x, dby, -(beta_y * dbx));
    const double q   = fma(-bde, bde, fma(sz, sz, fma(sy, sy, sx * sx)));
    // q <= 0: chiR = 0 or NaN in the reference, the pair is skipped (synchrotron.hpp:162);
    // NaN / inf inputs fail this or the range checks below
    if (!(q > 1e-280 && q < 1e280)) {
      return false;
    }
    const double chi = (q * rsqrt_nr(q)) * P.inv_B0;
    const double ep  = (P.e_scale * g2) * chi;
    // float(e_peak) must be a positive finite float (else x0 = e_syn / e_peak is
    // off the table on either side)
    if (!(ep > 1e-37 && ep < 3.4028234e38)) {
      return false;
    }
    const double c = fma(-log2_pos(ep), P.cells_per_octave, P.c0);
    if (!(c >= P.c_lo && c < P.c_hi)) {
      return false;
    }
    // floor(c) without FRND / F2I: round-to-nearest of c - 1/2 through the 2^52 trick;
    // an exact integer c may land on either neighbour, (K, fc = 0) and (K - 1, fc = 1)
    // being the same table coordinate
    const double magic = 6755399441055744.0; // 1.5 * 2^52
    const double rm    = (c - 0.5) + magic;
    int          ri    = __double2loint(rm);
    double       fl    = rm - magic;
    if (ri < P.kmin) {
      ri = P.kmin;
      fl = P.kmin_d;
    }
    bucket = (unsigned)(ri - P.kmin);
    fc     = (float)(c - fl);
    w      = (float)chi;
    return true;
  }

  // ---- TMA bulk copy + mbarrier (raw PTX; one transaction barrier per CTA)
  __device__ __forceinline__ unsigned smem_u32(const void* p) {
    return (unsigned)__cvta_generic_to_shared(p);
  }
  __device__ __forceinline__ void mbar_init(void* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  }
  __device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
  }
  __device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, void* bar) {
    asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
        smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
  }
  __device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
    unsigned ok;
    do {
      asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    } while (!ok);
  }

  // Lanes holding the same 11-bit key, from 11 ballots (MATCH.ANY takes several
  // hundred cycles when most of a warp's keys differ, which is the normal case here)
  __device__ __forceinline__ unsigned match_key11(unsigned key) {
    unsigned m = 0xffffffffu;
#pragma unroll
    for (int i = 0; i < 11; ++i) {
      const unsigned bit = (key >> i) & 1u;
      const unsigned bal = __ballot_sync(0xffffffffu, bit != 0u);
      m &= bit ? bal : ~bal;
    }
    return m;
  }

  // ---- kernel 1: per-particle prologue, streamed once over the particle columns
  // (36 B read, 10 B written per particle; HBM-bound).  Entries past nprtl up to the
  // end of the last tile are written as invalid so the pair kernel copies whole tiles.
  template <int MINB>
  __global__ void __launch_bounds__(256, MINB)
    sync_prologue_kernel(const __grid_constant__ PairParams P) {
    const std::size_t stride = (std::size_t)gridDim.x * blockDim.x * 4;
    for (std::size_t i0 = ((std::size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i0 < P.npad;
         i0 += stride) {
      float4 v[9];
      if (i0 < P.nprtl) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          v[d]     = __ldcs(reinterpret_cast<const float4*>(P.u[d] + i0));
          v[3 + d] = __ldcs(reinterpret_cast<const float4*>(P.e[d] + i0));
          v[6 + d] = __ldcs(reinterpret_cast<const float4*>(P.b[d] + i0));
        }
      }
      const float*   f = reinterpret_cast<const float*>(v);
      float          out_cw[8];
      unsigned short out_k[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        unsigned bucket = kInvalidKey;
        float    fc = 0.0f, w = 0.0f;
        bool     ok = false;
        if (i0 + k < P.nprtl) {
          ok = pair_prologue(P, f[0 * 4 + k], f[1 * 4 + k], f[2 * 4 + k], f[3 * 4 + k],
                             f[4 * 4 + k], f[5 * 4 + k], f[6 * 4 + k], f[7 * 4 + k],
                             f[8 * 4 + k], bucket, fc, w);
        }
        out_k[k]          = (unsigned short)(ok ? bucket : kInvalidKey);
        out_cw[2 * k]     = ok ? fc : 0.0f;
        out_cw[2 * k + 1] = ok ? w : 0.0f;
      }
      float4* cw4 = reinterpret_cast<float4*>(P.cw + i0);
      cw4[0]      = make_float4(out_cw[0], out_cw[1], out_cw[2], out_cw[3]);
      cw4[1]      = make_float4(out_cw[4], out_cw[5], out_cw[6], out_cw[7]);
      *reinterpret_cast<uint2*>(P.keys + i0) =
        make_uint2((unsigned)out_k[0] | ((unsigned)out_k[1] << 16),
                   (unsigned)out_k[2] | ((unsigned)out_k[3] << 16));
    }
  }

  // ---- kernel 2: bucket sort inside the tile + the pair loop
  template <int GPW>
  __global__ void __launch_bounds__(kPThreads, 2)
    sync_pair_kernel(const __grid_constant__ PairParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4*         coef     = reinterpret_cast<float4*>(smem_raw + P.o_coef);
    double*         s0tot    = reinterpret_cast<double*>(smem_raw + P.o_s0tot);
    double*         s1tot    = reinterpret_cast<double*>(smem_raw + P.o_s1tot);
    int*            seg_start = reinterpret_cast<int*>(smem_raw + P.o_start);   // [nb + 2]
    int*            seg_b     = reinterpret_cast<int*>(smem_raw + P.o_cstart);  // [nb + 2]
    float*          seg_s0    = reinterpret_cast<float*>(smem_raw + P.o_seg_s0);
    float*          seg_s1    = reinterpret_cast<float*>(smem_raw + P.o_seg_s1);
    unsigned short* hw16     = reinterpret_cast<unsigned short*>(smem_raw + P.o_hw);
    unsigned*       hw32     = reinterpret_cast<unsigned*>(smem_raw + P.o_hw);
    float2*         stage_cw = reinterpret_cast<float2*>(smem_raw + P.o_stage_cw);
    unsigned short* stage_k  = reinterpret_cast<unsigned short*>(smem_raw + P.o_stage_k);
    float2*         sorted   = reinterpret_cast<float2*>(smem_raw + P.o_sorted);
    PairEdge*       edge     = reinterpret_cast<PairEdge*>(smem_raw + P.o_edge);
    int*            scan_tmp = reinterpret_cast<int*>(smem_raw + P.o_scan);
    void*           mbar     = smem_raw + P.o_mbar;

    const int tid  = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int col  = warp % P.ncols;
    const int row  = warp / P.ncols;
    const int rows = kPWarps / P.ncols;
    const int nb   = P.nb;
    const int nbp  = P.nbp;

    const std::size_t ntiles = P.npad / kPTile;
    constexpr unsigned kStageBytesCW = kPTile * sizeof(float2);
    constexpr unsigned kStageBytesK  = kPTile * sizeof(unsigned short);
    // the staged (fc, w, key) of a tile arrive by TMA bulk copy; the copy of the next
    // tile is issued as soon as the sort no longer reads the staging buffers, so it
    // lands underneath the pair loop
    auto issue_tile_copy = [&](std::size_t tile) {
      mbar_expect_tx(mbar, kStageBytesCW + kStageBytesK);
      bulk_g2s(stage_cw, P.cw + tile * kPTile, kStageBytesCW, mbar);
      bulk_g2s(stage_k, P.keys + tile * kPTile, kStageBytesK, mbar);
    };
    if (tid == 0) {
      mbar_init(mbar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < P.n_pad; i += kPThreads) {
      coef[i] = P.coef_dh[i];
    }
    for (int i = tid; i < nb; i += kPThreads) {
      s0tot[i] = 0.0;
      s1tot[i] = 0.0;
    }
    if (tid < kPWarps * 2) {
      edge[tid].b = -1;
    }
    __syncthreads();
    if (tid == 0 && blockIdx.x < ntiles) {
      issue_tile_copy(blockIdx.x);
    }

    int    aoff[GPW];
    float  fa0[GPW], acc[GPW];
    double accd[GPW];
#pragma unroll
    for (int g = 0; g < GPW; ++g) {
      const int    slot = (col * GPW + g) * 32 + lane;
      const int2   si   = P.slot_i[slot];
      const float2 sf   = P.slot_f[slot];
      aoff[g] = si.x;
      fa0[g]  = sf.x;
      acc[g]  = 0.0f;
      accd[g] = 0.0;
    }

    const int bpt    = (nb + kPThreads - 1) / kPThreads; // buckets per thread in the scan
    unsigned  parity = 0;

    for (std::size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      // ---- per-warp bucket cursors (u16), zeroed; the previous tile's pair phase is
      // complete once every warp has passed this barrier
      for (int i = tid; i < kPWarps * nbp / 2; i += kPThreads) {
        hw32[i] = 0u;
      }
      __syncthreads();
      mbar_wait(mbar, parity);
      parity ^= 1u;
      // ---- pass A: stable rank of every particle among its warp's particles of the
      // same bucket (warp w owns tile entries [512 w, 512 w + 512), 32 per step).
      // The lanes of a step are grouped by key (match_key11); the group's first lane
      // advances the warp's cursor of that bucket.
      unsigned short* cur = hw16 + warp * nbp;
      unsigned        rank_pack[kPSteps / 2]; // u16 ranks, two per register
#pragma unroll
      for (int s4 = 0; s4 < kPSteps; s4 += 4) {
        unsigned key[4], m[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          key[j] = stage_k[warp * (kPTile / kPWarps) + (s4 + j) * 32 + lane];
          m[j]   = match_key11(key[j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool valid = key[j] != kInvalidKey;
          const int  lead  = __ffs(m[j]) - 1;
          int        basev = 0;
          if (valid && lane == lead) {
            basev       = cur[key[j]];
            cur[key[j]] = (unsigned short)(basev + __popc(m[j]));
          }
          basev = __shfl_sync(0xffffffffu, basev, lead);
          const unsigned rk = (unsigned)(basev + __popc(m[j] & ((1u << lane) - 1u)));
          if (((s4 + j) & 1) == 0) {
            rank_pack[(s4 + j) >> 1] = rk;
          } else {
            rank_pack[(s4 + j) >> 1] |= rk << 16;
          }
          __syncwarp();
        }
      }
      __syncthreads();
      // ---- scan: bucket totals (padded to even) -> per-warp cursors and the list
      // of non-empty buckets ("segments": bucket id + start in the sorted array).
      // Entries and segments are scanned together, packed 16 + 16 bits.
      {
        int sum = 0;
        for (int i = 0; i < bpt; ++i) {
          const int b = tid * bpt + i;
          if (b < nb) {
            int tot = 0;
#pragma unroll
            for (int wq = 0; wq < kPWarps; ++wq) {
              tot += hw16[wq * nbp + b];
            }
            const int pc = (tot + 1) & ~1;
            sum += pc + (pc ? 0x10000 : 0);
          }
        }
        int incl = sum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, off);
          if (lane >= off) {
            incl += t;
          }
        }
        if (lane == 31) {
          scan_tmp[warp] = incl;
        }
        __syncthreads();
        int warp_base = 0, total = 0;
#pragma unroll
        for (int wq = 0; wq < kPWarps; ++wq) {
          const int c = scan_tmp[wq];
          warp_base += wq < warp ? c : 0;
          total += c;
        }
        int packed = warp_base + incl - sum;
        for (int i = 0; i < bpt; ++i) {
          const int b = tid * bpt + i;
          if (b < nb) {
            const int off = packed & 0xffff, k = packed >> 16;
            int       run = off;
#pragma unroll
            for (int wq = 0; wq < kPWarps; ++wq) {
              const int c        = hw16[wq * nbp + b];
              hw16[wq * nbp + b] = (unsigned short)run;
              run += c;
            }
            if (run != off) {
              if ((run - off) & 1) {
                sorted[run] = make_float2(0.0f, 0.0f); // zero-weight pad
                ++run;
              }
              seg_start[k] = off;
              seg_b[k]     = b;
              seg_s0[k]    = 0.0f;
              seg_s1[k]    = 0.0f;
              packed += (run - off) + 0x10000;
            }
          }
        }
        if (tid == 0) {
          const int nseg  = total >> 16;
          seg_start[nseg] = total & 0xffff;
          seg_b[nseg]     = 0;
          scan_tmp[kPWarps] = nseg;
        }
      }
      __syncthreads();
      // ---- pass B: scatter to bucket order (cursor base of (warp, bucket) + rank)
      {
        const unsigned short* basep = hw16 + warp * nbp;
#pragma unroll
        for (int step = 0; step < kPSteps; ++step) {
          const int      idx = warp * (kPTile / kPWarps) + step * 32 + lane;
          const unsigned key = stage_k[idx];
          if (key != kInvalidKey) {
            const unsigned rk = (rank_pack[step >> 1] >> ((step & 1) * 16)) & 0xffffu;
            sorted[(unsigned)basep[key] + rk] = stage_cw[idx];
          }
        }
      }
      __syncthreads();
      // staging buffers are free again: fetch the next tile underneath the pair loop
      if (tid == 0 && tile + gridDim.x < ntiles) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue_tile_copy(tile + gridDim.x);
      }
      // ---- pair phase: the segment list is split between the warp rows by the cost
      // model  entries + seg_cost * segments;  a row boundary inside a bucket splits
      // that bucket's segment
      {
        const int nseg  = scan_tmp[kPWarps];
        const int total = seg_start[nseg];
        const int segc  = P.seg_cost;
        auto pos_of_cost = [&](int target, int& k_out) -> int {
          int l = 0, h = max(nseg - 1, 0);
          while (l < h) { // largest k with cost_before(k) <= target
            const int mid = (l + h + 1) >> 1;
            if (seg_start[mid] + segc * mid <= target) {
              l = mid;
            } else {
              h = mid - 1;
            }
          }
          k_out           = l;
          const int s0    = seg_start[l];
          const int within = max(0, target - (s0 + segc * l) - segc) & ~1;
          return min(seg_start[l + 1], s0 + within);
        };
        const int ctotal = total + segc * nseg;
        int       k = 0, kdummy = 0;
        const int lo = row == 0 ? 0 : pos_of_cost((int)(((long long)ctotal * row) / rows), k);
        const int hi = row == rows - 1
                         ? total
                         : pos_of_cost((int)(((long long)ctotal * (row + 1)) / rows), kdummy);
        if (lo < hi) {
          if (seg_start[k + 1] <= lo) {
            ++k;
          }
          const float4* sorted4 = reinterpret_cast<const float4*>(sorted);
          int           nedges  = 0;
          // coefficients of the current segment; those of the next one are fetched
          // before the pair loop runs so their latency hides underneath it
          int   b_cur = seg_b[k];
          int   s_beg = seg_start[k], s_end = seg_start[k + 1];
          float fap[GPW], sgn[GPW], ds[GPW];
#pragma unroll
          for (int g = 0; g < GPW; ++g) {
            const float4 dh = coef[max(aoff[g] + b_cur, 0)];
            ds[g]  = dh.x;
            sgn[g] = dh.z;
            fap[g] = (fa0[g] - dh.y) * dh.z;
          }
          for (;;) {
            const int  pos  = max(s_beg, lo);
            const int  end  = min(s_end, hi);
            const bool full = (pos == s_beg) && (end == s_end);
            const bool more = s_end < hi;
            // next segment (reads one past the list's end are harmless: k + 2 <= nseg + 1
            // is guarded by `more`)
            int    b_nxt = 0, n_beg = 0, n_end = 0;
            float4 dhn[GPW];
            if (more) {
              b_nxt = seg_b[k + 1];
              n_beg = s_end;
              n_end = seg_start[k + 2];
#pragma unroll
              for (int g = 0; g < GPW; ++g) {
                dhn[g] = coef[max(aoff[g] + b_nxt, 0)];
              }
            }
            float s2[GPW];
#pragma unroll
            for (int g = 0; g < GPW; ++g) {
              s2[g] = 0.0f;
            }
            // Two particles per broadcast LDS.128; the loads of the next two float4
            // are in flight while the current two are consumed (ping-pong registers,
            // no moves).  Per particle all hinges first, then all accumulates, so no
            // FFMA waits on the FFMA.SAT just before it.  Loads past `end` stay inside
            // the sorted buffer's slack and are never consumed.
            int       p  = pos >> 1;
            const int pe = end >> 1;
            float4    q0 = sorted4[p];
            float4    q1 = sorted4[p + 1];
#define RGC_PAIR_ONE(FC, W)                                                         \
  {                                                                                 \
    float r[GPW];                                                                   \
    _Pragma("unroll") for (int g = 0; g < GPW; ++g) { r[g] = __saturatef(fmaf((FC), sgn[g], fap[g])); } \
    _Pragma("unroll") for (int g = 0; g < GPW; ++g) { s2[g] = fmaf((W), r[g], s2[g]); }    \
  }
#define RGC_PAIR_BODY(Q) RGC_PAIR_ONE((Q).x, (Q).y) RGC_PAIR_ONE((Q).z, (Q).w)
            for (; p + 4 <= pe; p += 4) {
              const float4 a0 = sorted4[p + 2];
              const float4 a1 = sorted4[p + 3];
              RGC_PAIR_BODY(q0)
              RGC_PAIR_BODY(q1)
              q0 = sorted4[p + 4];
              q1 = sorted4[p + 5];
              RGC_PAIR_BODY(a0)
              RGC_PAIR_BODY(a1)
            }
            if (p + 2 <= pe) {
              RGC_PAIR_BODY(q0)
              RGC_PAIR_BODY(q1)
              if (p + 2 < pe) {
                const float4 a0 = sorted4[p + 2];
                RGC_PAIR_BODY(a0)
              }
            } else if (p < pe) {
              RGC_PAIR_BODY(q0)
            }
#undef RGC_PAIR_BODY
#undef RGC_PAIR_ONE
#pragma unroll
            for (int g = 0; g < GPW; ++g) {
              acc[g] = fmaf(ds[g], s2[g], acc[g]);
            }
            if (col == 0 && lane >= 30) {
              // spare lanes 30 / 31 of the last group carry S0 / S1 of this segment:
              // plain stores, folded into the fp64 bucket moments after the barrier
              const float v = s2[GPW - 1];
              if (full) {
                (lane == 30 ? seg_s0 : seg_s1)[k] = v;
              } else {
                PairEdge& ed = edge[row * 2 + nedges];
                ed.b         = b_cur;
                (lane == 30 ? ed.s0 : ed.s1) = v;
              }
            }
            if (!full) {
              ++nedges;
            }
            if (!more) {
              break;
            }
            ++k;
            b_cur = b_nxt;
            s_beg = n_beg;
            s_end = n_end;
#pragma unroll
            for (int g = 0; g < GPW; ++g) {
              ds[g]  = dhn[g].x;
              sgn[g] = dhn[g].z;
              fap[g] = (fa0[g] - dhn[g].y) * dhn[g].z;
            }
          }
        }
#pragma unroll
        for (int g = 0; g < GPW; ++g) {
          accd[g] += (double)acc[g];
          acc[g] = 0.0f;
        }
      }
      __syncthreads();
      // ---- fold the segments' moments into the fp64 bucket moments: one thread per
      // segment; the (at most two per row) pieces of segments cut by a row boundary
      // are added by the same thread, in row order
      {
        const int nseg = scan_tmp[kPWarps];
        for (int kk = tid; kk < nseg; kk += kPThreads) {
          const int b  = seg_b[kk];
          double    a0 = (double)seg_s0[kk], a1 = (double)seg_s1[kk];
          for (int i = 0; i < rows * 2; ++i) {
            if (edge[i].b == b) {
              a0 += (double)edge[i].s0;
              a1 += (double)edge[i].s1;
            }
          }
          s0tot[b] += a0;
          s1tot[b] += a1;
        }
      }
      __syncthreads();
      if (tid < kPWarps * 2) {
        edge[tid].b = -1;
      }
    }

    // ---- CTA reduction over warp rows (fixed order), one partial row per CTA
    __syncthreads();
    double* red = reinterpret_cast<double*>(smem_raw + P.o_stage_cw); // 8 * GPW * 32 doubles <= 16 KB
#pragma unroll
    for (int g = 0; g < GPW; ++g) {
      red[(warp * GPW + g) * 32 + lane] = accd[g];
    }
    __syncthreads();
    if (row == 0) {
#pragma unroll
      for (int g = 0; g < GPW; ++g) {
        double s = 0.0;
        for (int r = 0; r < rows; ++r) {
          s += red[((r * P.ncols + col) * GPW + g) * 32 + lane];
        }
        P.partials[(std::size_t)blockIdx.x * P.nslots + (col * GPW + g) * 32 + lane] = s;
      }
    }
    for (int i = tid; i < nb; i += kPThreads) {
      P.moments[(std::size_t)blockIdx.x * 2 * nb + i]      = s0tot[i];
      P.moments[(std::size_t)blockIdx.x * 2 * nb + nb + i] = s1tot[i];
    }
  }

  // msum[i] = sum over CTAs (in CTA order) of moments[cta][i], i < 2 * nb
  __global__ void pair_moments_kernel(const double* __restrict__ moments, int nctas, int n2,
                                      double* __restrict__ msum) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n2) {
      return;
    }
    double s = 0.0;
    for (int c = 0; c < nctas; ++c) {
      s += moments[(std::size_t)c * n2 + i];
    }
    msum[i] = s;
  }

  // out[slot] = sum_cta hinge partials + sum_b ( v_q S0_b + s_q (fa S0_b + S1_b) )
  __global__ void pair_final_kernel(const double* __restrict__ partials, int nctas, int nslots,
                                    const int2* __restrict__ slot_i,
                                    const float2* __restrict__ slot_f,
                                    const double2* __restrict__ coef_vs,
                                    const double* __restrict__ msum, int nb,
                                    double* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nslots) {
      return;
    }
    double s = 0.0;
    for (int c = 0; c < nctas; ++c) {
      s += partials[(std::size_t)c * nslots + j];
    }
    const int2   si  = slot_i[j];
    const double fa  = (double)slot_f[j].x;
    double       lin = 0.0;
    if (si.x >= 0) {
      for (int b = 0; b < nb; ++b) {
        const double2 vs = coef_vs[si.x + b];
        const double  S0 = msum[b], S1 = msum[nb + b];
        lin += fma(vs.x, S0, vs.y * fma(fa, S0, S1));
      }
    }
    out[j] = s + lin;
  }

  // ------------------------------------------------------------------ host side
  struct PairPlan {
    std::vector<int2>    slot_i;
    std::vector<float2>  slot_f;
    std::vector<float4>  coef_dh;
    std::vector<double2> coef_vs;
    std::vector<int>     bin_of_slot;
    int    ncols { 1 }, gpw { 1 }, nslots { 0 }, n_pad { 0 }, nb { 0 }, nbp { 0 }, kmin { 0 };
    double c0 { 0 }, c_lo { 0 }, c_hi { 0 };
  };

  bool pair_path_eligible(const TablePlan& tp, const float* bins_e_syn,
                          const std::vector<int>& bins) {
    if (bins.empty() || (int)bins.size() > kPMaxBins) {
      return false;
    }
    if (tp.y.front() != 0.0 || tp.y.back() != 0.0) {
      return false; // the interpolant jumps at a table end: gather kernel
    }
    double amin = 1e300, amax = -1e300;
    for (int j : bins) {
      const double a = (std::log10((double)bins_e_syn[j]) - tp.L0) / tp.dL;
      amin = std::min(amin, a);
      amax = std::max(amax, a);
    }
    const double spread = amax - amin;
    return (double)tp.T + std::ceil(spread) + 2.0 <= (double)kPMaxBuckets;
  }

  static void make_pair_plan(const TablePlan& tp, const float* bins_e_syn,
                             const std::vector<int>& bins, PairPlan& pp) {
    const int           nbin = (int)bins.size();
    std::vector<double> a(nbin);
    double              amin = 1e300, amax = -1e300;
    for (int s = 0; s < nbin; ++s) {
      a[s] = (std::log10((double)bins_e_syn[bins[s]]) - tp.L0) / tp.dL;
      amin = std::min(amin, a[s]);
      amax = std::max(amax, a[s]);
    }
    const double spread = amax - amin;
    const int    T      = (int)tp.T;
    const int    pad_lo = (int)std::ceil(spread) + 1;
    pp.n_pad            = pad_lo + T + (int)std::ceil(spread) + 3;
    // t_pad = (a_j - amin) + c',  c' = c + amin + pad_lo,  c = -(log10 e_peak)/dL
    pp.c0   = amin + (double)pad_lo;
    pp.c_lo = (double)pad_lo - spread;     // t_real > 0 for the highest bin
    pp.c_hi = (double)(T - 1 + pad_lo);    // t_real < T - 1 for the lowest bin
    pp.kmin = (int)std::floor(pp.c_lo);
    pp.nb   = (int)std::floor(pp.c_hi) - pp.kmin + 1;
    pp.nbp  = (pp.nb + 1) & ~1;
    // per padded cell: value at the cell's left edge, slope, slope change at the
    // node that ends the cell, position of that node relative to the left edge
    //
    // A cell pair (k, k+1) is written from the side that keeps exact zeros exact:
    //   L  F = v_k + s_k u + (s_{k+1} - s_k) max(0, u - h)      (line of cell k + hinge)
    //   R  F = v_k max(0, 1 - u)     when cell k+1 is identically zero (the table's
    //      upper end): particles in the zero cell then contribute exactly 0, as in
    //      the reference.  The node between the cells is taken at its nominal
    //      position here (it sits within ~1e-5 cell of it; the line is pinned at the
    //      cell's left edge, so F moves by < 1e-5 |v_k| inside this one cell).
    pp.coef_dh.assign(pp.n_pad, make_float4(0.0f, 1.0f, 1.0f, 0.0f));
    pp.coef_vs.assign(pp.n_pad, make_double2(0.0, 0.0));
    auto zero_cell = [&](int k) -> bool { // real cell k; outside [0, T-2] the table is 0
      return k < 0 || k > T - 2 || (tp.y[k] == 0.0 && tp.y[k + 1] == 0.0);
    };
    auto slope = [&](int k) -> double {
      if (k < 0 || k > T - 2) {
        return 0.0;
      }
      return (tp.y[k + 1] - tp.y[k]) / (tp.tx[k + 1] - tp.tx[k]);
    };
    for (int k = -1; k <= T - 2; ++k) {
      const int    q  = pad_lo + k;
      const double sk = slope(k);
      const double vk = k >= 0 ? tp.y[k] + sk * ((double)k - tp.tx[k]) : 0.0;
      if (!zero_cell(k) && zero_cell(k + 1)) {
        pp.coef_dh[q] = make_float4((float)vk, 1.0f, -1.0f, 0.0f);
      } else {
        pp.coef_vs[q] = make_double2(vk, sk);
        pp.coef_dh[q] = make_float4((float)(slope(k + 1) - sk),
                                    (float)(tp.tx[k + 1] - (double)k), 1.0f, 0.0f);
      }
    }
    // slots: every warp column keeps its last two lanes for S0 / S1
    const int cap1 = kPMaxGPW * 32 - 2;
    pp.ncols       = 1;
    while (pp.ncols < 8 && (nbin + pp.ncols - 1) / pp.ncols > cap1) {
      pp.ncols *= 2;
    }
    const int per_col = (nbin + pp.ncols - 1) / pp.ncols;
    pp.gpw            = (per_col + 2 + 31) / 32;
    pp.nslots         = pp.ncols * pp.gpw * 32;
    // spare slots sit on padded cell 0 = {0, h = 1, +1}:  fa' = fa0 - 1
    pp.slot_i.assign(pp.nslots, make_int2(-(1 << 20), 0));
    pp.slot_f.assign(pp.nslots, make_float2(1.0f, 0.0f));
    pp.bin_of_slot.assign(pp.nslots, -1);
    const int cap = pp.gpw * 32 - 2;
    for (int s = 0; s < nbin; ++s) {
      const int    c    = s / cap, r = s % cap;
      const int    slot = c * pp.gpw * 32 + r;
      const double rel  = a[s] - amin;
      double       A    = std::floor(rel);
      float        fa   = (float)(rel - A);
      if (fa >= 1.0f) { // rounding of the fraction to float
        fa = 0.0f;
        A += 1.0;
      }
      pp.slot_i[slot]      = make_int2((int)A + pp.kmin, 1);
      pp.slot_f[slot]      = make_float2(fa, 1.0f);
      pp.bin_of_slot[slot] = bins[s];
    }
    for (int c = 0; c < pp.ncols; ++c) {
      const int last = (c + 1) * pp.gpw * 32;
      pp.slot_f[last - 2] = make_float2(2.0f, 0.0f); // r = sat(1 + fc) = 1   -> S0
      pp.slot_f[last - 1] = make_float2(1.0f, 0.0f); // r = sat(fc)     = fc  -> S1
    }
  }

  template <int G>
  static int launch_pair_g(dim3 grid, std::size_t smem, cudaStream_t st, const PairParams& P) {
    auto kern = sync_pair_kernel<G>;
    RGC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kPThreads, smem, st>>>(P);
    return RGC_OK;
  }

  static int launch_pair(int gpw, dim3 grid, std::size_t smem, cudaStream_t st,
                         const PairParams& P) {
    switch (gpw) {
      case 1: return launch_pair_g<1>(grid, smem, st, P);
      case 2: return launch_pair_g<2>(grid, smem, st, P);
      case 3: return launch_pair_g<3>(grid, smem, st, P);
      case 4: return launch_pair_g<4>(grid, smem, st, P);
      case 5: return launch_pair_g<5>(grid, smem, st, P);
      case 6: return launch_pair_g<6>(grid, smem, st, P);
      case 7: return launch_pair_g<7>(grid, smem, st, P);
      case 8: return launch_pair_g<8>(grid, smem, st, P);
    }
    return fail(RGC_ERR_INVALID, "internal: bad groups per warp %d", gpw);
  }

  // One launch over one chunk of bins.  acc[s] = sum_i w_i F_is for s < bins.size()
  // (before the e_syn factor), in the caller's chunk order.
  int run_spectrum_pair(const rgc_particles_t* prtls, std::size_t n, float B0, float g_syn,
                        float e_at, const TablePlan& tp, const float* bins_e_syn,
                        const std::vector<int>& bins, std::vector<double>& acc, float* main_ms) {
    auto&    c = ctx();
    PairPlan pp;
    make_pair_plan(tp, bins_e_syn, bins, pp);
    const PairSmem    L      = pair_smem_layout(pp.n_pad, pp.nb, pp.nbp);
    const std::size_t smem   = L.total;
    const int         per_sm = smem <= 113 * 1024 ? 2 : 1;
    if (smem > 227 * 1024) {
      return fail(RGC_ERR_INVALID, "internal: pair kernel needs %zu B of shared memory", smem);
    }
    // particles are processed in chunks so the staged (fc, w, key) stay bounded
    // (10 B per particle); chunk results are summed on the host in chunk order
    const std::size_t chunk_max = std::size_t(1) << 27;
    const std::size_t nchunk0   = std::min(n, chunk_max);
    const std::size_t ntiles0   = (nchunk0 + kPTile - 1) / kPTile;
    const std::size_t npad0     = ntiles0 * kPTile;
    const int nctas = (int)std::min<std::size_t>((std::size_t)c.sm_count * per_sm,
                                                 std::max<std::size_t>(ntiles0, 1));
    auto align = [](std::size_t x) { return (x + 255) & ~std::size_t(255); };
    const std::size_t off_si   = 0;
    const std::size_t off_sf   = align(off_si + pp.nslots * sizeof(int2));
    const std::size_t off_dh   = align(off_sf + pp.nslots * sizeof(float2));
    const std::size_t off_vs   = align(off_dh + pp.n_pad * sizeof(float4));
    const std::size_t off_msum = align(off_vs + pp.n_pad * sizeof(double2));
    const std::size_t off_out  = align(off_msum + 2 * pp.nb * sizeof(double));
    const std::size_t off_part = align(off_out + pp.nslots * sizeof(double));
    const std::size_t off_mom  = align(off_part + (std::size_t)nctas * pp.nslots * sizeof(double));
    const std::size_t off_cw   = align(off_mom + (std::size_t)nctas * 2 * pp.nb * sizeof(double));
    const std::size_t off_keys = align(off_cw + npad0 * sizeof(float2));
    const std::size_t total    = off_keys + npad0 * sizeof(unsigned short);
    void*             scratch  = nullptr;
    RGC_TRY(ensure_scratch(total, &scratch));
    char* sb = static_cast<char*>(scratch);
    RGC_CUDA(cudaMemcpyAsync(sb + off_si, pp.slot_i.data(), pp.nslots * sizeof(int2),
                             cudaMemcpyHostToDevice, c.stream));
    RGC_CUDA(cudaMemcpyAsync(sb + off_sf, pp.slot_f.data(), pp.nslots * sizeof(float2),
                             cudaMemcpyHostToDevice, c.stream));
    RGC_CUDA(cudaMemcpyAsync(sb + off_dh, pp.coef_dh.data(), pp.n_pad * sizeof(float4),
                             cudaMemcpyHostToDevice, c.stream));
    RGC_CUDA(cudaMemcpyAsync(sb + off_vs, pp.coef_vs.data(), pp.n_pad * sizeof(double2),
                             cudaMemcpyHostToDevice, c.stream));
    PairParams P {};
    P.slot_i   = reinterpret_cast<const int2*>(sb + off_si);
    P.slot_f   = reinterpret_cast<const float2*>(sb + off_sf);
    P.coef_dh  = reinterpret_cast<const float4*>(sb + off_dh);
    P.n_pad    = pp.n_pad;
    P.nb       = pp.nb;
    P.nbp      = pp.nbp;
    P.ncols    = pp.ncols;
    P.kmin     = pp.kmin;
    P.kmin_d   = (double)pp.kmin;
    P.inv_B0           = 1.0 / (double)B0;
    P.e_scale          = (double)e_at / (double)(g_syn * g_syn);
    P.cells_per_octave = 0.30102999566398119521 / tp.dL;
    P.c0       = pp.c0;
    P.inv_dL   = 1.0 / tp.dL;
    P.c_lo     = pp.c_lo;
    P.c_hi     = pp.c_hi;
    P.cw       = reinterpret_cast<float2*>(sb + off_cw);
    P.keys     = reinterpret_cast<unsigned short*>(sb + off_keys);
    P.partials = reinterpret_cast<double*>(sb + off_part);
    P.moments  = reinterpret_cast<double*>(sb + off_mom);
    P.nslots   = pp.nslots;
    {
      const char* sc = std::getenv("RGC_PAIR_SEG_COST"); // tuning knob
      P.seg_cost     = sc ? std::max(0, std::atoi(sc)) : kSegCostDefault;
    }
    P.o_coef = (int)L.coef; P.o_s0tot = (int)L.s0tot; P.o_s1tot = (int)L.s1tot;
    P.o_start = (int)L.start; P.o_cstart = (int)L.cstart; P.o_hw = (int)L.hw;
    P.o_stage_cw = (int)L.stage_cw; P.o_stage_k = (int)L.stage_k; P.o_sorted = (int)L.sorted;
    P.o_edge = (int)L.edge; P.o_scan = (int)L.scan; P.o_mbar = (int)L.mbar;
    P.o_seg_s0 = (int)L.seg_s0; P.o_seg_s1 = (int)L.seg_s1;
    double* d_msum = reinterpret_cast<double*>(sb + off_msum);
    double* d_out  = reinterpret_cast<double*>(sb + off_out);
    std::vector<double> out_host(pp.nslots), out_sum(pp.nslots, 0.0);
    float               pro_ms = 0.f;
    const char*         pm       = std::getenv("RGC_PROLOGUE_MINB"); // tuning knob
    const int           pro_minb = pm ? std::atoi(pm) : 3;
    for (std::size_t off = 0; off < n; off += chunk_max) {
      const std::size_t cnt = std::min(chunk_max, n - off);
      for (int d = 0; d < 3; ++d) {
        P.u[d] = prtls->col[RGC_Q_U][d] + off;
        P.e[d] = prtls->col[RGC_Q_E][d] + off;
        P.b[d] = prtls->col[RGC_Q_B][d] + off;
      }
      P.nprtl = cnt;
      P.npad  = ((cnt + kPTile - 1) / kPTile) * kPTile;
      const int grid1 = (int)std::min<std::size_t>((std::size_t)c.sm_count * 8,
                                                   (P.npad / 4 + 255) / 256);
      const int grid2 = (int)std::min<std::size_t>((std::size_t)nctas, P.npad / kPTile);
      RGC_CUDA(cudaEventRecord(c.ev[2], c.stream));
      if (pro_minb == 4) {
        sync_prologue_kernel<4><<<grid1, 256, 0, c.stream>>>(P);
      } else {
        sync_prologue_kernel<3><<<grid1, 256, 0, c.stream>>>(P);
      }
      RGC_CUDA(cudaGetLastError());
      RGC_CUDA(cudaEventRecord(c.ev[3], c.stream));
      RGC_TRY(launch_pair(pp.gpw, dim3(grid2), smem, c.stream, P));
      RGC_CUDA(cudaGetLastError());
      RGC_CUDA(cudaEventRecord(c.ev[4], c.stream));
      pair_moments_kernel<<<(2 * pp.nb + 127) / 128, 128, 0, c.stream>>>(P.moments, grid2,
                                                                         2 * pp.nb, d_msum);
      RGC_CUDA(cudaGetLastError());
      pair_final_kernel<<<(pp.nslots + 63) / 64, 64, 0, c.stream>>>(
        P.partials, grid2, pp.nslots, P.slot_i, P.slot_f,
        reinterpret_cast<const double2*>(sb + off_vs), d_msum, pp.nb, d_out);
      RGC_CUDA(cudaGetLastError());
      count_launch(4);
      RGC_CUDA(cudaMemcpyAsync(out_host.data(), d_out, pp.nslots * sizeof(double),
                               cudaMemcpyDeviceToHost, c.stream));
      RGC_CUDA(cudaStreamSynchronize(c.stream));
      float ms = 0.f;
      RGC_CUDA(cudaEventElapsedTime(&ms, c.ev[3], c.ev[4]));
      if (main_ms) {
        *main_ms += ms;
      }
      RGC_CUDA(cudaEventElapsedTime(&ms, c.ev[2], c.ev[3]));
      pro_ms += ms;
      for (int s2 = 0; s2 < pp.nslots; ++s2) {
        out_sum[s2] += out_host[s2];
      }
    }
    c.last_ms[2] = pro_ms;
    out_host.swap(out_sum);
    // bins[] is in chunk order; slots were filled in the same order
    acc.assign(bins.size(), 0.0);
    const int cap = pp.gpw * 32 - 2;
    for (std::size_t s = 0; s < bins.size(); ++s) {
      const int cidx = (int)s / cap, r = (int)s % cap;
      acc[s]         = out_host[(std::size_t)cidx * pp.gpw * 32 + r];
    }
    return RGC_OK;
  }

} // namespace rgc
