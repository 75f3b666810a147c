// Bucketed hinge form of the particle synchrotron spectrum for sm_100a — the main path
// of SynchrotronSpectrum_<D>D.
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
//     sum_i w_i F_ij = v_q S0 + s_q (fa_j S0 + S1) + ds_q * sum_i w_i max(0, fc_i + fa'_j),
// fa'_j = fa_j - h_q.  S0 = sum w_i and S1 = sum w_i fc_i do not depend on the bin; only
// the hinge needs per-pair work:  r = sat(fc_i + fa'_j);  S2_j += w_i * r   (FADD.SAT +
// FFMA).
//
// Sub-bucket decomposition (round 2).  The hinge of bin j is identically 0 for every
// particle with fc <= -fa'_j and LINEAR in fc for every particle above.  Cut the bucket
// into sub-buckets by s = floor(8 fc + phi): a sub-bucket that lies entirely below the
// bin's threshold contributes nothing, one that lies entirely above contributes
// S1_s + fa'_j S0_s (its own moments), and only the ONE sub-bucket that contains the
// threshold needs the pair loop — about one pair in eight.  Bins are therefore grouped
// into 32-lane groups by the sub-bucket of their threshold, a piece of the sorted array is
// counting-sorted by sub-bucket, and run s is streamed through the groups of sub-bucket s
// only; pair_final_kernel adds the moment part from suffix sums over the sub-buckets.
// The phase phi puts the sub-bucket boundaries into the widest gap between the bins'
// thresholds (commensurate grids put every threshold ON a multiple of 1/8); where a float
// threshold still strays across a boundary with the table nodes' 1e-5 wiggle, the plan
// makes that sub-bucket's groups see the neighbouring run too (per bucket, `extmask`).
// The first two lanes of every sub-bucket's first group carry fa' = 1 and fa' = 0 and so
// deliver the run's S0 and S1 from the same two instructions.
//
// Pipeline (one pass per <= 2^27 particles), all on the compute stream:
//   1 sync_prologue_kernel  HBM-bound (36 B read, 10 B written per particle):
//       gamma, beta, chiR, e_peak with the reference's fp64 promotions ->
//       (fc, w) and the bucket key, in particle order; bucket counts per row (a row
//       is a contiguous range of 4096-particle tiles, one row per CTA of kernel 3)
//       from a shared-memory histogram flushed by integer atomics
//   2 pair_colscan_kernel   counts[bucket][row] -> exclusive offsets inside the
//       bucket + bucket totals (one warp per bucket)
//   3 sync_sort_kernel      HBM-bound (10 B read, 8 B written): a CTA walks its row
//       tile by tile: TMA bulk copy of the tile's (fc, w, key) into shared memory,
//       ranking, scan, scatter to bucket order INSIDE shared memory, then coalesced
//       write-out of every bucket run to its place in the GLOBAL bucket order
//       (scattered 8-byte global stores cost ~5 clk each per SM; runs of a tile are
//       contiguous).  Row order x tile order x (warp, step, lane) order: deterministic.
//   4 sync_pair_kernel      a CTA takes pieces of <= 4096 sorted entries of one bucket:
//       16 coalesced loads per thread, CTA-wide counting sort by sub-bucket in shared
//       memory ((thread, entry) order, no atomics), then warp w streams run w through
//       its lane groups with the roles swapped — a lane owns every 32nd particle of
//       the run and keeps the partial sums of the group's 32 bins in registers, a
//       butterfly transpose-reduction leaves bin j's total in lane j — and adds
//       ds_q * sum (fp64) into the CTA's row of partial sums with RED.ADD.F64
//       (a slot is only ever touched by one warp of the CTA: fixed order).
//   5 pair_moments_kernel / pair_final_kernel   fp64 sub-bucket moments (suffix sums)
//       from the runs' moment lanes; CTA partials + the linear part
//       v_q S0 + s_q (fa_j S0 + S1) + the hinge of the sub-buckets beyond the
//       threshold  ->  one value per bin.
// Hinge sums are float per run (<= 4096 / 32 terms per lane, then a 32-lane tree), folded
// into fp64 per run; every reduction runs in a fixed order: bitwise reproducible.
//
// Requires a table with F = 0 at both ends (true for sync::TabulateFfunc, whose
// first node is forced to 0 and whose nodes beyond x = 20 are 0); any other table,
// very wide bin ranges and the FromDist form use the gather kernel.
//
// Roofline: the pipeline is HBM-bound (64 B per particle over the three streaming
// kernels + 8 B read by the pair kernel); the pair loop issues about particles x 32 lanes
// x groups-per-sub-bucket evaluations (2 FP32-pipe instructions each) instead of
// particles x bins.
//
// Compiled with -fmad=false: every FMA below is an explicit fmaf()/fma().
#include "rgc_internal.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

namespace rgc {

  constexpr int kPThreads    = 256;
  constexpr int kPWarps      = kPThreads / 32;
  constexpr int kPMaxGPW     = 8;    // lane groups a warp evaluates at once (one chunk)
  constexpr int kSubDiv      = 8;    // a table cell is cut into eighths of the fraction fc ...
  constexpr int kSub         = kSubDiv + 1; // ... shifted by the plan's phase: 9 sub-buckets, s = floor(8 fc + phi)
  constexpr int kMomStride   = 2 * kSub + 2; // floats per piece in piece_mom (16-byte multiple)
  constexpr int kSubPk       = (kSub + 1) / 2; // sub-bucket counters packed two to a register
  constexpr int kPMaxBins    = 2032; // per launch
  constexpr int kPMaxGroups  = 96;   // lane groups per launch (bins by sub-bucket + moment lanes)
  constexpr int kPMaxBuckets = 1024;
  constexpr unsigned kInvalidKey = 0xffffu;
  constexpr int kPTile       = 4096; // particles per tile of the prologue / sort kernels
  constexpr int kPSteps      = kPTile / kPThreads;
  constexpr int kPieceLen    = 4096; // sorted entries per work unit of the pair kernel
  // the pair kernel's CTA: 8 run warps + two sort groups of 4 warps
  constexpr int kPairRunThreads = 256;
  constexpr int kPairSortGroup  = 128;
  constexpr int kPairSortWarps  = kPairSortGroup / 32;
  constexpr int kPairThreads    = kPairRunThreads + 2 * kPairSortGroup;
  constexpr int kPairSortEnt    = kPieceLen / kPairSortGroup; // entries per sort thread
  constexpr int kPairBufLen     = kPieceLen + 64 * kSub; // a sorted piece, every run padded to 64 entries
  constexpr int kPairBufs       = 4; // sorted pieces in flight: each sort group works up to two pieces ahead

  // fp64 constants of the prologue, read as constant-bank operands (an immediate double
  // whose low word is not zero costs two UMOVs per use otherwise)
  struct PairConsts {
    double l3, lm4, lm2, inv_ln2, split24, magic, flt_max, ep_lo, ep_hi, q_lo;
  };

  struct PairParams {
    PairConsts     kc;
    const double2* log_tab; // [256] {log2(1 + j/256), 1 / (1 + j/256)}
    const float*   u[3];
    const float* e[3];
    const float* b[3];
    std::size_t  nprtl;
    const int2*   slot_i;  // per slot {Aoff, -}: cell q = max(Aoff + bucket, 0); spare slots Aoff << 0
    const float2* slot_f;  // per slot {fa0, -}
    const float4* coef_dh; // per padded cell {hinge coefficient, hinge position, sign, 0}
    const int4*   chunks;  // {first lane group, groups (<= kPMaxGPW), carries the moment lanes, -}
    const unsigned char* na_tab; // [bucket][chunk] leading groups of the chunk that are on the table
    const unsigned*      extmask; // [bucket] bit s: sub-bucket s also sees run s - 1; bit kSub + s: run s + 1
    int chunk_first[kSub + 1]; // chunks of sub-bucket s: [chunk_first[s], chunk_first[s + 1])
    int n_pad, nb, nbp, nchunks;
    int kmin;              // bucket = floor(c) - kmin
    double kmin_d;
    double inv_B0;           // 1 / B0
    double e_scale;          // e_syn_at_g_syn / (g_syn * g_syn), the float product promoted
    double cells_per_octave; // log10(2) / dL
    double c0, inv_dL, c_lo, c_hi;
    // staged per-particle results of the prologue kernel (particle order)
    float2*         cw;   // (fc, w)
    unsigned short* keys; // bucket, kInvalidKey = not on the table
    // rows: row r owns tiles [r * tiles_per_row, (r + 1) * tiles_per_row) of kPTile particles
    int         rows, ntiles, tiles_per_row, tiles_per_cta1;
    int*        counts; // [nbp][rows]: per-row bucket counts, then offsets inside the bucket
    int*        tot;    // [nbp] valid particles per bucket
    float2*     sorted; // (fc, w) in global bucket order, every bucket padded to an even length
    float*      piece_mom; // [piece][kMomStride]: {S0, S1} of the piece's kSub sub-bucket runs (float sums)
    float       sub_phi;   // phase of the sub-bucket boundaries: s = floor(kSubDiv fc + sub_phi)
    int*        poison;    // != 0: a particle's chiR overflows float (see pair_prologue)
    unsigned long long* lane_evals; // hinge evaluations the pair kernel issued (roofline accounting)
    unsigned force_groups;          // != 0: lane groups are evaluated even when all their lanes sit on zero cells
    double*     partials;  // [cta][nslots] hinge sums (RED.ADD.F64 targets, zeroed per pass; a slot is
                           // only ever touched by one warp of the CTA: the order of additions is fixed)
    int         nslots;
    // shared-memory layout of the pair kernel (byte offsets, computed once on the host)
    int o_coef, o_bstart, o_pstart, o_tmp, o_slot, o_chunk, o_sorted, o_cur, o_wtot, o_run;
  };

  __host__ __device__ inline std::size_t pair_align16(std::size_t x) { return (x + 15) & ~std::size_t(15); }

  struct PairSmem {
    std::size_t coef, bstart, pstart, tmp, slot, chunk, sorted, cur, wtot, run, total;
  };
  // shared memory of the pair kernel: plan tables, the piece sorted by sub-bucket (every run
  // padded to a multiple of 16 entries), the per-thread cursors of the sort, packed warp totals,
  // the run table
  __host__ __device__ inline PairSmem pair_smem_layout(int n_pad, int nbp, int nslots, int nchunks) {
    PairSmem L;
    std::size_t o = 0;
    L.coef = o;    o = pair_align16(o + (std::size_t)n_pad * sizeof(float4));
    L.bstart = o;  o = pair_align16(o + (std::size_t)(nbp + 2) * sizeof(int));
    L.pstart = o;  o = pair_align16(o + (std::size_t)(nbp + 2) * sizeof(int));
    L.tmp = o;     o = pair_align16(o + (std::size_t)(2 * (kPairThreads / 32)) * sizeof(int));
    L.slot = o;    o = pair_align16(o + (std::size_t)nslots * sizeof(int2));
    L.chunk = o;   o = pair_align16(o + (std::size_t)nchunks * sizeof(int4));
    L.wtot = o;    o = pair_align16(o + (std::size_t)2 * 2 * kPairSortWarps * kSubPk * sizeof(unsigned));
    L.run = o;     o = pair_align16(o + (std::size_t)kPairBufs * kSub * sizeof(int2));
    L.cur = o;     o = pair_align16(o + (std::size_t)2 * kSub * kPairSortGroup * sizeof(int));
    o = (o + 127) & ~std::size_t(127);
    L.sorted = o;  o = o + (std::size_t)kPairBufs * kPairBufLen * sizeof(float2);
    L.total  = o;
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

  // log2 of a positive normal double that carries float precision (e_peak after its
  // rounding: 24 significant bits), |error| < 1e-12: exponent + table over the top 8
  // fraction bits, log2(m) = T[j] + log2(1 + r) with m = mh (1 + r), mh = 1 + j/256,
  // r = (m - mh) / mh in [0, 2^-8) and a cubic in r (r^5 / 5 < 2e-13).  tab[j] = {log2(mh),
  // 1 / mh}, staged in shared memory.  8 fp64 operations against 16 for the atanh series.
  __device__ __forceinline__ double log2_f24(const PairConsts& K, const double2* __restrict__ tab,
                                             double x) {
    const int     hi = __double2hiint(x);
    const int     ex = ((hi >> 20) & 0x7ff) - 1023;
    const int     j  = (hi >> 12) & 0xff; // top 8 fraction bits
    const double  m  = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
    const double  mh = __hiloint2double((hi & 0x000ff000) | 0x3ff00000, 0);
    const double2 t  = tab[j];
    const double  r  = (m - mh) * t.y;
    double        p  = fma(r, K.lm4, K.l3); // 1/3 - r/4
    p                = fma(r, p, K.lm2);     // -1/2 + r (...)
    p                = fma(r, p, 1.0);
    return (double)ex + fma(r * p, K.inv_ln2, t.x);
  }

  // x rounded to float precision (24 significant bits, nearest) without leaving the
  // fp64 pipe: Veltkamp's split with 2^29 + 1.  Equal to (double)(float)x for normal
  // floats; saves the two XU-pipe conversions (16 lanes/clk/SM against 64 for fp64).
  __device__ __forceinline__ double round24(const PairConsts& K, double x) {
    const double t = x * K.split24; // 2^29 + 1
    return t - (t - x);
  }

  // reference src/physics/synchrotron.hpp:193-231 — gamma, beta, beta.E, beta x B,
  // chiR, e_peak — in fp64 like the reference's promoted arithmetic, then the table
  // coordinate of e_peak split into bucket and fraction.  The float roundings of the
  // reference's sequence are kept (squares of U, chiR, e_peak); sqrt, quotients and
  // log10 go through rsqrt_nr / log2_f24 (error < 1e-12: far below a float ulp).
  // Returns true with (bucket, fc, w), false for a particle the reference skips.  A
  // particle that poisons the reference's result raises *P.poison (rare path): chiR =
  // real_t(sqrt(q) / B0) = +inf makes e_peak = +inf > 0, x0 = e_syn / e_peak = 0 < xmin,
  // F = yfill = 0 and the term e_syn * inf * 0 = NaN in EVERY photon bin
  // (synchrotron.hpp:162-171).
  __device__ __forceinline__ bool pair_prologue(const PairParams& P, const double2* __restrict__ ltab,
                                               float ux, float uy, float uz, float ex, float ey,
                                               float ez, float bx, float by, float bz,
                                               unsigned& bucket, float& fc, float& w) {
    const double dux = (double)ux, duy = (double)uy, duz = (double)uz;
    const double dex = (double)ex, dey = (double)ey, dez = (double)ez;
    const double dbx = (double)bx, dby = (double)by, dbz = (double)bz;
    // gamma^2 from the float squares, like `1.0 + ux * ux + uy * uy + uz * uz` in the
    // reference (each square rounded to float, then promoted)
    const PairConsts& K = P.kc;
    const double g2 = ((1.0 + (double)(ux * ux)) + (double)(uy * uy)) + (double)(uz * uz);
    const double rg  = rsqrt_nr(g2);
    const double beta_x = dux * rg, beta_y = duy * rg, beta_z = duz * rg;
    const double bde = fma(beta_z, dez, fma(beta_y, dey, beta_x * dex));
    const double sx  = dex + fma(beta_y, dbz, -(beta_z * dby));
    const double sy  = dey + fma(beta_z, dbx, -(beta_x * dbz));
    const double sz  = dez + fma(beta_x, dby, -(beta_y * dbx));
    const double q   = fma(-bde, bde, fma(sz, sz, fma(sy, sy, sx * sx)));
    // q <= 0: chiR = 0 or NaN in the reference, the pair is skipped (synchrotron.hpp:162);
    // NaN / inf inputs fail this or the range checks below
    if (!(q > K.q_lo)) {
      return false;
    }
    // (q = +inf gives chi = NaN here: 1/sqrt(inf) = 0, inf * 0)
    const double chi = (q * rsqrt_nr(q)) * P.inv_B0;
    if (!(chi <= K.flt_max)) { // chiR rounds to +inf as a float, or q is huge / inf
      const double chi_big = sqrt(q) * P.inv_B0;
      if (chi_big > K.flt_max && P.e_scale * g2 > 0.0) {
        atomicAdd(P.poison, 1);
      }
      return false;
    }
    // chiR and e_peak are rounded to float exactly where the reference rounds them
    // (synchrotron.hpp:230-231): a mono-energetic population has no other particles to
    // average a half-ulp coordinate shift away
    const float  chi_f = (float)chi; // the weight; off the critical path
    const double ep_d  = (P.e_scale * g2) * round24(K, chi);
    // float(e_peak) must be a positive finite float (else x0 = e_syn / e_peak is
    // off the table on either side)
    if (!(ep_d > K.ep_lo && ep_d < K.ep_hi)) {
      return false;
    }
    const double ep = round24(K, ep_d);
    const double c  = fma(-log2_f24(K, ltab, ep), P.cells_per_octave, P.c0);
    if (!(c >= P.c_lo && c < P.c_hi)) {
      return false;
    }
    // floor(c) without FRND / F2I: round-to-nearest of c - 1/2 through the 2^52 trick;
    // an exact integer c may land on either neighbour, (K, fc = 0) and (K - 1, fc = 1)
    // being the same table coordinate
    const double magic = K.magic; // 1.5 * 2^52
    const double rm    = (c - 0.5) + magic;
    int          ri    = __double2loint(rm);
    double       fl    = rm - magic;
    if (ri < P.kmin) {
      ri = P.kmin;
      fl = P.kmin_d;
    }
    bucket = (unsigned)(ri - P.kmin);
    fc     = (float)(c - fl);
    w      = chi_f;
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

  // Lanes holding the same key, from one ballot per key bit (MATCH.ANY takes several
  // hundred cycles when most of a warp's keys differ, which is the normal case here).
  // Valid keys are < 1024 (kPMaxBuckets); kInvalidKey has all low bits set, so 11 bits
  // always separate it from every bucket.
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

  // ---- probe behind the default ranking of the sort kernel.  sync_sort_kernel<true> takes a
  // particle's rank among its warp's particles of the same bucket from the return value of one
  // shared-memory atomicAdd on a warp-private cursor.  The CUDA programming model does not say
  // in which order the lanes of ONE instruction that hit the same address are served; the
  // result is reproducible only if that order is a fixed function of the lane ids.  This
  // kernel checks exactly that on the device in use (8 warps with private cursors and random
  // keys, as in the sort kernel, many rounds): ranks of same-key lanes must ascend with the
  // lane id.  One violation and the library ranks by ballots for the rest of the process
  // (order fixed by construction, ~0.3 ms per 1e8 particles slower).
  __global__ void __launch_bounds__(kPThreads)
    rank_order_probe_kernel(int rounds, unsigned seed, int* __restrict__ violations) {
    __shared__ int cur[kPWarps][64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned  x    = seed ^ (blockIdx.x * 2654435761u) ^ (threadIdx.x * 40503u);
    int       bad  = 0;
    for (int r = 0; r < rounds; ++r) {
      cur[warp][lane]      = 0;
      cur[warp][lane + 32] = 0;
      __syncwarp();
      x = x * 1664525u + 1013904223u;
      // few distinct keys: many same-address lanes per instruction
      const unsigned key = (x >> 24) % (1u + (unsigned)(r % 24));
      const int      rk  = atomicAdd(&cur[warp][key], 1);
      const unsigned same = __match_any_sync(0xffffffffu, key);
      const int      want = __popc(same & ((1u << lane) - 1u));
      bad += rk != want;
      __syncwarp();
    }
    if (bad) {
      atomicAdd(violations, bad);
    }
  }

  // Exclusive scans over the buckets, by one CTA of kPThreads threads:
  //   bstart[b] = first sorted entry of bucket b (every bucket padded to an even length)
  //   pstart[b] = first piece of bucket b (pieces of kPieceLen entries, the last one short)
  // entries [0, nb]; tmp: 2 * kPWarps ints.  Ends with a __syncthreads().
  template <int NT = kPThreads>
  __device__ __forceinline__ void block_scan_buckets(const int* __restrict__ tot, int nb,
                                                     int* bstart, int* pstart, int* tmp) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kPer   = kPMaxBuckets / NT; // consecutive buckets per thread
    constexpr int kWarps = NT / 32;
    int pe[kPer], pp[kPer];
    int esum = 0, psum = 0;
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      const int b   = tid * kPer + i;
      const int t   = b < nb ? tot[b] : 0;
      const int pad = (t + 1) & ~1;
      pe[i]         = pad;
      pp[i]         = (pad + kPieceLen - 1) / kPieceLen;
      esum += pad;
      psum += pp[i];
    }
    int ei = esum, pi = psum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int te = __shfl_up_sync(0xffffffffu, ei, off);
      const int tp = __shfl_up_sync(0xffffffffu, pi, off);
      if (lane >= off) {
        ei += te;
        pi += tp;
      }
    }
    if (lane == 31) {
      tmp[warp]           = ei;
      tmp[kWarps + warp] = pi;
    }
    __syncthreads();
    int ebase = 0, pbase = 0;
#pragma unroll
    for (int wq = 0; wq < kWarps; ++wq) {
      ebase += wq < warp ? tmp[wq] : 0;
      pbase += wq < warp ? tmp[kWarps + wq] : 0;
    }
    int erun = ebase + ei - esum, prun = pbase + pi - psum;
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      const int b = tid * kPer + i;
      if (b <= nb) {
        bstart[b] = erun;
        pstart[b] = prun;
      }
      erun += pe[i];
      prun += pp[i];
    }
    __syncthreads();
  }

  // ---- kernel 1: per-particle prologue, streamed once over the particle columns
  // (36 B read, 10 B written per particle; HBM-bound), plus the bucket counts of
  // every row.  A CTA owns a contiguous range of tiles; its shared-memory histogram
  // is added to the row's global counts (integer atomics: order-independent) whenever
  // the row changes.  Entries past nprtl up to the end of the last tile are written
  // as invalid so the sort kernel copies whole tiles.
  template <int MINB>
  __global__ void __launch_bounds__(kPThreads, MINB)
    sync_prologue_kernel(const __grid_constant__ PairParams P) {
    __shared__ int     hist[kPMaxBuckets];
    __shared__ double2 ltab[256];
    const int          tid = threadIdx.x;
    for (int i = tid; i < P.nbp; i += kPThreads) {
      hist[i] = 0;
    }
    ltab[tid] = P.log_tab[tid]; // kPThreads == 256
    __syncthreads();
    const int t0 = blockIdx.x * P.tiles_per_cta1;
    const int t1 = min(t0 + P.tiles_per_cta1, P.ntiles);
    auto flush = [&](int row) {
      __syncthreads();
      for (int b = tid; b < P.nb; b += kPThreads) {
        const int v = hist[b];
        if (v) {
          atomicAdd(&P.counts[(std::size_t)b * P.rows + row], v);
          hist[b] = 0;
        }
      }
      __syncthreads();
    };
    for (int tile = t0; tile < t1; ++tile) {
      const std::size_t base = (std::size_t)tile * kPTile;
#pragma unroll 1
      for (int step = 0; step < kPTile / (kPThreads * 4); ++step) {
        const std::size_t i0 = base + (std::size_t)step * (kPThreads * 4) + (std::size_t)tid * 4;
        float4            v[9];
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
            ok = pair_prologue(P, ltab, f[0 * 4 + k], f[1 * 4 + k], f[2 * 4 + k], f[3 * 4 + k],
                               f[4 * 4 + k], f[5 * 4 + k], f[6 * 4 + k], f[7 * 4 + k],
                               f[8 * 4 + k], bucket, fc, w);
          }
          out_k[k]          = (unsigned short)(ok ? bucket : kInvalidKey);
          out_cw[2 * k]     = ok ? fc : 0.0f;
          out_cw[2 * k + 1] = ok ? w : 0.0f;
          if (ok) {
            atomicAdd(&hist[bucket], 1);
          }
        }
        float4* cw4 = reinterpret_cast<float4*>(P.cw + i0);
        cw4[0]      = make_float4(out_cw[0], out_cw[1], out_cw[2], out_cw[3]);
        cw4[1]      = make_float4(out_cw[4], out_cw[5], out_cw[6], out_cw[7]);
        *reinterpret_cast<uint2*>(P.keys + i0) =
          make_uint2((unsigned)out_k[0] | ((unsigned)out_k[1] << 16),
                     (unsigned)out_k[2] | ((unsigned)out_k[3] << 16));
      }
      if (tile + 1 == t1 || (tile + 1) / P.tiles_per_row != tile / P.tiles_per_row) {
        flush(tile / P.tiles_per_row);
      }
    }
  }

  // ---- kernel 2: per bucket (one warp each), exclusive scan of the row counts in
  // row order -> offset of every row's first particle inside the bucket; bucket totals
  __global__ void __launch_bounds__(kPThreads)
    pair_colscan_kernel(int* __restrict__ counts, int rows, int nb, int* __restrict__ tot) {
    const int lane = threadIdx.x & 31;
    const int b    = blockIdx.x * kPWarps + (threadIdx.x >> 5);
    if (b >= nb) {
      return;
    }
    int* col = counts + (std::size_t)b * rows;
    int  run = 0;
#pragma unroll 4
    for (int r0 = 0; r0 < rows; r0 += 32) {
      const int r    = r0 + lane;
      const int v    = r < rows ? col[r] : 0;
      int       incl = v;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) {
          incl += t;
        }
      }
      if (r < rows) {
        col[r] = run + incl - v;
      }
      run += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) {
      tot[b] = run;
    }
  }

  struct SortSmem {
    std::size_t bstart, pstart, tmp, gcur, seg_off, hw, stage_cw, stage_k, sorted, sorted_k, mbar, total;
  };

  __host__ __device__ inline SortSmem sort_smem_layout(int nbp) {
    SortSmem    L;
    std::size_t o = 0;
    L.bstart = o;   o = pair_align16(o + (std::size_t)(nbp + 2) * sizeof(int));
    L.pstart = o;   o = pair_align16(o + (std::size_t)(nbp + 2) * sizeof(int));
    L.tmp = o;      o = pair_align16(o + (std::size_t)(2 * kPWarps + 2) * sizeof(int));
    L.gcur = o;     o = pair_align16(o + (std::size_t)nbp * sizeof(int));
    L.seg_off = o;  o = pair_align16(o + (std::size_t)(nbp + 2) * sizeof(int));
    L.hw = o;       o = pair_align16(o + (std::size_t)kPWarps * nbp * sizeof(int));
    o = (o + 127) & ~std::size_t(127);
    L.stage_cw = o; o = pair_align16(o + (std::size_t)kPTile * sizeof(float2));
    L.stage_k = o;  o = pair_align16(o + (std::size_t)kPTile * sizeof(unsigned short));
    L.sorted = o;   o = pair_align16(o + (std::size_t)kPTile * sizeof(float2));
    L.sorted_k = o; o = pair_align16(o + (std::size_t)kPTile * sizeof(unsigned short));
    L.mbar = o;     o = pair_align16(o + 16);
    L.total = o;
    return L;
  }

  // ---- kernel 3: bucket sort of every tile inside shared memory, coalesced write-out
  // of the tile's bucket runs into the global bucket order.  One CTA per row.
  // ATOMIC_RANK: a particle's rank among its warp's particles of the same bucket is
  // the return value of one shared-memory atomic on the warp-private cursor (lanes of
  // one instruction that hit the same cursor are replayed by the hardware in a fixed
  // order, so the result is reproducible in practice — tests/test_gpu_parity.py
  // test_repeatable checks it on the device); otherwise 11 ballots per step group the
  // lanes by key and the group's first lane advances the cursor (order guaranteed by
  // construction; RGC_SORT_RANK=ballot selects it).
  template <bool ATOMIC_RANK>
  __global__ void __launch_bounds__(kPThreads, 2)
    sync_sort_kernel(const __grid_constant__ PairParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const SortSmem  L        = sort_smem_layout(P.nbp);
    int*            bstart   = reinterpret_cast<int*>(smem_raw + L.bstart);
    int*            pstart   = reinterpret_cast<int*>(smem_raw + L.pstart);
    int*            scan_tmp = reinterpret_cast<int*>(smem_raw + L.tmp);
    int*            gcur     = reinterpret_cast<int*>(smem_raw + L.gcur);
    int*            seg_off  = reinterpret_cast<int*>(smem_raw + L.seg_off);
    int*            hw       = reinterpret_cast<int*>(smem_raw + L.hw);
    float2*         stage_cw = reinterpret_cast<float2*>(smem_raw + L.stage_cw);
    unsigned short* stage_k  = reinterpret_cast<unsigned short*>(smem_raw + L.stage_k);
    float2*         sorted   = reinterpret_cast<float2*>(smem_raw + L.sorted);
    unsigned short* sorted_k = reinterpret_cast<unsigned short*>(smem_raw + L.sorted_k);
    void*           mbar     = smem_raw + L.mbar;

    const int tid  = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int nb   = P.nb;
    const int nbp  = P.nbp;
    const int row  = blockIdx.x;
    const int t0   = row * P.tiles_per_row;
    const int t1   = min(t0 + P.tiles_per_row, P.ntiles);

    constexpr unsigned kStageBytesCW = kPTile * sizeof(float2);
    constexpr unsigned kStageBytesK  = kPTile * sizeof(unsigned short);
    auto issue_tile_copy = [&](int tile) {
      mbar_expect_tx(mbar, kStageBytesCW + kStageBytesK);
      bulk_g2s(stage_cw, P.cw + (std::size_t)tile * kPTile, kStageBytesCW, mbar);
      bulk_g2s(stage_k, P.keys + (std::size_t)tile * kPTile, kStageBytesK, mbar);
    };
    if (tid == 0) {
      mbar_init(mbar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    block_scan_buckets(P.tot, nb, bstart, pstart, scan_tmp); // ends with __syncthreads()
    if (tid == 0 && t0 < t1) {
      issue_tile_copy(t0);
    }
    if (row == 0) {
      for (int b = tid; b < nb; b += kPThreads) {
        const int t = P.tot[b];
        if (t & 1) {
          P.sorted[bstart[b] + t] = make_float2(0.0f, 0.0f); // zero-weight pad of the bucket
        }
      }
    }
    for (int b = tid; b < nb; b += kPThreads) {
      gcur[b] = bstart[b] + P.counts[(std::size_t)b * P.rows + row];
    }

    constexpr int kPer = kPMaxBuckets / kPThreads; // most buckets per thread in the scan
    const int     bpt    = (nb + kPThreads - 1) / kPThreads;
    unsigned      parity = 0;
    for (int tile = t0; tile < t1; ++tile) {
      // ---- per-warp bucket cursors, zeroed; the previous tile's write-out is
      // complete once every warp has passed the barrier below
      for (int i = tid; i < kPWarps * nbp; i += kPThreads) {
        hw[i] = 0;
      }
      __syncthreads();
      mbar_wait(mbar, parity);
      parity ^= 1u;
      // ---- pass A: rank of every particle among its warp's particles of the same
      // bucket (warp w owns tile entries [512 w, 512 w + 512), 32 per step)
      int*     cur = hw + warp * nbp;
      unsigned rank_pack[kPSteps / 2]; // u16 ranks, two per register
      unsigned key_pack[kPSteps / 2];  // the keys of pass A, kept for pass B
      if (ATOMIC_RANK) {
#pragma unroll
        for (int st = 0; st < kPSteps; ++st) {
          const unsigned key = stage_k[warp * (kPTile / kPWarps) + st * 32 + lane];
          unsigned       rk  = 0;
          if (key != kInvalidKey) {
            rk = (unsigned)atomicAdd(&cur[key], 1);
          }
          if ((st & 1) == 0) {
            rank_pack[st >> 1] = rk;
            key_pack[st >> 1]  = key;
          } else {
            rank_pack[st >> 1] |= rk << 16;
            key_pack[st >> 1] |= key << 16;
          }
        }
      } else {
        // The lanes of a step are grouped by key (match_key11); the group's first lane
        // advances the warp's cursor of that bucket.
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
              cur[key[j]] = basev + __popc(m[j]);
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
      }
      __syncthreads();
      // ---- scan: tile totals per bucket -> first sorted entry of every bucket
      // (seg_off) and of every (warp, bucket)
      // (bpt consecutive buckets per thread: 1 for up to 256 buckets, so the scan is
      // spread over all warps)
      int tcnt[kPer];
      {
        int sum = 0;
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
          const int b   = tid * bpt + i;
          int       tot = 0;
          if (i < bpt && b < nb) {
#pragma unroll
            for (int wq = 0; wq < kPWarps; ++wq) {
              tot += hw[wq * nbp + b];
            }
          }
          tcnt[i] = tot;
          sum += tot;
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
        int run = warp_base + incl - sum;
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
          const int b = tid * bpt + i;
          if (i < bpt && b < nb) {
            // global position of sorted entry i of bucket b: i + (gcur[b] - run); the
            // row's cursor moves on by this tile's count right away
            seg_off[b] = gcur[b] - run;
            gcur[b] += tcnt[i];
            int r2 = run;
#pragma unroll
            for (int wq = 0; wq < kPWarps; ++wq) {
              const int c      = hw[wq * nbp + b];
              hw[wq * nbp + b] = r2;
              r2 += c;
            }
            run = r2;
          }
        }
        if (tid == 0) {
          scan_tmp[2 * kPWarps] = total;
        }
      }
      __syncthreads();
      // ---- pass B: scatter to bucket order (cursor base of (warp, bucket) + rank)
      {
        const int* basep = hw + warp * nbp;
#pragma unroll
        for (int step = 0; step < kPSteps; ++step) {
          const int      idx = warp * (kPTile / kPWarps) + step * 32 + lane;
          const unsigned key = ATOMIC_RANK ? (key_pack[step >> 1] >> ((step & 1) * 16)) & 0xffffu
                                           : (unsigned)stage_k[idx];
          if (key != kInvalidKey) {
            const unsigned rk  = (rank_pack[step >> 1] >> ((step & 1) * 16)) & 0xffffu;
            const unsigned pos = (unsigned)basep[key] + rk;
            sorted[pos]        = stage_cw[idx];
            sorted_k[pos]      = (unsigned short)key;
          }
        }
      }
      __syncthreads();
      // staging buffers are free again: fetch the next tile underneath the write-out
      if (tid == 0 && tile + 1 < t1) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue_tile_copy(tile + 1);
      }
      // ---- write-out: consecutive threads write consecutive sorted entries; a bucket's
      // run of this tile continues where the row's previous tiles left the bucket
      {
        const int total = scan_tmp[2 * kPWarps];
        for (int i = tid; i < total; i += kPThreads) {
          P.sorted[i + seg_off[sorted_k[i]]] = sorted[i];
        }
      }
      // no barrier here: the next tile's first barrier (after the cursor reset, which
      // touches nothing the write-out reads) orders everything that follows
    }
  }

  // ---- kernel 4: the pair loop over the globally bucket-sorted (fc, w).
  //
  // A piece is <= kPieceLen sorted entries of one bucket.  The hinge r = max(0, fc + fa') of
  // bin j is identically 0 for every particle with fc <= -fa' and linear in fc for every
  // particle above, so only particles whose fc lies in the same eighth of the cell as the
  // bin's threshold -fa' need per-pair work; the others enter through the moments (S0, S1)
  // of their sub-bucket, which the first two lanes of every sub-bucket's first lane group
  // deliver for free (pair_final_kernel adds that part).
  //
  // One CTA of 16 warps per SM, warp-specialised:
  //   * warps 8-15 are two SORT groups of four warps.  Group g takes every other piece of the
  //     CTA into one of four buffers: 16 coalesced 16-byte loads (32 entries) per thread,
  //     counting sort by sub-bucket s = floor(8 fc + phi) in (thread, entry) order (6-bit
  //     per-thread counters of the 9 sub-buckets packed into one 64-bit register, warp scans of
  //     the counters packed two to a register, warp totals through shared memory, private
  //     cursor columns for the scatter; no atomics), every run padded to a multiple of 64
  //     entries with zero-weight entries.
  //   * warps 0-7 are the RUN warps: warp w streams run w of the piece (warp 0 also run 8: with
  //     the phase shift runs 0 and 8 are the two parts of one eighth) through the lane groups
  //     that hold the bins whose threshold lies in that sub-bucket (plus, rarely, a
  //     neighbouring sub-bucket's groups, `extmask`), and adds ds_q * sum (fp64) into the CTA's
  //     row of partial sums with RED.ADD.F64 (a slot is only ever touched by one warp of the
  //     CTA: the order of additions is fixed).  Pieces are dealt to the CTAs in consecutive
  //     pairs; a pair of the same bucket is streamed as one (one lane-group set-up and one
  //     transpose-reduction for both) — decided by the piece indices alone, never by timing.
  // The hand-over is a pair of named barriers per buffer (FULL: 128 sort threads arrive, 256
  // run threads wait; EMPTY: the other way round); with four buffers both sort groups work up to
  // two of their pieces ahead of the run warps.
  // A cell pair next to the table's zero tail (sign < 0) needs sum w max(0, 1 - u): its
  // lanes run the same two instructions on sat(fc + fa0) and the run's S0 lane gives
  // S0 - sum w sat(u), which is exactly 0 when every particle of the run is beyond the
  // tail (both lanes then execute the identical float sequence and reduction tree).
  __device__ __forceinline__ void named_bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
  }
  __device__ __forceinline__ void named_bar_arrive(int id, int count) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
  }

  __global__ void __launch_bounds__(kPairThreads, 1)
    sync_pair_kernel(const __grid_constant__ PairParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* coef     = reinterpret_cast<float4*>(smem_raw + P.o_coef);
    int*    bstart   = reinterpret_cast<int*>(smem_raw + P.o_bstart);
    int*    pstart   = reinterpret_cast<int*>(smem_raw + P.o_pstart);
    int*    tmp      = reinterpret_cast<int*>(smem_raw + P.o_tmp);
    int2*   slot_tab = reinterpret_cast<int2*>(smem_raw + P.o_slot);
    int4*   chunks   = reinterpret_cast<int4*>(smem_raw + P.o_chunk);

    const int tid  = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int nb   = P.nb;
    for (int i = tid; i < P.n_pad; i += kPairThreads) {
      coef[i] = P.coef_dh[i];
    }
    for (int i = tid; i < P.nslots; i += kPairThreads) {
      slot_tab[i] = make_int2(P.slot_i[i].x, __float_as_int(P.slot_f[i].x));
    }
    for (int i = tid; i < P.nchunks; i += kPairThreads) {
      chunks[i] = P.chunks[i];
    }
    block_scan_buckets<kPairThreads>(P.tot, nb, bstart, pstart, tmp); // ends with __syncthreads()

    // pieces are dealt to the CTAs in consecutive pairs (two consecutive pieces mostly share their
    // bucket; buckets differ in cost, so the pairs of a CTA are spread over all of them): the
    // CTA's k-th piece is piece_of(k)
    const int npieces   = pstart[nb];
    const int npairs    = (npieces + 1) / 2;
    const int my_pairs  = (int)blockIdx.x < npairs ? (npairs - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int last_pair = (int)blockIdx.x + (my_pairs - 1) * (int)gridDim.x;
    // (the globally last pair may hold a single piece)
    const int my_pieces = my_pairs == 0 ? 0 : 2 * my_pairs - ((2 * last_pair + 1 >= npieces) ? 1 : 0);
    auto piece_of = [&](int k) { return 2 * ((int)blockIdx.x + (k >> 1) * (int)gridDim.x) + (k & 1); };
    constexpr int kBarGroup = 1, kBarFull = 3, kBarEmpty = 3 + kPairBufs; // named barriers (+ group / buffer)
    constexpr int kHandover = kPairSortGroup + kPairRunThreads;

    if (warp >= kPairRunThreads / 32) {
      // =============================================================== sort groups
      const int g     = (warp - kPairRunThreads / 32) / (kPairSortGroup / 32);
      const int pt    = tid - kPairRunThreads - g * kPairSortGroup; // thread of the group
      const int pwarp = pt >> 5;                                    // warp of the group
      int*      cur    = reinterpret_cast<int*>(smem_raw + P.o_cur) + g * (kSub * kPairSortGroup);
      unsigned* wtot   = reinterpret_cast<unsigned*>(smem_raw + P.o_wtot) + g * (2 * kPairSortWarps * kSubPk);
      // sub-bucket of a fraction fc in [0, 1]: floor(8 fc + phi) as the round-to-nearest of
      // 8 fc + (phi - 1/2), taken from the mantissa after adding 1.5 * 2^23 (no F2I on the XU
      // pipe; phi < 1 keeps the result <= 8; a tie lies within the plan's eps of a boundary,
      // where either side is valid)
      const float sub_bias = P.sub_phi - 0.5f;
      auto sub_of = [&](float fc) {
        return min(__float_as_int(fmaf(fc, (float)kSubDiv, sub_bias) + 12582912.0f) & 15, kSub - 1);
      };
      int bcur = 0; // the group's pieces ascend: the bucket search resumes where it stopped
      for (int k = g; k < my_pieces; k += 2) {
        const int piece = piece_of(k);
        const int buf   = k % kPairBufs; // the group's buffers alternate: it works up to two pieces ahead
        float2*   B      = reinterpret_cast<float2*>(smem_raw + P.o_sorted) + (std::size_t)buf * kPairBufLen;
        int2*     runtab = reinterpret_cast<int2*>(smem_raw + P.o_run) + buf * kSub;
        while (pstart[bcur + 1] <= piece) {
          ++bcur;
        }
        const int beg = bstart[bcur] + (piece - pstart[bcur]) * kPieceLen;
        const int n   = min(kPieceLen, bstart[bcur + 1] - beg); // even
        // ---- the piece: 16 coalesced 16-byte loads (two entries) per thread, all in flight at
        // once (entries beyond a short piece become zero-weight pads; beg and n are even)
        float2        ent[kPairSortEnt];
        const float4* src = reinterpret_cast<const float4*>(P.sorted + beg);
#pragma unroll
        for (int st = 0; st < kPairSortEnt / 2; ++st) {
          const int    e2 = st * kPairSortGroup + pt;
          const float4 v  = 2 * e2 < n ? __ldcs(src + e2) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
          ent[2 * st]     = make_float2(v.x, v.y);
          ent[2 * st + 1] = make_float2(v.z, v.w);
        }
        unsigned long long cnt = 0ull;
#pragma unroll
        for (int st = 0; st < kPairSortEnt; ++st) {
          cnt += 1ull << (6 * sub_of(ent[st].x));
        }
        unsigned pk[kSubPk], incl[kSubPk];
#pragma unroll
        for (int j = 0; j < kSubPk; ++j) {
          const unsigned lo = (unsigned)((cnt >> (12 * j)) & 63ull);
          const unsigned hi = 2 * j + 1 < kSub ? (unsigned)((cnt >> (12 * j + 6)) & 63ull) : 0u;
          pk[j]   = lo | (hi << 16);
          incl[j] = pk[j];
        }
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
#pragma unroll
          for (int j = 0; j < kSubPk; ++j) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl[j], off);
            if (lane >= off) {
              incl[j] += t;
            }
          }
        }
        // warp totals, double-buffered by the group's piece parity (a fast warp may write the
        // next piece's totals while a slow one still reads this piece's)
        unsigned* wt = wtot + ((k >> 1) & 1) * (kPairSortWarps * kSubPk);
        if (lane == 31) {
#pragma unroll
          for (int j = 0; j < kSubPk; ++j) {
            wt[pwarp * kSubPk + j] = incl[j];
          }
        }
        named_bar_sync(kBarGroup + g, kPairSortGroup);
        // every warp for itself: (packed) prefix over the group's warps before it, totals,
        // padded run starts
        unsigned pre_pk = 0u, tot_pk = 0u;
        if (lane < kSubPk) {
#pragma unroll
          for (int w = 0; w < kPairSortWarps; ++w) {
            const unsigned t = wt[w * kSubPk + lane];
            pre_pk += w < pwarp ? t : 0u;
            tot_pk += t;
          }
        }
        const unsigned tp  = __shfl_sync(0xffffffffu, tot_pk, lane >> 1);
        const int      tot = lane < kSub ? (int)((tp >> ((lane & 1) * 16)) & 0xffffu) : 0;
        const int      pad = (tot + 63) & ~63; // two particles per lane and loop iteration
        int            inc = pad;
#pragma unroll
        for (int off = 1; off < 16; off <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, inc, off);
          if (lane >= off) {
            inc += t;
          }
        }
        const int start_k = inc - pad; // lane k: first sorted entry of run k
        // the buffer is free again once the run warps have left the piece that used it before
        if (k >= kPairBufs) {
          named_bar_sync(kBarEmpty + buf, kHandover);
        }
        if (pwarp == 0 && lane < kSub) {
          runtab[lane] = make_int2(start_k, pad);
          for (int i = tot; i < pad; ++i) {
            B[start_k + i] = make_float2(0.0f, 0.0f); // zero-weight pad
          }
        }
#pragma unroll
        for (int kk = 0; kk < kSub; ++kk) {
          const unsigned sh   = (kk & 1) * 16;
          const int      st_k = __shfl_sync(0xffffffffu, start_k, kk);
          const unsigned pr   = __shfl_sync(0xffffffffu, pre_pk, kk >> 1);
          cur[kk * kPairSortGroup + pt] =
            st_k + (int)((pr >> sh) & 0xffffu) + (int)(((incl[kk >> 1] - pk[kk >> 1]) >> sh) & 0xffffu);
        }
        // the thread's cursor column is its own: read, bump, store
#pragma unroll
        for (int st = 0; st < kPairSortEnt; ++st) {
          int* c        = cur + sub_of(ent[st].x) * kPairSortGroup + pt;
          const int pos = *c;
          *c            = pos + 1;
          B[pos]        = ent[st];
        }
        named_bar_arrive(kBarFull + buf, kHandover);
      }
      return;
    }

    // ================================================================= run warps
    // Two consecutive pieces of the same bucket are streamed as one: the set-up of a lane group
    // (32 broadcasts of fa') and its transpose-reduction are paid once for both.
    double*   prow    = P.partials + (std::size_t)blockIdx.x * P.nslots;
    unsigned long long lane_evals = 0; // hinge evaluations issued by this warp
    int bcur = 0;
    for (int k = 0; k < my_pieces;) {
      const int piece = piece_of(k);
      const int buf   = k % kPairBufs;
      while (pstart[bcur + 1] <= piece) {
        ++bcur;
      }
      const int  b      = bcur;
      // the second piece of the pair, if it belongs to the same bucket
      const bool paired = (k & 1) == 0 && k + 1 < my_pieces && pstart[b + 1] > piece + 1;
      const int  buf2   = (k + 1) % kPairBufs;
      const float2* BA   = reinterpret_cast<const float2*>(smem_raw + P.o_sorted) + (std::size_t)buf * kPairBufLen;
      const float2* BB   = reinterpret_cast<const float2*>(smem_raw + P.o_sorted) + (std::size_t)buf2 * kPairBufLen;
      const int2*   runA = reinterpret_cast<const int2*>(smem_raw + P.o_run) + buf * kSub;
      const int2*   runB = reinterpret_cast<const int2*>(smem_raw + P.o_run) + buf2 * kSub;
      // the bucket's chunk table (these loads complete while the warp waits for the piece)
      const unsigned       em     = P.extmask[b];
      const unsigned char* na_row = P.na_tab + (std::size_t)b * P.nchunks;
      unsigned             na_reg[3];
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        na_reg[q] = q * 32 + lane < P.nchunks ? (unsigned)na_row[q * 32 + lane] : 0u;
      }
      named_bar_sync(kBarFull + buf, kHandover);
      if (paired) {
        named_bar_sync(kBarFull + buf2, kHandover);
      }
      // ---- the runs.  The moments of a pair of pieces go to the second piece's slot.
      float* pm_zero = P.piece_mom + (std::size_t)piece * kMomStride;
      float* pm      = paired ? pm_zero + kMomStride : pm_zero;
      float  s0run   = 0.0f;
      // One lane group (32 bins, lane = bin in the slot tables) over the run(s), with the roles
      // swapped for the loop: a lane owns every 32nd PARTICLE of a run and keeps the partial sums
      // of all 32 bins in registers (the 32 fa' are warp-uniform register copies), so the loop is
      // r = sat(fc + fa'_j), acc_j += w r for j = 0..31 per particle with no shared-memory traffic
      // beyond one coalesced 8-byte load per particle (a broadcast load per particle made the
      // lanes = bins form shared-memory bound at one lane group: 1 wavefront per 2 instructions).
      // A butterfly transpose-reduction then leaves bin j's total in lane j.  The same instruction
      // sequence and the same reduction tree serve every bin, so the zero-tail bins
      // (S0 - sum w sat(u)) are exactly 0 when every particle of the run is beyond the tail.
      auto run_group = [&](const int group, const bool moments, const int r,
                           const float2* __restrict__ reA, const int lenA,
                           const float2* __restrict__ reB, const int lenB) {
        const int2   sl    = slot_tab[group * 32 + lane];
        const float  fa0   = __int_as_float(sl.y);
        const float4 dh    = coef[max(sl.x + b, 0)];
        const bool   spare = sl.x < 0;              // moment lanes and unused lanes
        const bool   tail  = !spare && dh.z < 0.0f; // cell pair next to the zero tail
        const float  fap   = spare ? fa0 - 1.0f : (tail ? fa0 : fa0 - dh.y);
        const float  ds    = spare ? 0.0f : dh.x;
        float f[32], acc[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          f[j]   = __shfl_sync(0xffffffffu, fap, j);
          acc[j] = 0.0f;
        }
        // runs are padded to a multiple of 64 entries: two particles per lane and iteration, the
        // next two in flight
        auto stream = [&](const float2* __restrict__ re, const int len) {
          float2 p0 = re[lane], p1 = re[32 + lane];
#pragma unroll 1
          for (int e = 64; e < len; e += 64) {
            const float2 n0 = re[e + lane], n1 = re[e + 32 + lane];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              acc[j] = fmaf(p0.y, __saturatef(p0.x + f[j]), acc[j]);
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              acc[j] = fmaf(p1.y, __saturatef(p1.x + f[j]), acc[j]);
            }
            p0 = n0;
            p1 = n1;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            acc[j] = fmaf(p0.y, __saturatef(p0.x + f[j]), acc[j]);
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            acc[j] = fmaf(p1.y, __saturatef(p1.x + f[j]), acc[j]);
          }
        };
        if (lenA > 0) {
          stream(reA, lenA);
        }
        if (lenB > 0) {
          stream(reB, lenB);
        }
        // transpose-reduce: after the stage with partner lane ^ off a lane keeps the half of its
        // values whose bin index has that bit equal to its own; lane j ends with bin j
#pragma unroll
        for (int off = 16, n = 32; off >= 1; off >>= 1, n >>= 1) {
          const bool upper = (lane & off) != 0;
#pragma unroll
          for (int i = 0; i < n / 2; ++i) {
            const float send = upper ? acc[i] : acc[i + n / 2];
            const float keep = upper ? acc[i + n / 2] : acc[i];
            acc[i]           = keep + __shfl_xor_sync(0xffffffffu, send, off);
          }
        }
        const float s2 = acc[0];
        if (moments) {
          s0run = __shfl_sync(0xffffffffu, s2, 0);
          if (lane < 2) {
            pm[r * 2 + lane] = s2; // lanes 0 / 1 of the sub-bucket's first group: S0 / S1 of the run(s)
          }
        }
        if (ds != 0.0f) {
          const float v = tail ? s0run - s2 : s2;
          // RED (no return value: nothing waits for the L2 round trip)
          asm volatile("red.global.add.f64 [%0], %1;" ::"l"(prow + group * 32 + lane),
                       "d"((double)ds * (double)v)
                       : "memory");
        }
      };
      auto do_chunks = [&](const int sub, const bool own, const int r, const float2* reA, const int lenA,
                           const float2* reB, const int lenB) {
        for (int c = P.chunk_first[sub]; c < P.chunk_first[sub + 1]; ++c) {
          const int4     ch  = chunks[c];
          const unsigned nav = c < 32 ? na_reg[0] : (c < 64 ? na_reg[1] : na_reg[2]);
          const int      na  = P.force_groups != 0u ? ch.y : (int)__shfl_sync(0xffffffffu, nav, c & 31);
          // the first na lane groups of the chunk are on the table for this bucket
          lane_evals += (unsigned long long)(lenA + lenB) * (unsigned)(na * 32);
          for (int g = 0; g < na; ++g) {
            run_group(ch.x + g, own && ch.z != 0 && g == 0, r, reA, lenA, reB, lenB);
          }
        }
      };
      // warp w streams run w (warp 0 also run 8: with the phase shift runs 0 and 8 are the two
      // parts of one eighth)
#pragma unroll 1
      for (int r = warp; r < kSub; r += kPairRunThreads / 32) {
        const int2 rtA = runA[r];
        const int2 rtB = paired ? runB[r] : make_int2(0, 0);
        if (paired && lane < 2) {
          pm_zero[r * 2 + lane] = 0.0f; // the pair's moments are booked under the second piece
        }
        if (rtA.y + rtB.y == 0) {
          if (lane < 2) {
            pm[r * 2 + lane] = 0.0f;
          }
          continue;
        }
        const float2* reA = BA + rtA.x;
        const float2* reB = BB + rtB.x;
        do_chunks(r, true, r, reA, rtA.y, reB, rtB.y); // the sub-bucket's own bins (first group: moment lanes)
        if (r + 1 < kSub && ((em >> (r + 1)) & 1u)) {
          do_chunks(r + 1, false, r, reA, rtA.y, reB, rtB.y); // a threshold of sub-bucket r + 1 strays down here
        }
        if (r > 0 && ((em >> (kSub + r - 1)) & 1u)) {
          do_chunks(r - 1, false, r, reA, rtA.y, reB, rtB.y); // a threshold of sub-bucket r - 1 strays up here
        }
      }
      // hand the buffers back if a sort group will use them again
      if (k + kPairBufs < my_pieces) {
        named_bar_arrive(kBarEmpty + buf, kHandover);
      }
      if (paired && k + 1 + kPairBufs < my_pieces) {
        named_bar_arrive(kBarEmpty + buf2, kHandover);
      }
      k += paired ? 2 : 1;
    }
    if (lane == 0 && lane_evals != 0ull) {
      atomicAdd(P.lane_evals, lane_evals);
    }
  }

  // msum[b * (kSub + 1) + a] = {sum_{s >= a} S0_{b,s}, sum_{s >= a} S1_{b,s}}: suffix sums over the
  // sub-buckets of bucket b (a = 0: the whole bucket, a = kSub: 0) of the fp64 sums over the
  // bucket's pieces, lane-strided then a fixed shuffle tree.  Also counts (roofline accounting,
  // integer atomics) the pairs of the bucket whose cell pair is not identically zero: what an
  // ideal kernel evaluating every pair would have to touch.
  __global__ void __launch_bounds__(kPThreads)
    pair_moments_kernel(const int* __restrict__ tot, int nb, const float* __restrict__ piece_mom,
                        double2* __restrict__ msum, const int2* __restrict__ slot_i,
                        const int* __restrict__ bin_of_slot, const float4* __restrict__ coef_dh,
                        int nslots, unsigned long long* __restrict__ ontable) {
    __shared__ int    bstart[kPMaxBuckets + 2], pstart[kPMaxBuckets + 2], tmp[2 * kPWarps];
    __shared__ double red[kPWarps][2 * kSub];
    __shared__ int    red_live[kPWarps];
    block_scan_buckets(tot, nb, bstart, pstart, tmp);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b   = blockIdx.x; // one CTA per bucket
    double    s0[kSub], s1[kSub];
#pragma unroll
    for (int s = 0; s < kSub; ++s) {
      s0[s] = 0.0;
      s1[s] = 0.0;
    }
    for (int p = pstart[b] + tid; p < pstart[b + 1]; p += kPThreads) {
      const float4* m4 = reinterpret_cast<const float4*>(piece_mom + (std::size_t)p * kMomStride);
#pragma unroll
      for (int h = 0; h < kMomStride / 4; ++h) {
        const float4 m = m4[h];
        s0[2 * h] += (double)m.x;
        s1[2 * h] += (double)m.y;
        if (2 * h + 1 < kSub) {
          s0[2 * h + 1] += (double)m.z;
          s1[2 * h + 1] += (double)m.w;
        }
      }
    }
    int live = 0;
    for (int sl = tid; sl < nslots; sl += kPThreads) {
      live += (bin_of_slot[sl] >= 0 && coef_dh[max(slot_i[sl].x + b, 0)].x != 0.0f) ? 1 : 0;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
      for (int s = 0; s < kSub; ++s) {
        s0[s] += __shfl_xor_sync(0xffffffffu, s0[s], off);
        s1[s] += __shfl_xor_sync(0xffffffffu, s1[s], off);
      }
      live += __shfl_xor_sync(0xffffffffu, live, off);
    }
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < kSub; ++s) {
        red[warp][2 * s]     = s0[s];
        red[warp][2 * s + 1] = s1[s];
      }
      red_live[warp] = live;
    }
    __syncthreads();
    if (tid == 0) {
      double2* out = msum + (std::size_t)b * (kSub + 1);
      double   a0 = 0.0, a1 = 0.0;
      out[kSub]   = make_double2(0.0, 0.0);
      for (int s = kSub - 1; s >= 0; --s) {
        for (int w = 0; w < kPWarps; ++w) { // fixed order
          a0 += red[w][2 * s];
          a1 += red[w][2 * s + 1];
        }
        out[s] = make_double2(a0, a1);
      }
      int lv = 0;
      for (int w = 0; w < kPWarps; ++w) {
        lv += red_live[w];
      }
      if (tot[b] != 0 && lv != 0) {
        atomicAdd(ontable, (unsigned long long)tot[b] * (unsigned long long)lv);
      }
    }
  }

  // out[slot] = sum_rows hinge partials
  //           + sum_b ( v_q S0_b + s_q (fa S0_b + S1_b) )                       line part
  //           + sum_b ds_q * ( S1_{b,>pw} + fa' S0_{b,>pw} )                    hinge, sub-buckets above
  //             the bin's threshold (r = fc + fa' there; sub-buckets below contribute 0), or, for the
  //             cell pair next to the zero tail, ds_q * ( (1 - fa) S0_{b,<pw} - S1_{b,<pw} );
  // one warp per slot, lane-strided sums and a fixed shuffle tree
  __global__ void __launch_bounds__(kPThreads)
    pair_final_kernel(const double* __restrict__ partials, int nrows, int nslots,
                      const int2* __restrict__ slot_i, const float2* __restrict__ slot_f,
                      const int* __restrict__ slot_sub, const unsigned* __restrict__ extmask,
                      const float4* __restrict__ coef_dh,
                      const double2* __restrict__ coef_vs, const double2* __restrict__ msum,
                      int nb, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int j    = blockIdx.x * kPWarps + (threadIdx.x >> 5);
    if (j >= nslots) {
      return;
    }
    double s = 0.0;
    for (int c = lane; c < nrows; c += 32) {
      s += partials[(std::size_t)c * nslots + j];
    }
    const int2   si  = slot_i[j];
    const float  faf = slot_f[j].x;
    const double fa  = (double)faf;
    const int    s0  = slot_sub[j]; // the sub-bucket of the bin's hinge threshold
    if (si.x >= 0) {
      for (int b = lane; b < nb; b += 32) {
        const double2* mb = msum + (std::size_t)b * (kSub + 1);
        const double2  vs = coef_vs[si.x + b];
        const float4   dh = coef_dh[si.x + b];
        const double2  m0 = mb[0];
        double         t  = fma(vs.x, m0.x, vs.y * fma(fa, m0.x, m0.y));
        if (dh.x != 0.0f) {
          // sub-buckets [pw_lo, pw_hi] were evaluated pair by pair for this bucket
          const unsigned em    = extmask[b];
          const int      pw_lo = s0 - (int)((em >> s0) & 1u);
          const int      pw_hi = s0 + (int)((em >> (kSub + s0)) & 1u);
          if (dh.z > 0.0f) {
            const float   fap = faf - dh.y; // the pair kernel's float fa'
            const double2 ma  = mb[pw_hi + 1];
            t = fma((double)dh.x, fma((double)fap, ma.x, ma.y), t);
          } else {
            const double2 mlo = mb[pw_lo];
            t = fma((double)dh.x, fma(1.0 - fa, m0.x - mlo.x, -(m0.y - mlo.y)), t);
          }
        }
        s += t;
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, off);
    }
    if (lane == 0) {
      out[j] = s;
    }
  }
  // ------------------------------------------------------------------ host side
  struct PairPlan {
    std::vector<int2>    slot_i;
    std::vector<float2>  slot_f;
    std::vector<int>     slot_sub; // sub-bucket of the slot's lane group
    std::vector<float4>  coef_dh;
    std::vector<double2> coef_vs;
    std::vector<int>     bin_of_slot;
    std::vector<int4>    chunks;   // {first lane group, groups, carries the moment lanes, -}
    std::vector<unsigned char> na_tab; // [bucket][chunk] leading groups of the chunk on the table
    std::vector<unsigned>      extmask; // [bucket] sub-buckets that also see a neighbouring run
    int    chunk_first[kSub + 1] {};
    int    ngroups { 0 }, nslots { 0 }, n_pad { 0 }, nb { 0 }, nbp { 0 }, kmin { 0 };
    double c0 { 0 }, c_lo { 0 }, c_hi { 0 };
    float  sub_phi { 0 };          // s = floor(kSubDiv fc + sub_phi)
    bool   ok { true };            // false: a hinge threshold strays beyond the neighbouring sub-bucket
  };

  static bool pair_shape_eligible(const TablePlan& tp, const float* bins_e_syn,
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
    return (double)tp.T + std::ceil(spread) + 4.0 <= (double)kPMaxBuckets;
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
    // ---- bins by the sub-bucket of their hinge threshold.  For a particle of fraction fc the
    // hinge of bin j (fraction fa, node at h ~ 1) is active for fc > h - fa; the bin goes to
    // sub-bucket s0 = floor(kSubDiv (1 - fa) + phi).  The phase phi puts the sub-bucket boundaries
    // into the widest gap between the thresholds (bin grids commensurate with the table grid --
    // 200 bins over 7 decades on the 200-point table over 8 -- put every threshold on a multiple
    // of 1/8).  Every sub-bucket owns whole lane groups; the
    // first two lanes of its first group are the moment lanes (S0: fa0 = 2 -> r = 1,
    // S1: fa0 = 1 -> r = fc; a sub-bucket without bins still gets them).  Inside a
    // sub-bucket the bins ascend in energy, so the groups a bucket cannot reach are the
    // trailing ones.
    struct BinSlot {
      int    bin, A, s0;
      float  fa;
      double a;
    };
    std::vector<int> order(nbin);
    for (int s = 0; s < nbin; ++s) {
      order[s] = s;
    }
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return a[x] < a[y]; });
    std::vector<std::vector<BinSlot>> by_sub(kSub);
    std::vector<float>                fa_of(nbin);
    std::vector<double>               A_of(nbin);
    {
      std::vector<double> v(nbin); // position of the threshold inside its eighth
      for (int s = 0; s < nbin; ++s) {
        const double rel = a[s] - amin;
        double       A   = std::floor(rel);
        float        fa  = (float)(rel - A);
        if (fa >= 1.0f) { // rounding of the fraction to float
          fa = 0.0f;
          A += 1.0;
        }
        fa_of[s]       = fa;
        A_of[s]        = A;
        const double t = (1.0 - (double)fa) * (double)kSubDiv;
        v[s]           = t - std::floor(t);
      }
      std::sort(v.begin(), v.end());
      double best = -1.0, centre = 0.5;
      for (int s = 0; s < nbin; ++s) { // circular gaps
        const double lo = v[s], hi = s + 1 < nbin ? v[s + 1] : v[0] + 1.0;
        if (hi - lo > best) {
          best   = hi - lo;
          centre = 0.5 * (lo + hi);
        }
      }
      centre -= std::floor(centre);
      // boundaries where kSubDiv t + phi is an integer, i.e. frac(kSubDiv t) = 1 - phi
      double phi = 1.0 - centre;
      phi -= std::floor(phi);
      pp.sub_phi = (float)phi;
      if (const char* fp = std::getenv("RGC_PAIR_PHI")) { // test knob: force the phase (e.g. 0 puts
        pp.sub_phi = (float)std::atof(fp);                // BASELINE's thresholds ON the boundaries)
      }
      if (!(pp.sub_phi >= 0.0f && pp.sub_phi < 1.0f)) {
        pp.sub_phi = 0.0f;
      }
    }
    const double phi = (double)pp.sub_phi;
    for (int k = 0; k < nbin; ++k) {
      const int    s  = order[k];
      const float  fa = fa_of[s];
      const double A  = A_of[s];
      int s0 = (int)std::floor((1.0 - (double)fa) * (double)kSubDiv + phi);
      s0     = std::max(0, std::min(kSub - 1, s0));
      by_sub[s0].push_back(BinSlot { bins[s], (int)A + pp.kmin, s0, fa, a[s] });
    }
    int first_group[kSub + 1];
    first_group[0] = 0;
    for (int s = 0; s < kSub; ++s) {
      first_group[s + 1] = first_group[s] + ((int)by_sub[s].size() + 2 + 31) / 32;
    }
    pp.ngroups = first_group[kSub];
    pp.nslots  = pp.ngroups * 32;
    // spare slots sit on padded cell 0 (the kernel gives them fa' = fa0 - 1 and ds = 0)
    pp.slot_i.assign(pp.nslots, make_int2(-(1 << 20), 0));
    pp.slot_f.assign(pp.nslots, make_float2(1.0f, 0.0f));
    pp.slot_sub.assign(pp.nslots, 0);
    pp.bin_of_slot.assign(pp.nslots, -1);
    // ---- per bucket: which neighbouring run a sub-bucket's groups must also see (the float
    // threshold -fa' = h_q - fa of an L cell wanders with the node position h_q, 1 +- 1e-5, and
    // may leave the bin's nominal sub-bucket by a hair), and which lane groups are on the table
    pp.extmask.assign(pp.nb, 0u);
    std::vector<unsigned char> group_live((std::size_t)pp.nb * pp.ngroups, 0);
    pp.ok = true;
    for (int s = 0; s < kSub; ++s) {
      const int base = first_group[s] * 32;
      pp.slot_f[base]     = make_float2(2.0f, 0.0f); // r = sat(1 + fc) = 1 -> S0
      pp.slot_f[base + 1] = make_float2(1.0f, 0.0f); // r = sat(fc)     = fc -> S1
      for (int g = first_group[s]; g < first_group[s + 1]; ++g) {
        for (int l = 0; l < 32; ++l) {
          pp.slot_sub[g * 32 + l] = s;
        }
      }
      for (std::size_t k = 0; k < by_sub[s].size(); ++k) {
        const BinSlot& bs   = by_sub[s][k];
        const int      slot = base + 2 + (int)k;
        pp.slot_i[slot]      = make_int2(bs.A, 1);
        pp.slot_f[slot]      = make_float2(bs.fa, 1.0f);
        pp.bin_of_slot[slot] = bs.bin;
        for (int b = 0; b < pp.nb; ++b) {
          const float4 dh = pp.coef_dh[std::max(bs.A + b, 0)];
          if (dh.x == 0.0f) {
            continue; // zero cell pair
          }
          group_live[(std::size_t)b * pp.ngroups + slot / 32] = 1;
          // threshold in fc: the hinge of an L cell is active for fc > -fa' (the kernel's float
          // fa'), the zero-tail form changes at fc = 1 - fa.  Sub-buckets below the evaluated
          // range must lie entirely on one side, those above entirely on the other; eps covers
          // the float rounding of the kernel's floor(kSubDiv fc + phi)
          const double tt  = dh.z < 0.0f ? 1.0 - (double)bs.fa : -(double)(bs.fa - dh.y);
          const double eps = 1e-6;
          auto edge = [&](int k) { return ((double)k - phi) / (double)kSubDiv; }; // lower edge of sub-bucket k
          if (s > 0 && tt < edge(s) + eps) {
            pp.extmask[b] |= 1u << s;
            if (s > 1 && tt < edge(s - 1) + eps) {
              pp.ok = false;
            }
          }
          if (s < kSub - 1 && tt > edge(s + 1) - eps) {
            pp.extmask[b] |= 1u << (kSub + s);
            if (s < kSub - 2 && tt > edge(s + 2) - eps) {
              pp.ok = false;
            }
          }
        }
      }
    }
    // ---- chunks of <= kPMaxGPW groups
    pp.chunks.clear();
    for (int s = 0; s < kSub; ++s) {
      pp.chunk_first[s] = (int)pp.chunks.size();
      for (int g = first_group[s]; g < first_group[s + 1]; g += kPMaxGPW) {
        pp.chunks.push_back(make_int4(g, std::min(kPMaxGPW, first_group[s + 1] - g),
                                      g == first_group[s] ? 1 : 0, 0));
      }
    }
    pp.chunk_first[kSub] = (int)pp.chunks.size();
    const int nchunks = (int)pp.chunks.size();
    pp.na_tab.assign((std::size_t)pp.nb * nchunks, 0);
    for (int b = 0; b < pp.nb; ++b) {
      for (int c = 0; c < nchunks; ++c) {
        int na = pp.chunks[c].z ? 1 : 0; // the moment lanes always run
        for (int g = 0; g < pp.chunks[c].y; ++g) {
          if (group_live[(std::size_t)b * pp.ngroups + pp.chunks[c].x + g]) {
            na = g + 1;
          }
        }
        pp.na_tab[(std::size_t)b * nchunks + c] = (unsigned char)na;
      }
    }
  }

  // Host-only summary of the plan (rgc_pair_plan_describe; CPU tests of the lane-group layout)
  void pair_plan_describe(const TablePlan& tp, const float* bins_e_syn, const std::vector<int>& bins,
                          int info[8], float* phase, int* slot_bin, std::size_t cap) {
    info[0] = 0;
    for (int k = 1; k < 8; ++k) {
      info[k] = 0;
    }
    if (!pair_shape_eligible(tp, bins_e_syn, bins)) {
      return;
    }
    PairPlan pp;
    make_pair_plan(tp, bins_e_syn, bins, pp);
    int ext = 0, most = 0;
    for (unsigned m : pp.extmask) {
      ext += m != 0u ? 1 : 0;
    }
    for (int s = 0; s < kSub; ++s) {
      int groups = 0;
      for (int c = pp.chunk_first[s]; c < pp.chunk_first[s + 1]; ++c) {
        groups += pp.chunks[c].y;
      }
      most = std::max(most, groups);
    }
    const bool fits =
      pair_smem_layout(pp.n_pad, pp.nbp, pp.nslots, (int)pp.chunks.size()).total <= 226 * 1024;
    info[0] = (pp.ok && pp.ngroups <= kPMaxGroups && pp.nb < kPMaxBuckets && pp.nbp <= kPMaxBuckets && fits) ? 1 : 0;
    info[1] = pp.ngroups;
    info[2] = pp.nslots;
    info[3] = pp.nb;
    info[4] = (int)pp.chunks.size();
    info[5] = ext;
    info[6] = most;
    info[7] = kSub;
    if (phase) {
      *phase = pp.sub_phi;
    }
    for (std::size_t i = 0; slot_bin && i < cap && i < (std::size_t)pp.nslots; ++i) {
      slot_bin[i] = pp.bin_of_slot[i];
    }
  }

  // Plans are kept with their device copy (slot tables, hinge and line coefficients):
  // a repeated call with the same photon bins and F table uploads nothing.
  struct CachedPlan {
    std::vector<float>  key_bins;  // e_syn of the chunk's bins, in slot order
    std::vector<int>    key_index; // their indices in the caller's bin array
    std::vector<double> key_tx, key_y;
    std::string         key_phi; // RGC_PAIR_PHI (test knob) the plan was made under
    PairPlan            pp;
    char*               dev { nullptr };
    std::size_t         off_map { 0 }, off_si { 0 }, off_sf { 0 }, off_dh { 0 }, off_vs { 0 };
    std::size_t         off_sub { 0 }, off_chunk { 0 }, off_na { 0 }, off_ext { 0 };
  };
  static std::vector<CachedPlan>& plan_cache() {
    static std::vector<CachedPlan> cache;
    return cache;
  }

  static double2* g_log_tab = nullptr;
  static int g_rank_order_ok = -1; // -1 unknown, 0 lanes are NOT served in lane order, 1 verified

  static int ensure_log_table(const double2** out) {
    if (!g_log_tab) {
      std::vector<double2> tab(256);
      for (int j = 0; j < 256; ++j) {
        const double mh = 1.0 + (double)j / 256.0;
        tab[j]          = make_double2(std::log2(mh), 1.0 / mh);
      }
      RGC_CUDA(cudaMalloc(reinterpret_cast<void**>(&g_log_tab), tab.size() * sizeof(double2)));
      RGC_CUDA(cudaMemcpy(g_log_tab, tab.data(), tab.size() * sizeof(double2), cudaMemcpyHostToDevice));
    }
    *out = g_log_tab;
    return RGC_OK;
  }

  void pair_release_plans() {
    if (g_log_tab) {
      cudaFree(g_log_tab);
      g_log_tab = nullptr;
    }
    for (auto& e : plan_cache()) {
      if (e.dev) {
        cudaFree(e.dev);
      }
    }
    plan_cache().clear();
    g_rank_order_ok = -1;
  }

  static int cached_pair_plan(const TablePlan& tp, const float* bins_e_syn,
                              const std::vector<int>& bins, const CachedPlan** out) {
    auto&              cache = plan_cache();
    std::vector<float> kb(bins.size());
    for (std::size_t s = 0; s < bins.size(); ++s) {
      kb[s] = bins_e_syn[bins[s]];
    }
    const char*       fp = std::getenv("RGC_PAIR_PHI");
    const std::string key_phi = fp ? fp : "";
    for (const auto& e : cache) {
      if (e.key_phi == key_phi && e.key_index == bins && e.key_bins.size() == kb.size() &&
          std::memcmp(e.key_bins.data(), kb.data(), kb.size() * sizeof(float)) == 0 &&
          e.key_tx == tp.tx && e.key_y == tp.y) {
        *out = &e;
        return RGC_OK;
      }
    }
    if (cache.size() >= 8) {
      RGC_CUDA(cudaStreamSynchronize(ctx().stream));
      cudaFree(cache.front().dev);
      cache.erase(cache.begin());
    }
    CachedPlan e;
    e.key_bins  = kb;
    e.key_index = bins;
    e.key_phi   = key_phi;
    e.key_tx    = tp.tx;
    e.key_y     = tp.y;
    make_pair_plan(tp, bins_e_syn, bins, e.pp);
    const PairPlan& pp = e.pp;
    auto align = [](std::size_t x) { return (x + 255) & ~std::size_t(255); };
    e.off_map = 0;
    e.off_si  = align(e.off_map + pp.nslots * sizeof(int));
    e.off_sf  = align(e.off_si + pp.nslots * sizeof(int2));
    e.off_dh  = align(e.off_sf + pp.nslots * sizeof(float2));
    e.off_vs  = align(e.off_dh + pp.n_pad * sizeof(float4));
    e.off_sub   = align(e.off_vs + pp.n_pad * sizeof(double2));
    e.off_chunk = align(e.off_sub + pp.nslots * sizeof(int));
    e.off_na    = align(e.off_chunk + pp.chunks.size() * sizeof(int4));
    e.off_ext   = align(e.off_na + pp.na_tab.size());
    const std::size_t total = align(e.off_ext + pp.extmask.size() * sizeof(unsigned));
    RGC_CUDA(cudaMalloc(reinterpret_cast<void**>(&e.dev), total));
    cudaStream_t st = ctx().stream;
    RGC_CUDA(cudaMemcpyAsync(e.dev + e.off_map, pp.bin_of_slot.data(), pp.nslots * sizeof(int),
                             cudaMemcpyHostToDevice, st));
    RGC_CUDA(cudaMemcpyAsync(e.dev + e.off_si, pp.slot_i.data(), pp.nslots * sizeof(int2),
                             cudaMemcpyHostToDevice, st));
    RGC_CUDA(cudaMemcpyAsync(e.dev + e.off_sf, pp.slot_f.data(), pp.nslots * sizeof(float2),
                             cudaMemcpyHostToDevice, st));
    RGC_CUDA(cudaMemcpyAsync(e.dev + e.off_dh, pp.coef_dh.data(), pp.n_pad * sizeof(float4),
                             cudaMemcpyHostToDevice, st));
    RGC_CUDA(cudaMemcpyAsync(e.dev + e.off_vs, pp.coef_vs.data(), pp.n_pad * sizeof(double2),
                             cudaMemcpyHostToDevice, st));
    RGC_CUDA(cudaMemcpyAsync(e.dev + e.off_sub, pp.slot_sub.data(), pp.nslots * sizeof(int),
                             cudaMemcpyHostToDevice, st));
    RGC_CUDA(cudaMemcpyAsync(e.dev + e.off_chunk, pp.chunks.data(), pp.chunks.size() * sizeof(int4),
                             cudaMemcpyHostToDevice, st));
    RGC_CUDA(cudaMemcpyAsync(e.dev + e.off_na, pp.na_tab.data(), pp.na_tab.size(),
                             cudaMemcpyHostToDevice, st));
    RGC_CUDA(cudaMemcpyAsync(e.dev + e.off_ext, pp.extmask.data(), pp.extmask.size() * sizeof(unsigned),
                             cudaMemcpyHostToDevice, st));
    cache.push_back(std::move(e));
    *out = &cache.back();
    return RGC_OK;
  }

  // The hinge path takes a chunk of bins when its shape fits and the plan keeps every hinge
  // threshold within one sub-bucket of its nominal place (always, for tables on a log grid).
  bool pair_path_eligible(const TablePlan& tp, const float* bins_e_syn,
                          const std::vector<int>& bins) {
    if (!pair_shape_eligible(tp, bins_e_syn, bins)) {
      return false;
    }
    const CachedPlan* cp = nullptr;
    if (cached_pair_plan(tp, bins_e_syn, bins, &cp) != RGC_OK) {
      return false;
    }
    const PairPlan& pp = cp->pp;
    // (a very wide table with ~2000 bins can outgrow the pair kernel's shared memory)
    const bool fits = pair_smem_layout(pp.n_pad, pp.nbp, pp.nslots, (int)pp.chunks.size()).total <= 226 * 1024;
    return pp.ok && pp.ngroups <= kPMaxGroups && pp.nb < kPMaxBuckets && pp.nbp <= kPMaxBuckets && fits;
  }

  // Runs rank_order_probe_kernel once per process (per device context) and remembers the verdict.
  static int rank_order_verified(bool* ok) {
    if (g_rank_order_ok < 0) {
      auto& c       = ctx();
      void* scratch = nullptr; // (laid out anew by the caller right after)
      RGC_TRY(ensure_scratch(64, &scratch));
      int* d_bad = static_cast<int*>(scratch);
      RGC_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), c.stream));
      rank_order_probe_kernel<<<2 * c.sm_count, kPThreads, 0, c.stream>>>(256, 0x9e3779b9u, d_bad);
      RGC_CUDA(cudaGetLastError());
      count_launch(1);
      int bad = 0;
      RGC_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
      RGC_CUDA(cudaStreamSynchronize(c.stream));
      g_rank_order_ok = bad == 0 ? 1 : 0;
    }
    *ok = g_rank_order_ok == 1;
    return RGC_OK;
  }

  int pair_rank_mode() { return g_rank_order_ok; }

  // particles per pipeline pass (bounds the staged and sorted arrays: 18 B per particle)
  static std::size_t pair_pass_max() {
    std::size_t chunk_max = std::size_t(1) << 27;
    if (const char* pm = std::getenv("RGC_PAIR_PASS_MAX")) { // test knob
      const long long v = std::atoll(pm);
      if (v >= kPTile && (std::size_t)v < chunk_max) {
        chunk_max = (std::size_t)v / kPTile * kPTile;
      }
    }
    return chunk_max;
  }

  bool pair_single_pass(std::size_t n) { return n <= pair_pass_max(); }

  // One pipeline pass per <= 2^27 particles over one chunk of bins.
  // d_acc[bins[s]] += sum_i w_i F_is (before the e_syn factor), on the device;
  // *d_poison is raised when a particle's chiR overflows float (see pair_prologue).
  int run_spectrum_pair(const rgc_particles_t* prtls, std::size_t n, float B0, float g_syn,
                        float e_at, const TablePlan& tp, const float* bins_e_syn,
                        const std::vector<int>& bins, double* d_acc, int* d_poison,
                        float* main_ms, bool defer_sync) {
    auto&             c  = ctx();
    const CachedPlan* cp = nullptr;
    RGC_TRY(cached_pair_plan(tp, bins_e_syn, bins, &cp));
    const PairPlan& pp = cp->pp;
    if (pp.nb >= kPMaxBuckets || pp.nbp > kPMaxBuckets) {
      return fail(RGC_ERR_INVALID, "internal: %d buckets exceed the pair path's limit", pp.nb);
    }
    const PairSmem    L    = pair_smem_layout(pp.n_pad, pp.nbp, pp.nslots, (int)pp.chunks.size());
    const std::size_t smem = L.total;
    if (smem > 226 * 1024) {
      return fail(RGC_ERR_INVALID, "internal: pair kernel needs %zu B of shared memory", smem);
    }
    // particles are processed in passes so the staged and sorted (fc, w, key) stay
    // bounded (18 B per particle); pass results are summed on the host in pass order
    const std::size_t chunk_max = pair_pass_max();
    const std::size_t cnt0      = std::min(n, chunk_max);
    // "ballot": ranking whose order is fixed by construction; "atomic": skip the probe
    const char* sr          = std::getenv("RGC_SORT_RANK");
    bool        atomic_rank = !(sr && std::strcmp(sr, "ballot") == 0);
    if (atomic_rank && !(sr && std::strcmp(sr, "atomic") == 0)) {
      RGC_TRY(rank_order_verified(&atomic_rank));
    }
    const int sort_ctas_per_sm = 2;
    const char* pm       = std::getenv("RGC_PROLOGUE_MINB"); // tuning knob: prologue CTAs per SM
    const int   pro_minb = (pm && std::atoi(pm) == 4) ? 4 : 3;
    struct Geom {
      int ntiles, rows, tiles_per_row, ctas1, tiles_per_cta1;
    };
    auto geom_for = [&](std::size_t cnt) {
      Geom g;
      g.ntiles         = (int)((cnt + kPTile - 1) / kPTile);
      g.rows           = std::max(1, std::min(c.sm_count * sort_ctas_per_sm, g.ntiles));
      g.tiles_per_row  = (g.ntiles + g.rows - 1) / g.rows;
      g.rows           = (g.ntiles + g.tiles_per_row - 1) / g.tiles_per_row;
      g.ctas1          = std::max(1, std::min(c.sm_count * pro_minb, g.ntiles));
      g.tiles_per_cta1 = (g.ntiles + g.ctas1 - 1) / g.ctas1;
      g.ctas1          = (g.ntiles + g.tiles_per_cta1 - 1) / g.tiles_per_cta1;
      return g;
    };
    const Geom        g0    = geom_for(cnt0);
    const int         rows0 = g0.rows;
    const std::size_t npad0 = (std::size_t)g0.ntiles * kPTile;
    const int  pair_ctas = c.sm_count; // one warp-specialised CTA of 16 warps per SM
    const std::size_t max_pieces  = cnt0 / kPieceLen + (std::size_t)pp.nbp + 2;
    auto align = [](std::size_t x) { return (x + 255) & ~std::size_t(255); };
    const std::size_t off_msum = 0;
    const std::size_t off_out  = align(off_msum + (std::size_t)pp.nb * (kSub + 1) * sizeof(double2));
    const std::size_t off_part = align(off_out + pp.nslots * sizeof(double));
    const std::size_t part_bytes = (std::size_t)pair_ctas * pp.nslots * sizeof(double);
    const std::size_t off_tot  = align(off_part + part_bytes);
    const std::size_t off_cnt  = align(off_tot + (std::size_t)pp.nbp * sizeof(int));
    const std::size_t off_mom  = align(off_cnt + (std::size_t)pp.nbp * rows0 * sizeof(int));
    const std::size_t off_cw   = align(off_mom + max_pieces * kMomStride * sizeof(float));
    const std::size_t off_keys = align(off_cw + npad0 * sizeof(float2));
    const std::size_t off_sort = align(off_keys + npad0 * sizeof(unsigned short));
    const std::size_t total    = off_sort + (cnt0 + pp.nbp + 64) * sizeof(float2);
    void*             scratch  = nullptr;
    RGC_TRY(ensure_scratch(total, &scratch));
    char* sb = static_cast<char*>(scratch);
    PairParams P {};
    P.kc = PairConsts { 1.0 / 3.0, -0.25, -0.5, 1.4426950408889634, 536870913.0, 6755399441055744.0,
                        3.4028235677973366e38, 1e-37, 3.4028234e38, 1e-280 };
    RGC_TRY(ensure_log_table(&P.log_tab));
    P.slot_i   = reinterpret_cast<const int2*>(cp->dev + cp->off_si);
    P.slot_f   = reinterpret_cast<const float2*>(cp->dev + cp->off_sf);
    P.coef_dh  = reinterpret_cast<const float4*>(cp->dev + cp->off_dh);
    P.n_pad    = pp.n_pad;
    P.nb       = pp.nb;
    P.nbp      = pp.nbp;
    P.chunks   = reinterpret_cast<const int4*>(cp->dev + cp->off_chunk);
    P.na_tab   = reinterpret_cast<const unsigned char*>(cp->dev + cp->off_na);
    P.extmask  = reinterpret_cast<const unsigned*>(cp->dev + cp->off_ext);
    P.nchunks  = (int)pp.chunks.size();
    P.sub_phi  = pp.sub_phi;
    for (int r = 0; r <= kSub; ++r) {
      P.chunk_first[r] = pp.chunk_first[r];
    }
    P.kmin     = pp.kmin;
    P.kmin_d   = (double)pp.kmin;
    P.inv_B0           = 1.0 / (double)B0;
    P.e_scale          = (double)e_at / (double)(g_syn * g_syn);
    P.cells_per_octave = 0.30102999566398119521 / tp.dL;
    P.c0        = pp.c0;
    P.inv_dL    = 1.0 / tp.dL;
    P.c_lo      = pp.c_lo;
    P.c_hi      = pp.c_hi;
    P.cw        = reinterpret_cast<float2*>(sb + off_cw);
    P.keys      = reinterpret_cast<unsigned short*>(sb + off_keys);
    P.counts    = reinterpret_cast<int*>(sb + off_cnt);
    P.tot       = reinterpret_cast<int*>(sb + off_tot);
    P.poison    = d_poison;
    {
      const char* ns  = std::getenv("RGC_PAIR_NO_SKIP"); // test knob
      P.force_groups  = (ns && ns[0] == '1') ? 1u : 0u;
    }
    P.lane_evals = reinterpret_cast<unsigned long long*>(d_poison) + 1; // [issued, on-table] behind the flag
    P.sorted    = reinterpret_cast<float2*>(sb + off_sort);
    P.piece_mom = reinterpret_cast<float*>(sb + off_mom);
    P.partials  = reinterpret_cast<double*>(sb + off_part);
    P.nslots    = pp.nslots;
    P.o_coef = (int)L.coef; P.o_bstart = (int)L.bstart; P.o_pstart = (int)L.pstart;
    P.o_tmp = (int)L.tmp; P.o_slot = (int)L.slot; P.o_chunk = (int)L.chunk;
    P.o_sorted = (int)L.sorted; P.o_cur = (int)L.cur; P.o_wtot = (int)L.wtot; P.o_run = (int)L.run;
    double2* d_msum = reinterpret_cast<double2*>(sb + off_msum);
    RGC_CUDA(cudaFuncSetAttribute(sync_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    double* d_out  = reinterpret_cast<double*>(sb + off_out);
    float               pro_ms = 0.f, sort_ms = 0.f;
    const std::size_t   sort_smem = sort_smem_layout(pp.nbp).total;
    RGC_CUDA(cudaFuncSetAttribute(sync_sort_kernel<true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem));
    RGC_CUDA(cudaFuncSetAttribute(sync_sort_kernel<false>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem));
    for (std::size_t off = 0; off < n; off += chunk_max) {
      const std::size_t cnt = std::min(chunk_max, n - off);
      for (int d = 0; d < 3; ++d) {
        P.u[d] = prtls->col[RGC_Q_U][d] + off;
        P.e[d] = prtls->col[RGC_Q_E][d] + off;
        P.b[d] = prtls->col[RGC_Q_B][d] + off;
      }
      const Geom g     = geom_for(cnt);
      P.nprtl          = cnt;
      P.ntiles         = g.ntiles;
      P.rows           = g.rows;
      P.tiles_per_row  = g.tiles_per_row;
      P.tiles_per_cta1 = g.tiles_per_cta1;
      RGC_CUDA(cudaEventRecord(c.ev[2], c.stream));
      RGC_CUDA(cudaMemsetAsync(P.counts, 0, (std::size_t)pp.nbp * g.rows * sizeof(int), c.stream));
      if (pro_minb == 4) {
        sync_prologue_kernel<4><<<g.ctas1, kPThreads, 0, c.stream>>>(P);
      } else {
        sync_prologue_kernel<3><<<g.ctas1, kPThreads, 0, c.stream>>>(P);
      }
      RGC_CUDA(cudaGetLastError());
      RGC_CUDA(cudaEventRecord(c.ev[3], c.stream));
      pair_colscan_kernel<<<(pp.nb + kPWarps - 1) / kPWarps, kPThreads, 0, c.stream>>>(
        P.counts, P.rows, pp.nb, P.tot);
      RGC_CUDA(cudaGetLastError());
      if (atomic_rank) {
        sync_sort_kernel<true><<<g.rows, kPThreads, sort_smem, c.stream>>>(P);
      } else {
        sync_sort_kernel<false><<<g.rows, kPThreads, sort_smem, c.stream>>>(P);
      }
      RGC_CUDA(cudaGetLastError());
      RGC_CUDA(cudaEventRecord(c.ev[5], c.stream));
      RGC_CUDA(cudaMemsetAsync(P.partials, 0, part_bytes, c.stream));
      sync_pair_kernel<<<pair_ctas, kPairThreads, smem, c.stream>>>(P);
      RGC_CUDA(cudaGetLastError());
      RGC_CUDA(cudaEventRecord(c.ev[4], c.stream));
      pair_moments_kernel<<<pp.nb, kPThreads, 0, c.stream>>>(
        P.tot, pp.nb, P.piece_mom, d_msum, P.slot_i, reinterpret_cast<const int*>(cp->dev + cp->off_map),
        P.coef_dh, pp.nslots, P.lane_evals + 1);
      RGC_CUDA(cudaGetLastError());
      pair_final_kernel<<<(pp.nslots + kPWarps - 1) / kPWarps, kPThreads, 0, c.stream>>>(
        P.partials, pair_ctas, pp.nslots, P.slot_i, P.slot_f,
        reinterpret_cast<const int*>(cp->dev + cp->off_sub), P.extmask, P.coef_dh,
        reinterpret_cast<const double2*>(cp->dev + cp->off_vs), d_msum, pp.nb, d_out);
      RGC_CUDA(cudaGetLastError());
      count_launch(6);
      RGC_TRY(launch_scatter_add(d_out, reinterpret_cast<const int*>(cp->dev + cp->off_map), pp.nslots,
                                 d_acc));
      if (defer_sync && n <= chunk_max) {
        // single pass: the caller synchronises once and then collects the times
        c.last_ms[2] = c.last_ms[3] = 0.f;
        return RGC_OK;
      }
      // the events (and the staged arrays) are reused by the next pass: wait here
      RGC_CUDA(cudaStreamSynchronize(c.stream));
      float ms = 0.f;
      RGC_CUDA(cudaEventElapsedTime(&ms, c.ev[5], c.ev[4]));
      if (main_ms) {
        *main_ms += ms;
      }
      RGC_CUDA(cudaEventElapsedTime(&ms, c.ev[2], c.ev[3]));
      pro_ms += ms;
      RGC_CUDA(cudaEventElapsedTime(&ms, c.ev[3], c.ev[5]));
      sort_ms += ms;
    }
    c.last_ms[2] = pro_ms;
    c.last_ms[3] = sort_ms;
    return RGC_OK;
  }

  int collect_pair_times(float* main_ms) {
    auto& c  = ctx();
    float ms = 0.f;
    RGC_CUDA(cudaEventElapsedTime(&ms, c.ev[5], c.ev[4]));
    if (main_ms) {
      *main_ms += ms;
    }
    RGC_CUDA(cudaEventElapsedTime(&c.last_ms[2], c.ev[2], c.ev[3]));
    RGC_CUDA(cudaEventElapsedTime(&c.last_ms[3], c.ev[3], c.ev[5]));
    return RGC_OK;
  }

} // namespace rgc
