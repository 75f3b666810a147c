// Synchrotron spectrum entry points, the FromDist kernel and the gather fallback for
// sm_100a.  (The main path of the particle spectrum is the bucketed hinge pipeline of
// rgc_sync_pair.cu; run_spectrum below routes every eligible bin chunk to it and
// keeps the gather kernel of this file for tables that do not vanish at both ends and
// for bin sets spanning more than ~1000 table cells.  SynchrotronSpectrumFromDist is
// the small fp64 kernel sync_dist_kernel.)
//
// Replaces (reference paths relative to haykh/ragnar @ fceb6b08):
//   sync::Kernel<D>::operator() / OmegaSync_ChiR   src/physics/synchrotron.hpp:145-232
//   sync::KernelFromDist::operator()               src/physics/synchrotron.hpp:72-95
//   InterpolateTabulatedFunction<true>             src/containers/tabulation.hpp:19-42
//   SynchrotronSpectrum<D> / ...FromDist drivers   src/physics/synchrotron.cpp:69-145
//
// What one (particle i, photon bin j) evaluation is in the reference:
//     x0 = e_syn[j] / e_peak_i;  F = loglog-table(x0);  spec[j] += e_syn[j] * chiR_i * F
// with ~5 log10f + 5 divisions per pair and (chiR_i, e_peak_i) recomputed per bin.
//
// Gather kernel (same numbers, different arithmetic):
//   * the table is linear in t = (log10 x0 - log10 x[0]) / dL between its nodes,
//     and t = a_j + c_i with a_j = (log10 e_syn[j] - log10 x[0]) / dL per bin and
//     c_i = -log10(e_peak_i) / dL per particle.  Both are computed ONCE in fp64 and
//     stored as unsigned 12.20 fixed point, so t is one exact integer add per pair
//     (resolution 2^-20 of a table cell; the reference's own float index is only
//     good to ~1e-5 cell).
//   * cell = t >> 20 indexes a zero-padded table of per-cell lines held in shared
//     memory as float2 (A_k, B_k) with F = A_k + B_k * m, m = 1 + frac/8 built
//     by OR-ing the 20 fraction bits under a 1.0f exponent (no int->float
//     conversion).  Lines are built in fp64 from the table's ACTUAL node
//     positions log10(x[k]), so they are the reference's interpolant.
//     Out-of-table pairs land in the zero padding: no per-pair range checks.
//   * lanes own photon bins (32 consecutive bins per group, GPW groups per warp,
//     accumulators in registers); particles are broadcast from shared memory.
//     A pair costs IMAD + SHF + LOP3 + LOP3 + LDS.64 + 2 FFMA and no atomics.
//   * per-particle prologue (gamma, beta, chiR, e_peak: reference's fp64
//     promotions, rounded to float exactly where the reference rounds) runs once
//     per particle per launch, 4 particles per thread from 16-byte column loads.
//   * float accumulators are flushed into fp64 registers every tile (<=128 terms
//     each), CTA partials are summed in a fixed order by a second kernel:
//     deterministic results, fp64 accumulation (the reference accumulates in
//     float; parity is defined against its float terms summed wide).
//
// Roofline: 8 B of shared-memory gather per evaluation against
// 128 B/clk/SM, and 7 issue slots per 32 evaluations against 4 issue/clk/SM.
// HBM: 36 B per particle amortised over nbins evaluations.
//
// Compiled with -fmad=false: every FMA below is an explicit fmaf()/fma(), all
// other float/double arithmetic is unfused like the reference's default build.
#include "rgc_hist_device.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace rgc {

  constexpr int      kThreads  = 256;
  constexpr int      kWarps    = kThreads / 32;
  constexpr int      kTile     = 1024; // particles per CTA tile (4 per thread)
  constexpr int      kFracBits = 20;
  constexpr unsigned kFracMask = (1u << kFracBits) - 1u;
  constexpr int      kMaxPad   = 4096; // table entries addressable by 12 index bits
  constexpr int      kMaxGroups = 64;  // 8 warp columns x 8 groups = 2048 bins per launch

  struct SpectrumParams {
    // particle columns
    const float* u[3];
    const float* e[3];
    const float* b[3];
    std::size_t  nprtl;
    // per-bin fixed-point coordinates, padded to ngroups * 32
    const unsigned* a_fx;
    // zero-padded table of per-cell lines
    const float2*   table;
    int             n_pad;
    int             ncols;   // warp columns (1, 2, 4 or 8)
    int             one;     // == 1, opaque to the compiler (keeps the t-add on the FMA pipe)
    unsigned        exp_one; // == 0x3f800000, opaque to the compiler
    unsigned        c_pad;   // an in-table particle coordinate for zero-weight padding
    // prologue constants
    float  B0, g_syn, e_at;
    double c0;        // amin + pad_lo
    double inv_dL;    // 1 / dL
    double c_lo, c_hi; // in-range window of the shifted coordinate
    // output: partials[cta][ngroups * 32]
    double* partials;
    int     nbins_pad;
    int*    poison; // raised when a particle's chiR is +inf (every bin becomes NaN)
  };

  __device__ __forceinline__ int2 particle_prologue(const SpectrumParams& P, float ux,
                                                    float uy, float uz, float ex, float ey,
                                                    float ez, float bx, float by, float bz) {
    // reference src/physics/synchrotron.hpp:193-231; float products promoted to
    // double exactly where `1.0 + ux * ux + ...` promotes them
    const double gamma  = sqrt(((1.0 + (double)(ux * ux)) + (double)(uy * uy)) + (double)(uz * uz));
    const double beta_x = (double)ux / gamma;
    const double beta_y = (double)uy / gamma;
    const double beta_z = (double)uz / gamma;
    const double bde    = (beta_x * ex + beta_y * ey) + beta_z * ez;
    const double cx     = beta_y * bz - beta_z * by;
    const double cy     = beta_z * bx - beta_x * bz;
    const double cz     = beta_x * by - beta_y * bx;
    const double sx = ex + cx, sy = ey + cy, sz = ez + cz;
    const double ssq  = (sx * sx + sy * sy) + sz * sz;
    const float  chiR = (float)(sqrt(ssq - bde * bde) / (double)P.B0);
    const float  e_peak =
      (float)((((double)P.e_at * gamma) * gamma) * (double)chiR / (double)(P.g_syn * P.g_syn));
    int2 out = make_int2(0, 0);
    // reference synchrotron.hpp:162 `if (e_peak > 0.0)`; +inf passes there but
    // gives x0 = 0 < xmin, i.e. F = yfill = 0: nothing for a finite chiR, and the term
    // e_syn * inf * 0 = NaN in every bin for chiR = +inf
    if (chiR == __int_as_float(0x7f800000) && e_peak == __int_as_float(0x7f800000)) {
      atomicAdd(P.poison, 1);
    }
    if (e_peak > 0.0f && e_peak < __int_as_float(0x7f800000)) {
      const double c = P.c0 - log10((double)e_peak) * P.inv_dL;
      if (c >= P.c_lo && c < P.c_hi) {
        out.x = (int)(unsigned)__double2ll_rn(c * (double)(1u << kFracBits));
        out.y = __float_as_int(chiR);
      }
    }
    return out;
  }

  template <int GPW>
  __global__ void __launch_bounds__(kThreads, 2)
    sync_spectrum_kernel(const __grid_constant__ SpectrumParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* tab = reinterpret_cast<float2*>(smem_raw);
    int2*   cw  = reinterpret_cast<int2*>(smem_raw + (std::size_t)P.n_pad * sizeof(float2));
    __shared__ int warp_valid[kWarps];

    const int tid  = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int col  = warp % P.ncols;
    const int row  = warp / P.ncols;
    const int rows = kWarps / P.ncols;

    for (int i = tid; i < P.n_pad; i += kThreads) {
      tab[i] = P.table[i];
    }

    unsigned a[GPW];
    float    acc[GPW];
    double   accd[GPW];
#pragma unroll
    for (int g = 0; g < GPW; ++g) {
      a[g]    = P.a_fx[(col * GPW + g) * 32 + lane];
      acc[g]  = 0.0f;
      accd[g] = 0.0;
    }

    // opaque to the compiler: `one` keeps the per-pair add an IMAD (FMA pipe,
    // the ALU pipe already carries the shift and the two logic ops), `exp_one`
    // in a register lets (t & mask) | 0x3f800000 be a single LOP3
    const unsigned    one     = (unsigned)P.one;
    const unsigned    exp_one = P.exp_one;
    const std::size_t ntiles  = (P.nprtl + kTile - 1) / kTile;

    for (std::size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const std::size_t base = tile * kTile;
      // ---- phase 1: per-particle prologue, 4 consecutive particles per thread
      const std::size_t i0 = base + (std::size_t)tid * 4;
      int2              r[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        r[k] = make_int2(0, 0);
      }
      if (i0 < P.nprtl) {
        float4 v[9];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          v[d]     = *reinterpret_cast<const float4*>(P.u[d] + i0);
          v[3 + d] = *reinterpret_cast<const float4*>(P.e[d] + i0);
          v[6 + d] = *reinterpret_cast<const float4*>(P.b[d] + i0);
        }
        const float* f = reinterpret_cast<const float*>(v);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (i0 + k < P.nprtl) {
            r[k] = particle_prologue(P, f[0 * 4 + k], f[1 * 4 + k], f[2 * 4 + k],
                                     f[3 * 4 + k], f[4 * 4 + k], f[5 * 4 + k],
                                     f[6 * 4 + k], f[7 * 4 + k], f[8 * 4 + k]);
          }
        }
      }
      // compact the particles that reach at least one bin (weight bits != 0),
      // in particle order: ballots within the warp, warp totals across the CTA
      int      my_off = 0, warp_total = 0;
      unsigned keep   = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool     v = r[k].y != 0;
        const unsigned m = __ballot_sync(0xffffffffu, v);
        if (v) {
          keep |= 1u << k;
        }
        // position among this warp's survivors ordered by (lane, k)
        warp_total += __popc(m);
        my_off += __popc(m & ((1u << lane) - 1u));
      }
      // (lane, k) order: survivors of lower lanes first, then own lower k
      // my_off so far counts lower lanes over all k; that is exactly the rank of
      // this lane's first survivor
      if (lane == 0) {
        warp_valid[warp] = warp_total;
      }
      __syncthreads(); // also: previous tile's phase 2 is complete
      int warp_base = 0, nvalid = 0;
#pragma unroll
      for (int wq = 0; wq < kWarps; ++wq) {
        const int c = warp_valid[wq];
        warp_base += wq < warp ? c : 0;
        nvalid += c;
      }
      {
        int o = warp_base + my_off;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (keep & (1u << k)) {
            cw[o++] = r[k];
          }
        }
      }
      // rows take equal even-sized shares; the shares are padded with zero-weight
      // entries at an in-table coordinate
      const int per_row = (((nvalid + rows - 1) / rows) + 1) & ~1;
      if (tid < per_row * rows - nvalid) {
        cw[nvalid + tid] = make_int2((int)P.c_pad, 0);
      }
      __syncthreads();
      // ---- phase 2: every warp row sweeps its share of the tile over its bins
      const int2* mine = cw + row * per_row;
      for (int p = 0; p < per_row; p += 2) {
        const int4     pc = *reinterpret_cast<const int4*>(mine + p); // broadcast
        const unsigned c0 = (unsigned)pc.x, c1 = (unsigned)pc.z;
        const float    w0 = __int_as_float(pc.y), w1 = __int_as_float(pc.w);
#pragma unroll
        for (int g = 0; g < GPW; ++g) {
          const unsigned t0 = a[g] * one + c0;
          const unsigned t1 = a[g] * one + c1;
          const float2   l0 = tab[t0 >> kFracBits];
          const float2   l1 = tab[t1 >> kFracBits];
          const float    m0 = __uint_as_float((t0 & kFracMask) | exp_one);
          const float    m1 = __uint_as_float((t1 & kFracMask) | exp_one);
          acc[g]            = fmaf(w0, fmaf(m0, l0.y, l0.x), acc[g]);
          acc[g]            = fmaf(w1, fmaf(m1, l1.y, l1.x), acc[g]);
        }
      }
#pragma unroll
      for (int g = 0; g < GPW; ++g) {
        accd[g] += (double)acc[g];
        acc[g] = 0.0f;
      }
      __syncthreads(); // phase 2 done before the next tile's survivors overwrite cw
    }

    // ---- CTA reduction over warp rows (fixed order), one partial row per CTA
    __syncthreads();
    double* red = reinterpret_cast<double*>(cw); // kWarps * GPW * 32 doubles <= 16 KB
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
        P.partials[(std::size_t)blockIdx.x * P.nbins_pad + (col * GPW + g) * 32 + lane] = s;
      }
    }
  }

  // sums the per-CTA partial rows in CTA order: out[j] (+)= sum_cta partials[cta][j]
  __global__ void reduce_partials_kernel(const double* __restrict__ partials, int nctas,
                                         int nbins_pad, int nbins, double* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nbins) {
      return;
    }
    double s = 0.0;
    for (int c = 0; c < nctas; ++c) {
      s += partials[(std::size_t)c * nbins_pad + j];
    }
    out[j] = s;
  }

  // d_acc[binmap[s]] += src[s]: folds one launch's slot sums into the per-bin result
  // on the device (launches of one call are stream-ordered; a bin belongs to exactly
  // one slot of one bin chunk, so there is no concurrent update)
  __global__ void scatter_add_kernel(const double* __restrict__ src, const int* __restrict__ binmap,
                                     int nslots, double* __restrict__ d_acc) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < nslots) {
      const int b = binmap[s];
      if (b >= 0) {
        d_acc[b] += src[s];
      }
    }
  }

  // d_acc[j] = e_syn[j] * sum_i w_i F_ij — the factor every term of a bin shares, applied
  // before the all-reduce so that every rank contributes finished values whichever path
  // it took (the literal path's terms already carry it).  A poisoned population turns
  // every bin into NaN, as in the reference.
  __global__ void finalize_acc_kernel(const int* __restrict__ poison, int nbins,
                                      const float* __restrict__ e_syn, double* __restrict__ d_acc) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < nbins) {
      d_acc[j] = *poison != 0 ? __longlong_as_double(0x7ff8000000000000ll)
                              : (double)e_syn[j] * d_acc[j];
    }
  }

  int launch_scatter_add(const double* src, const int* binmap, int nslots, double* d_acc) {
    scatter_add_kernel<<<(nslots + 127) / 128, 128, 0, ctx().stream>>>(src, binmap, nslots, d_acc);
    RGC_CUDA(cudaGetLastError());
    count_launch(1);
    return RGC_OK;
  }

  // ------------------------------------------------------------------ host side
  static int make_table_plan(const float* tab_x, const float* tab_y, std::size_t T,
                             TablePlan& tp) {
    if (T < 2) {
      return fail(RGC_ERR_INVALID, "F table needs at least 2 points");
    }
    for (std::size_t k = 0; k < T; ++k) {
      if (!(tab_x[k] > 0.0f) || !std::isfinite(tab_x[k]) || (k > 0 && !(tab_x[k] > tab_x[k - 1]))) {
        return fail(RGC_ERR_INVALID, "xmin <= 0.0 in Logspace TabulatedFunction");
      }
    }
    tp.T  = T;
    tp.L0 = std::log10((double)tab_x[0]);
    tp.dL = (std::log10((double)tab_x[T - 1]) - tp.L0) / (double)(T - 1);
    tp.tx.resize(T);
    tp.y.resize(T);
    for (std::size_t k = 0; k < T; ++k) {
      tp.tx[k] = (std::log10((double)tab_x[k]) - tp.L0) / tp.dL;
      tp.y[k]  = (double)tab_y[k];
      // the reference indexes the table as a uniform log grid
      // (tabulation.hpp:33-35); nodes must sit within a small fraction of a cell
      // of their nominal position for the per-cell lines to be its interpolant
      if (std::fabs(tp.tx[k] - (double)k) > 0.05) {
        return fail(RGC_ERR_INVALID, "F table is not a uniform logarithmic grid");
      }
    }
    return RGC_OK;
  }

  struct LaunchPlan {
    std::vector<unsigned> a_fx;  // ngroups * 32
    std::vector<float2>   table; // n_pad
    std::vector<int>      bin_of_slot; // slot -> original bin index (or -1)
    int    ngroups { 0 }, gpw { 0 }, ncols { 0 }, n_pad { 0 };
    double c0 { 0 }, c_lo { 0 }, c_hi { 0 };
  };

  // bins: indices (into bins_e_syn) handled by this launch, all with e_syn > 0 finite
  static int make_launch_plan(const TablePlan& tp, const float* bins_e_syn,
                              const std::vector<int>& bins, LaunchPlan& lp) {
    const int nb = (int)bins.size();
    std::vector<double> a(nb);
    double amin = 1e300, amax = -1e300;
    for (int s = 0; s < nb; ++s) {
      a[s] = (std::log10((double)bins_e_syn[bins[s]]) - tp.L0) / tp.dL;
      amin = std::min(amin, a[s]);
      amax = std::max(amax, a[s]);
    }
    const double spread = amax - amin;
    const int    pad_lo = (int)std::ceil(spread) + 1;
    lp.n_pad            = (pad_lo + (int)tp.T + (int)std::ceil(spread) + 2 + 1) & ~1; // even: int4 stores follow
    if (lp.n_pad > kMaxPad) {
      return fail(RGC_ERR_INVALID, "internal: bin chunk spans too many table cells (%d)", lp.n_pad);
    }
    lp.ngroups = (nb + 31) / 32;
    if (lp.ngroups <= 8) {
      lp.ncols = 1;
    } else if (lp.ngroups <= 16) {
      lp.ncols = 2;
    } else if (lp.ngroups <= 32) {
      lp.ncols = 4;
    } else {
      lp.ncols = 8;
    }
    lp.gpw     = (lp.ngroups + lp.ncols - 1) / lp.ncols;
    lp.ngroups = lp.gpw * lp.ncols;
    lp.a_fx.assign((std::size_t)lp.ngroups * 32, 0u);
    lp.bin_of_slot.assign((std::size_t)lp.ngroups * 32, -1);
    const double scale = (double)(1u << kFracBits);
    for (int s = 0; s < nb; ++s) {
      lp.a_fx[s]        = (unsigned)std::llrint((a[s] - amin) * scale);
      lp.bin_of_slot[s] = bins[s];
    }
    // per-cell lines F = A + B * m, m = 1 + frac/8 in [1, 1.125)
    lp.table.assign(lp.n_pad, make_float2(0.f, 0.f));
    for (std::size_t k = 0; k + 1 < tp.T; ++k) {
      const double slope = (tp.y[k + 1] - tp.y[k]) / (tp.tx[k + 1] - tp.tx[k]);
      const double A     = tp.y[k] + slope * ((double)k - 8.0 - tp.tx[k]);
      const double B     = 8.0 * slope;
      lp.table[pad_lo + k] = make_float2((float)A, (float)B);
    }
    // t = a_j + c, c = -log10(e_peak)/dL; shifted coordinate c' = c + amin + pad_lo
    lp.c0   = amin + (double)pad_lo;
    lp.c_lo = (double)pad_lo - spread;               // c >= -amax
    lp.c_hi = (double)(tp.T - 1) + (double)pad_lo;   // c <  T-1-amin
    return RGC_OK;
  }

  static void launch_spectrum(int gpw, dim3 grid, std::size_t smem, cudaStream_t st,
                              const SpectrumParams& P) {
    switch (gpw) {
#define RGC_CASE(G)                                                                   \
  case G:                                                                             \
    sync_spectrum_kernel<G><<<grid, kThreads, smem, st>>>(P);                         \
    break;
      RGC_CASE(1)
      RGC_CASE(2)
      RGC_CASE(3)
      RGC_CASE(4)
      RGC_CASE(5)
      RGC_CASE(6)
      RGC_CASE(7)
      RGC_CASE(8)
#undef RGC_CASE
    }
  }

  // splits the valid photon bins into launches of <= 2048 bins whose spread in
  // table cells keeps the padded table within kMaxPad entries
  static void chunk_bins(const TablePlan& tp, const float* bins_e_syn, std::size_t nbins,
                         std::vector<std::vector<int>>& chunks, std::vector<int>& nan_bins) {
    std::vector<std::pair<double, int>> valid;
    for (std::size_t j = 0; j < nbins; ++j) {
      const float e = bins_e_syn[j];
      if (std::isnan(e)) {
        nan_bins.push_back((int)j);
      } else if (e > 0.0f && std::isfinite(e)) {
        valid.emplace_back(std::log10((double)e) / tp.dL, (int)j);
      }
    }
    // keep the caller's order when it already fits one launch (the common case:
    // ascending bins, adjacent lanes hit adjacent table cells)
    auto fits = [&](double lo, double hi, std::size_t count) {
      const double spread = hi - lo;
      return count <= (std::size_t)kMaxGroups * 32 &&
             2.0 * std::ceil(spread) + (double)tp.T + 3.0 <= (double)kMaxPad;
    };
    if (valid.empty()) {
      return;
    }
    double lo = 1e300, hi = -1e300;
    for (auto& v : valid) {
      lo = std::min(lo, v.first);
      hi = std::max(hi, v.first);
    }
    if (fits(lo, hi, valid.size())) {
      chunks.emplace_back();
      for (auto& v : valid) {
        chunks.back().push_back(v.second);
      }
      return;
    }
    std::stable_sort(valid.begin(), valid.end(),
                     [](const auto& x, const auto& y) { return x.first < y.first; });
    std::size_t s = 0;
    while (s < valid.size()) {
      std::size_t e = s + 1;
      while (e < valid.size() && fits(valid[s].first, valid[e].first, e - s + 1)) {
        ++e;
      }
      chunks.emplace_back();
      for (std::size_t k = s; k < e; ++k) {
        chunks.back().push_back(valid[k].second);
      }
      s = e;
    }
  }

  struct SpectrumSource {
    const rgc_particles_t* prtls { nullptr }; // particles path
    std::size_t            n { 0 };
    float                  B0 { 1 }, g_syn { 1 }, e_at { 1 };
  };

  // Runs all launches, leaves acc64[nbins] = sum_i w_i F_ij (before the e_syn factor)
  // on the device (all-reduced when requested) and copies it to the host.
  static int run_spectrum(const SpectrumSource& src, const float* bins_e_syn,
                          std::size_t nbins, const float* tab_x, const float* tab_y,
                          std::size_t tab_n, bool allreduce, std::vector<double>& acc_host) {
    auto& c = ctx();
    TablePlan tp;
    RGC_TRY(make_table_plan(tab_x, tab_y, tab_n, tp));
    std::vector<std::vector<int>> chunks;
    std::vector<int>              nan_bins;
    chunk_bins(tp, bins_e_syn, nbins, chunks, nan_bins);

    acc_host.assign(nbins, 0.0);
    if (nbins == 0) {
      return RGC_OK; // an empty Bins gives an empty Array1D (every rank alike)
    }
    // per-bin sums are accumulated on the device (one small buffer that survives the
    // scratch re-layouts of the launches), all-reduced once and read back once:
    // [nbins doubles | poison flag, issued evaluations, on-table pairs | nbins floats: e_syn]
    void* result = nullptr;
    RGC_TRY(ensure_result(nbins * sizeof(double) + 24 + nbins * sizeof(float), &result));
    double* d_acc    = static_cast<double*>(result);
    int*    d_poison = reinterpret_cast<int*>(d_acc + nbins);
    float*  d_bins   = reinterpret_cast<float*>(d_acc + nbins + 3);
    RGC_CUDA(cudaEventRecord(c.ev[0], c.stream));
    RGC_CUDA(cudaMemsetAsync(d_acc, 0, nbins * sizeof(double) + 24, c.stream));
    float main_ms  = 0.f;
    bool  deferred = false;
    // ---- small populations: the reference's own float arithmetic per pair
    // (rgc_sync_literal.cu); nothing averages its rounding there
    const bool literal = src.n > 0 && src.n <= literal_max_n();
    if (literal) {
      RGC_TRY(run_spectrum_literal(src.prtls, src.n, src.B0, src.g_syn, src.e_at, nullptr, nullptr,
                                   nullptr, bins_e_syn, nbins, tab_x, tab_y, tab_n, d_acc));
      chunks.clear();
      nan_bins.clear();
    } else {
      RGC_TRY(copy_h2d(d_bins, bins_e_syn, nbins * sizeof(float), c.stream));
    }
    // ---- bucketed hinge path (rgc_sync_pair.cu) for every chunk it can take;
    // RGC_SPECTRUM_PATH=gather forces the gather kernel below (A/B checks)
    if (src.n == 0) {
      chunks.clear(); // a rank without particles launches nothing but still joins the all-reduce
    }
    if (src.n > 0 && !literal) {
      const char* force = std::getenv("RGC_SPECTRUM_PATH");
      const bool  allow = !(force && std::strcmp(force, "gather") == 0);
      std::vector<std::vector<int>> rest;
      int npair = 0;
      for (auto& chunk : chunks) {
        npair += (allow && pair_path_eligible(tp, bins_e_syn, chunk)) ? 1 : 0;
      }
      // one pair launch and nothing else: no intermediate synchronisation at all
      deferred = npair == 1 && chunks.size() == 1;
      for (auto& chunk : chunks) {
        if (allow && pair_path_eligible(tp, bins_e_syn, chunk)) {
          RGC_TRY(run_spectrum_pair(src.prtls, src.n, src.B0, src.g_syn, src.e_at, tp, bins_e_syn,
                                    chunk, d_acc, d_poison, &main_ms, deferred));
        } else {
          rest.push_back(std::move(chunk));
        }
      }
      chunks.swap(rest);
    }
    // device scratch: acc64[nbins] | per launch: a_fx, table, partials
    const int ctas_per_sm = 2;
    const std::size_t ntiles = (src.n + kTile - 1) / kTile;
    int nctas = (int)std::min<std::size_t>((std::size_t)c.sm_count * ctas_per_sm,
                                           std::max<std::size_t>(ntiles, 1));
    std::size_t max_slots = 32, max_pad = 1;
    std::vector<LaunchPlan> plans(chunks.size());
    for (std::size_t k = 0; k < chunks.size(); ++k) {
      RGC_TRY(make_launch_plan(tp, bins_e_syn, chunks[k], plans[k]));
      max_slots = std::max<std::size_t>(max_slots, plans[k].a_fx.size());
      max_pad   = std::max<std::size_t>(max_pad, plans[k].table.size());
    }
    auto align = [](std::size_t x) { return (x + 255) & ~std::size_t(255); };
    const std::size_t off_map   = 0;
    const std::size_t off_out   = align(off_map + max_slots * sizeof(int));
    const std::size_t off_afx   = align(off_out + max_slots * sizeof(double));
    const std::size_t off_table = align(off_afx + max_slots * sizeof(unsigned));
    const std::size_t off_part  = align(off_table + max_pad * sizeof(float2));
    const std::size_t total     = off_part + (std::size_t)nctas * max_slots * sizeof(double);
    void*             scratch   = nullptr;
    RGC_TRY(ensure_scratch(total, &scratch));
    char*   sbase   = static_cast<char*>(scratch);
    int*    d_map   = reinterpret_cast<int*>(sbase + off_map);
    double* d_out   = reinterpret_cast<double*>(sbase + off_out);
    auto*   d_afx   = reinterpret_cast<unsigned*>(sbase + off_afx);
    auto*   d_table = reinterpret_cast<float2*>(sbase + off_table);
    double* d_part  = reinterpret_cast<double*>(sbase + off_part);

    for (std::size_t k = 0; k < plans.size(); ++k) {
      const LaunchPlan& lp = plans[k];
      RGC_CUDA(cudaMemcpyAsync(d_afx, lp.a_fx.data(), lp.a_fx.size() * sizeof(unsigned),
                               cudaMemcpyHostToDevice, c.stream));
      RGC_CUDA(cudaMemcpyAsync(d_table, lp.table.data(), lp.table.size() * sizeof(float2),
                               cudaMemcpyHostToDevice, c.stream));
      SpectrumParams P {};
      for (int d = 0; d < 3; ++d) {
        P.u[d] = src.prtls->col[RGC_Q_U][d];
        P.e[d] = src.prtls->col[RGC_Q_E][d];
        P.b[d] = src.prtls->col[RGC_Q_B][d];
      }
      P.nprtl     = src.n;
      P.a_fx      = d_afx;
      P.table     = d_table;
      P.n_pad     = lp.n_pad;
      P.ncols     = lp.ncols;
      P.one       = 1;
      P.exp_one   = 0x3f800000u;
      P.c_pad     = (unsigned)std::llrint(lp.c_lo * (double)(1u << kFracBits)) + 1u;
      P.B0        = src.B0;
      P.g_syn     = src.g_syn;
      P.e_at      = src.e_at;
      P.c0        = lp.c0;
      P.inv_dL    = 1.0 / tp.dL;
      P.c_lo      = lp.c_lo;
      P.c_hi      = lp.c_hi;
      P.partials  = d_part;
      P.nbins_pad = lp.ngroups * 32;
      P.poison    = d_poison;
      const std::size_t smem = (std::size_t)lp.n_pad * sizeof(float2) +
                               std::max<std::size_t>((std::size_t)(kTile + 2 * kWarps) * sizeof(int2),
                                                     (std::size_t)kWarps * lp.gpw * 32 * sizeof(double));
      RGC_CUDA(cudaEventRecord(c.ev[2], c.stream));
      launch_spectrum(lp.gpw, dim3(nctas), smem, c.stream, P);
      RGC_CUDA(cudaGetLastError());
      RGC_CUDA(cudaEventRecord(c.ev[3], c.stream));
      const int nslots = lp.ngroups * 32;
      reduce_partials_kernel<<<(nslots + 127) / 128, 128, 0, c.stream>>>(d_part, nctas, nslots,
                                                                         nslots, d_out);
      RGC_CUDA(cudaGetLastError());
      count_launch(2);
      RGC_CUDA(cudaMemcpyAsync(d_map, lp.bin_of_slot.data(), nslots * sizeof(int),
                               cudaMemcpyHostToDevice, c.stream));
      RGC_TRY(launch_scatter_add(d_out, d_map, nslots, d_acc));
      // the plan's host arrays and the events are reused by the next launch
      RGC_CUDA(cudaStreamSynchronize(c.stream));
      float ms = 0.f;
      RGC_CUDA(cudaEventElapsedTime(&ms, c.ev[2], c.ev[3]));
      main_ms += ms;
    }
    if (!literal && src.n > 0) {
      finalize_acc_kernel<<<(unsigned)((nbins + 127) / 128), 128, 0, c.stream>>>(d_poison, (int)nbins,
                                                                                 d_bins, d_acc);
      RGC_CUDA(cudaGetLastError());
      count_launch(1);
    }
    if (allreduce) {
      RGC_TRY(allreduce_sum_f64(d_acc, nbins));
    }
    // one D2H: [per-bin sums | poison flag | issued hinge evaluations]
    std::vector<double> back(nbins + 3);
    RGC_CUDA(cudaMemcpyAsync(back.data(), d_acc, (nbins + 3) * sizeof(double),
                             cudaMemcpyDeviceToHost, c.stream));
    RGC_CUDA(cudaEventRecord(c.ev[1], c.stream));
    RGC_CUDA(cudaStreamSynchronize(c.stream));
    if (allreduce) {
      RGC_TRY(exchange_check());
    }
    std::copy(back.begin(), back.begin() + nbins, acc_host.begin());
    {
      unsigned long long le = 0;
      std::memcpy(&le, &back[nbins + 1], sizeof(le));
      c.last_lane_evals = (double)le;
      std::memcpy(&le, &back[nbins + 2], sizeof(le));
      c.last_ontable_evals = (double)le;
    }
    if (deferred && pair_single_pass(src.n)) {
      RGC_TRY(collect_pair_times(&main_ms));
    }
    if (literal) {
      RGC_CUDA(cudaEventElapsedTime(&main_ms, c.ev[2], c.ev[3]));
      c.last_ms[2] = c.last_ms[3] = 0.f;
    }
    float total_ms = 0.f;
    RGC_CUDA(cudaEventElapsedTime(&total_ms, c.ev[0], c.ev[1]));
    c.last_ms[0] = total_ms;
    c.last_ms[1] = main_ms;
    for (int j : nan_bins) {
      acc_host[j] = std::nan("");
    }
    return RGC_OK;
  }

  // acc holds finished per-bin values (e_syn factor applied on the device)
  static void finish_spectrum(const std::vector<double>& acc, std::size_t nbins, float* out_spec,
                              double* out_spec64) {
    for (std::size_t j = 0; j < nbins; ++j) {
      if (out_spec64) {
        out_spec64[j] = acc[j];
      }
      if (out_spec) {
        out_spec[j] = (float)acc[j];
      }
    }
  }

} // namespace rgc

using namespace rgc;

extern "C" {

  int rgc_sync_spectrum_particles(const rgc_particles_t* p, size_t nactive,
                                  const float* bins_e_syn, size_t nbins, const float* tab_x,
                                  const float* tab_y, size_t tab_n, float B0, float g_syn,
                                  float e_syn_at_g_syn, float* out_spec, double* out_spec64) {
    RGC_REQUIRE_INIT();
    RGC_NVTX("SynchrotronSpectrum");
    // an unallocated container is an empty shard (the reference launches 0 x M threads over
    // empty Views): it still joins a multi-rank all-reduce, so no rank is left waiting
    const bool empty = (!p || !p->allocated) && nactive == 0;
    if (!empty && (!p || !p->allocated)) {
      return fail(RGC_ERR_INVALID, "Particles not allocated");
    }
    if (!empty && nactive > p->nalloc) {
      return fail(RGC_ERR_INVALID, "nactive %zu exceeds allocation %zu", nactive, p->nalloc);
    }
    SpectrumSource src;
    src.prtls = p;
    src.n     = nactive;
    src.B0    = B0;
    src.g_syn = g_syn;
    src.e_at  = e_syn_at_g_syn;
    std::vector<double> acc;
    RGC_TRY(run_spectrum(src, bins_e_syn, nbins, tab_x, tab_y, tab_n, true, acc));
    finish_spectrum(acc, nbins, out_spec, out_spec64);
    return RGC_OK;
  }

  int rgc_hist_and_spectrum(const rgc_particles_t* p, size_t nactive, const float* gbins, size_t ng,
                            int log_spaced, int fourvel, float* out_hist, double* out_hist64,
                            const float* bins_e_syn, size_t nbins, const float* tab_x,
                            const float* tab_y, size_t tab_n, float B0, float g_syn,
                            float e_syn_at_g_syn, float* out_spec, double* out_spec64) {
    RGC_REQUIRE_INIT();
    RGC_NVTX("ComputeEnergyDistribution + SynchrotronSpectrum");
    if (!p || !p->allocated) {
      return fail(RGC_ERR_INVALID, "Particles not allocated");
    }
    if (nactive > p->nalloc) {
      return fail(RGC_ERR_INVALID, "nactive %zu exceeds allocation %zu", nactive, p->nalloc);
    }
    if (ng > 5000) {
      return fail(RGC_ERR_INVALID, "energy histogram supports at most 5000 bins (got %zu)", ng);
    }
    // The histogram is enqueued (kernel, fold, exchange) without a wait, the spectrum pipeline
    // runs right behind it on the same stream, and both results are collected after the
    // spectrum's one synchronisation: the same kernels as the two separate entry points,
    // bit-identical results, one host round trip less per species.  Its device arrays live in
    // the result buffer, past the spectrum's part (the scratch is re-laid-out by the pipeline).
    auto&   c = ctx();
    HistJob job;
    if (ng > 0) {
      const std::size_t spec_bytes = (nbins * sizeof(double) + 24 + nbins * sizeof(float) + 255) & ~std::size_t(255);
      void*             result     = nullptr;
      RGC_TRY(ensure_result(spec_bytes + hist_job_bytes(ng, c.sm_count), &result));
      RGC_TRY(hist_enqueue(p, nactive, gbins, ng, log_spaced != 0, fourvel != 0, false,
                           static_cast<char*>(result) + spec_bytes, job));
    }
    RGC_TRY(rgc_sync_spectrum_particles(p, nactive, bins_e_syn, nbins, tab_x, tab_y, tab_n, B0, g_syn,
                                        e_syn_at_g_syn, out_spec, out_spec64));
    if (ng > 0) {
      RGC_TRY(hist_collect(job, out_hist, nullptr, out_hist64));
    }
    return RGC_OK;
  }

  int rgc_last_pair_lane_evals(double* lane_evals) {
    if (lane_evals) {
      *lane_evals = ctx().last_lane_evals;
    }
    return RGC_OK;
  }

  int rgc_pair_plan_describe(const float* bins_e_syn, size_t nbins, const float* tab_x,
                             const float* tab_y, size_t tab_n, int info[8], float* phase,
                             int* slot_bin, size_t cap) {
    if (!bins_e_syn || !tab_x || !tab_y || !info) {
      return fail(RGC_ERR_INVALID, "rgc_pair_plan_describe: null argument");
    }
    TablePlan tp;
    RGC_TRY(make_table_plan(tab_x, tab_y, tab_n, tp));
    std::vector<int> idx;
    for (std::size_t j = 0; j < nbins; ++j) {
      if (bins_e_syn[j] > 0.0f && std::isfinite(bins_e_syn[j])) {
        idx.push_back((int)j);
      }
    }
    pair_plan_describe(tp, bins_e_syn, idx, info, phase, slot_bin, cap);
    return RGC_OK;
  }

  int rgc_sort_rank_mode(int* mode) {
    if (mode) {
      *mode = pair_rank_mode();
    }
    return RGC_OK;
  }

  int rgc_last_pair_ontable_evals(double* evals) {
    RGC_REQUIRE_INIT();
    if (evals) {
      *evals = ctx().last_ontable_evals;
    }
    return RGC_OK;
  }

  int rgc_sync_spectrum_dist_batch(const float* gbeta, const float* f, size_t nbatch, size_t ndist,
                                   int islog_bins_prtls, const float* bins_e_syn, size_t nbins,
                                   const float* tab_x, const float* tab_y, size_t tab_n,
                                   float g_syn, float e_syn_at_g_syn, int mode, float* out_spec,
                                   double* out_spec64) {
    RGC_REQUIRE_INIT();
    RGC_NVTX("SynchrotronSpectrum (FromDist)");
    if (nbins == 0 || nbatch == 0) {
      return RGC_OK;
    }
    if (ndist > (std::size_t)1 << 30 || nbins > (std::size_t)1 << 30) {
      return fail(RGC_ERR_INVALID, "SynchrotronSpectrumFromDist: grid too large");
    }
    if (mode < -1 || mode > 1) {
      return fail(RGC_ERR_INVALID, "rgc_sync_spectrum_dist_batch: mode must be -1, 0 or 1");
    }
    auto& c = ctx();
    if (tab_n < 2) {
      return fail(RGC_ERR_INVALID, "F table needs at least 2 points");
    }
    // measured on B200 (profiles/r2_fromdist_batch.txt): the kernel-matrix build costs one
    // literal pass over G x M pairs, so the contraction pays from ~64 distributions on (kernel time; 16384: 0.24 ms against 5.2 ms)
    const bool contract = mode == 1 || (mode == -1 && nbatch >= 64);
    // per distribution bin, on the host (ndist values): e_peak exactly as the reference
    // forms it in float (synchrotron.hpp:78-79); the term ((f * e_syn) [* gbeta]) * F is
    // formed per pair by the literal kernel (synchrotron.hpp:89-93)
    std::vector<float> ep(ndist);
    for (std::size_t g = 0; g < ndist; ++g) {
      const float gb = gbeta[g];
      ep[g]          = e_syn_at_g_syn * gb * gb / (g_syn * g_syn);
    }
    void* result = nullptr;
    RGC_TRY(ensure_result(nbatch * nbins * sizeof(double), &result));
    double* d_out = static_cast<double*>(result);
    RGC_CUDA(cudaEventRecord(c.ev[0], c.stream));
    if (ndist == 0) {
      RGC_CUDA(cudaMemsetAsync(d_out, 0, nbatch * nbins * sizeof(double), c.stream));
      RGC_CUDA(cudaEventRecord(c.ev[2], c.stream));
      RGC_CUDA(cudaEventRecord(c.ev[3], c.stream));
    } else {
      RGC_TRY(run_spectrum_literal(nullptr, ndist, 1.0f, g_syn, e_syn_at_g_syn, ep.data(), f,
                                   islog_bins_prtls ? gbeta : nullptr, bins_e_syn, nbins, tab_x,
                                   tab_y, tab_n, d_out, nbatch, contract));
    }
    std::vector<double> acc(nbatch * nbins);
    RGC_CUDA(cudaMemcpyAsync(acc.data(), d_out, acc.size() * sizeof(double), cudaMemcpyDeviceToHost,
                             c.stream));
    RGC_CUDA(cudaEventRecord(c.ev[1], c.stream));
    RGC_CUDA(cudaStreamSynchronize(c.stream));
    RGC_CUDA(cudaEventElapsedTime(&c.last_ms[0], c.ev[0], c.ev[1]));
    RGC_CUDA(cudaEventElapsedTime(&c.last_ms[1], c.ev[2], c.ev[3]));
    finish_spectrum(acc, nbatch * nbins, out_spec, out_spec64);
    return RGC_OK;
  }

  int rgc_sync_spectrum_dist(const float* gbeta, const float* f, size_t ndist,
                             int islog_bins_prtls, const float* bins_e_syn, size_t nbins,
                             const float* tab_x, const float* tab_y, size_t tab_n, float g_syn,
                             float e_syn_at_g_syn, float* out_spec, double* out_spec64) {
    return rgc_sync_spectrum_dist_batch(gbeta, f, 1, ndist, islog_bins_prtls, bins_e_syn, nbins, tab_x,
                                        tab_y, tab_n, g_syn, e_syn_at_g_syn, 0, out_spec, out_spec64);
  }

} // extern "C"
