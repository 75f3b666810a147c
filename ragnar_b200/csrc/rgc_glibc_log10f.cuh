// glibc's log10f, restated for the device so that the literal pair kernel
// (rgc_sync_literal.cu) forms the reference's float terms bit for bit.
//
// The reference evaluates InterpolateTabulatedFunction<true> with
// Kokkos::log10(float) = std::log10(float) = libm's log10f
// (haykh/ragnar @ fceb6b08, src/containers/tabulation.hpp:33-41).  In the image's
// glibc (2.39) that is the fdlibm wrapper  z = y*log10_2lo + ivln10*logf(m);
// return z + y*log10_2hi  around the table-driven logf of the ARM optimized
// routines (16-entry {1/c, log c} table, cubic in r = z/c - 1, double arithmetic,
// one rounding to float).  Published algorithm, constants read back from the image's
// libm.so.6 (`__logf_data`).  tools/check_log10f.c compares this restatement with
// the host's log10f / logf over ALL positive normal floats (0 mismatches, with and
// without FMA contraction of the double chain: the final float rounding absorbs it).
//
// __host__ __device__: the CPU test tests/test_log10f_cpu.py compiles the same
// header with g++ and checks it against libm on the host.
#ifndef RGC_GLIBC_LOG10F_CUH
#define RGC_GLIBC_LOG10F_CUH

#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define RGC_HD __host__ __device__ __forceinline__
#else
#define RGC_HD inline
#endif

namespace rgc {

  struct LogfEntry {
    double invc, logc;
  };

  // __logf_data.tab of glibc 2.39 (sysdeps/ieee754/flt-32/e_logf_data.c); device code
  // keeps its own copy (shared memory, filled from RGC_LOGF_TAB_INIT)
#define RGC_LOGF_TAB_INIT                                                                          \
  {                                                                                                \
    { 0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2 }, { 0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2 }, \
    { 0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2 },  { 0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3 }, \
    { 0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3 }, { 0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3 },    \
    { 0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4 }, { 0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4 }, \
    { 0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5 }, { 0x1p+0, 0x0p+0 },                              \
    { 0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5 },  { 0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4 },   \
    { 0x1.b2036576afce6p-1, 0x1.526e57720db08p-3 },  { 0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3 },   \
    { 0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2 },  { 0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2 },  \
  }

  RGC_HD std::uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    std::uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
#endif
  }
  RGC_HD float u2f(std::uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    std::memcpy(&f, &u, 4);
    return f;
#endif
  }

  // logf of a positive normal float (glibc's __logf without its special cases)
  RGC_HD float glibc_logf_normal(float x, const LogfEntry* tab) {
    const std::uint32_t ix = f2u(x);
    if (ix == 0x3f800000u) {
      return 0.0f;
    }
    const std::uint32_t tmp  = ix - 0x3f330000u;
    const int           i    = (int)((tmp >> 19) & 15u);
    const int           k    = (std::int32_t)tmp >> 23;
    const std::uint32_t iz   = ix - (tmp & (0x1ffu << 23));
    const double        invc = tab[i].invc, logc = tab[i].logc;
    const double        z  = (double)u2f(iz);
    const double        r  = z * invc - 1.0;
    const double        y0 = logc + (double)k * 0x1.62e42fefa39efp-1;
    const double        r2 = r * r;
    double              y  = 0x1.5575b0be00b6ap-2 * r + -0x1.ffffef20a4123p-2;
    y                      = -0x1.00ea348b88334p-2 * r2 + y;
    y                      = y * r2 + (y0 + r);
    return (float)y;
  }

  // log10f as glibc 2.39 computes it (sysdeps/ieee754/flt-32/e_log10f.c), every
  // argument class: zero -> -inf, negative -> NaN, subnormals scaled by 2^25, inf / NaN
  // returned as x + x.
  RGC_HD float glibc_log10f(float x, const LogfEntry* tab) {
    std::int32_t hx = (std::int32_t)f2u(x);
    std::int32_t k  = 0;
    if (hx < 0x00800000) { // x < 2^-126 (or negative: sign bit makes hx < 0)
      if ((hx & 0x7fffffff) == 0) {
        return -u2f(0x7f800000u); // log(+-0) = -inf
      }
      if (hx < 0) {
        return u2f(0x7fc00000u); // log(-#) = NaN
      }
      k -= 25;
      x *= 33554432.0f; // 2^25
      hx = (std::int32_t)f2u(x);
    }
    if (hx >= 0x7f800000) {
      return x + x;
    }
    k += (hx >> 23) - 127;
    const std::int32_t i = (std::int32_t)(((std::uint32_t)k & 0x80000000u) >> 31);
    hx                   = (hx & 0x007fffff) | ((0x7f - i) << 23);
    const float y        = (float)(k + i);
    const float m        = u2f((std::uint32_t)hx);
    const float z        = y * 7.9034151668e-07f + 4.3429449201e-01f * glibc_logf_normal(m, tab);
    return z + y * 3.0102920532e-01f;
  }

} // namespace rgc

#endif // RGC_GLIBC_LOG10F_CUH
