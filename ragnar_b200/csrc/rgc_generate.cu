// Synthetic particle populations generated on the device (bench / test inputs
// for particle counts that cannot be staged through host memory).
// Counter-based Philox4x32-10 keyed by (seed, global particle index): particle
// g is the same numbers whichever GPU, shard size or launch shape produces it,
// so G-way sharded runs see exactly the particles of the 1-GPU run.
// The distributions mirror the reference's own particle test,
// src/tests/synchrotron.py:62-87 (power-law |U|, isotropic unit B).
#include "rgc_internal.hpp"

#include <algorithm>

namespace rgc {

  struct GenParams {
    float*        u[3];
    float*        e[3];
    float*        b[3];
    std::size_t   start, n;
    std::uint64_t seed, global_offset;
    int           kind;
    float         umin, umax;
  };

  __device__ __forceinline__ void philox_round(uint4& c, uint2& k) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }

  __device__ __forceinline__ uint4 philox4x32_10(std::uint64_t index, unsigned stream,
                                                 std::uint64_t seed) {
    uint4 c = make_uint4((unsigned)index, (unsigned)(index >> 32), stream, 0u);
    uint2 k = make_uint2((unsigned)seed, (unsigned)(seed >> 32));
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      philox_round(c, k);
    }
    return c;
  }

  // uniform in (0, 1), 24 random bits
  __device__ __forceinline__ float u01(unsigned x) {
    return (float)(x >> 8) * 5.9604645e-8f + 2.9802322e-8f;
  }

  __device__ __forceinline__ float3 isotropic(float r1, float r2) {
    const float mu = 2.0f * r1 - 1.0f;
    const float st = sqrtf(fmaxf(0.0f, 1.0f - mu * mu));
    float       s, c;
    sincospif(2.0f * r2, &s, &c);
    return make_float3(st * c, st * s, mu);
  }

  // dN/du ~ u^-2 on [umin, umax] by inverse CDF
  __device__ __forceinline__ float plaw_m2(float r, float umin, float umax) {
    const float a = 1.0f / umin, bq = 1.0f / umax;
    return 1.0f / (a + (bq - a) * r);
  }

  __global__ void generate_kernel(const GenParams P) {
    const std::size_t stride = (std::size_t)gridDim.x * blockDim.x;
    for (std::size_t i = (std::size_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += stride) {
      const std::uint64_t gidx = P.global_offset + i;
      const std::size_t   o    = P.start + i;
      const uint4         ra   = philox4x32_10(gidx, 0u, P.seed);
      const float         un   = plaw_m2(u01(ra.x), P.umin, P.umax);
      float3              U, E = make_float3(0.f, 0.f, 0.f), B = make_float3(0.f, 0.f, 0.f);
      if (P.kind == 0) {
        U = make_float3(un, 0.f, 0.f);
        B = isotropic(u01(ra.y), u01(ra.z));
      } else if (P.kind == 1) {
        const uint4  rb = philox4x32_10(gidx, 1u, P.seed);
        const float3 nu = isotropic(u01(ra.y), u01(ra.z));
        U               = make_float3(un * nu.x, un * nu.y, un * nu.z);
        const float  bn = 0.5f + 1.5f * u01(ra.w);
        const float3 nb = isotropic(u01(rb.x), u01(rb.y));
        B               = make_float3(bn * nb.x, bn * nb.y, bn * nb.z);
        const float3 nr = isotropic(u01(rb.z), u01(rb.w));
        E = make_float3(0.1f * (B.y * nr.z - B.z * nr.y), 0.1f * (B.z * nr.x - B.x * nr.z),
                        0.1f * (B.x * nr.y - B.y * nr.x));
      } else {
        const float3 nu = isotropic(u01(ra.y), u01(ra.z));
        U               = make_float3(un * nu.x, un * nu.y, un * nu.z);
      }
      P.u[0][o] = U.x;
      P.u[1][o] = U.y;
      P.u[2][o] = U.z;
      P.e[0][o] = E.x;
      P.e[1][o] = E.y;
      P.e[2][o] = E.z;
      P.b[0][o] = B.x;
      P.b[1][o] = B.y;
      P.b[2][o] = B.z;
    }
  }

} // namespace rgc

using namespace rgc;

extern "C" int rgc_particles_generate(rgc_particles_t* p, int kind, uint64_t seed,
                                      uint64_t global_offset, size_t start, size_t n,
                                      float umin, float umax) {
  RGC_REQUIRE_INIT();
  if (!p || !p->allocated) {
    return fail(RGC_ERR_INVALID, "Particles not allocated");
  }
  if (kind < 0 || kind > 2) {
    return fail(RGC_ERR_INVALID, "unknown synthetic population %d", kind);
  }
  if (start + n > p->nalloc) {
    return fail(RGC_ERR_INVALID, "generate [%zu, %zu) exceeds allocation %zu", start, start + n,
                p->nalloc);
  }
  if (!(umin > 0.0f) || !(umax > umin)) {
    return fail(RGC_ERR_INVALID, "need 0 < umin < umax");
  }
  if (n == 0) {
    return RGC_OK;
  }
  GenParams P {};
  for (int d = 0; d < 3; ++d) {
    P.u[d] = p->col[RGC_Q_U][d];
    P.e[d] = p->col[RGC_Q_E][d];
    P.b[d] = p->col[RGC_Q_B][d];
  }
  P.start         = start;
  P.n             = n;
  P.seed          = seed;
  P.global_offset = global_offset;
  P.kind          = kind;
  P.umin          = umin;
  P.umax          = umax;
  const int blocks = (int)std::min<std::size_t>((n + 255) / 256, (std::size_t)ctx().sm_count * 16);
  generate_kernel<<<blocks, 256, 0, ctx().stream>>>(P);
  RGC_CUDA(cudaGetLastError());
  count_launch(1);
  return RGC_OK;
}
