// Micro-benchmark: does a packed FFMA2 accumulate relieve the register-file read pressure
// of the pair loop?  Scalar hinge (FFMA.SAT / FADD.SAT) + packed accumulate against the
// all-scalar product form, 8 particles per loop iteration, NA = 8 lane groups.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -o tools/_build/mb_pair3 tools/mb_pair3.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

constexpr int kStageLen = 64;
typedef unsigned long long u64;

__device__ __forceinline__ u64 pack(float a, float b) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack(u64 v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// V = 0: scalar FFMA.SAT + scalar FFMA (product)
// V = 1: scalar FFMA.SAT + packed FFMA2
// V = 2: scalar FADD.SAT + packed FFMA2
// V = 3: packed FFMA2 only (2 per particle and pair)
// V = 4: scalar FFMA only (2 per particle and group)
// V = 5: V3 + one independent integer op per FFMA2 (does it hide in the second cycle?)
template <int V, int NA, int T>
__global__ void __launch_bounds__(T) k(float* out, const float4* __restrict__ prt, int stages) {
  __shared__ float4 ring[4 * kStageLen / 2 * (T / 32)];
  for (int i = threadIdx.x; i < 4 * kStageLen / 2 * (T / 32); i += T) {
    ring[i] = prt[i & 255];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  float fap[NA], sgn[NA], s2[NA];
  u64   s2p[NA / 2], xp[NA / 2];
  unsigned junk[4] = {threadIdx.x, 1u, 2u, 3u};
#pragma unroll
  for (int g = 0; g < NA; ++g) {
    fap[g] = (float)((threadIdx.x * NA + g) & 1023) * (1.0f / 1024.0f);
    sgn[g] = (threadIdx.x + g) & 64 ? 1.0f : -1.0f;
    s2[g]  = 0.f;
  }
#pragma unroll
  for (int h = 0; h < NA / 2; ++h) {
    s2p[h] = pack(0.f, 0.f);
    xp[h]  = pack(fap[2 * h], fap[2 * h + 1]);
  }
#define ONE(FC, W)                                                                       \
  {                                                                                      \
    if (V == 0) {                                                                        \
      float r[NA];                                                                       \
      _Pragma("unroll") for (int g = 0; g < NA; ++g) { r[g] = __saturatef(fmaf((FC), sgn[g], fap[g])); } \
      _Pragma("unroll") for (int g = 0; g < NA; ++g) { s2[g] = fmaf((W), r[g], s2[g]); } \
    } else if (V == 1 || V == 2) {                                                       \
      float r[NA];                                                                       \
      const u64 ww = pack((W), (W));                                                     \
      _Pragma("unroll") for (int g = 0; g < NA; ++g) {                                   \
        r[g] = V == 1 ? __saturatef(fmaf((FC), sgn[g], fap[g])) : __saturatef((FC) + fap[g]); \
      }                                                                                  \
      _Pragma("unroll") for (int h = 0; h < NA / 2; ++h) {                               \
        s2p[h] = fma2(ww, pack(r[2 * h], r[2 * h + 1]), s2p[h]);                         \
      }                                                                                  \
    } else if (V == 3 || V == 5) {                                                       \
      const u64 ww = pack((W), (W));                                                     \
      const u64 cc = pack((FC), (FC));                                                   \
      _Pragma("unroll") for (int h = 0; h < NA / 2; ++h) {                               \
        s2p[h] = fma2(ww, xp[h], s2p[h]);                                                \
        if (V == 5) { junk[h & 3] = junk[h & 3] * 3u + 1u; }                             \
        s2p[h] = fma2(cc, xp[h], s2p[h]);                                                \
        if (V == 5) { junk[(h + 2) & 3] ^= junk[(h + 2) & 3] >> 3; }                     \
      }                                                                                  \
    } else {                                                                             \
      _Pragma("unroll") for (int g = 0; g < NA; ++g) { s2[g] = fmaf((W), fap[g], s2[g]); } \
      _Pragma("unroll") for (int g = 0; g < NA; ++g) { s2[g] = fmaf((FC), fap[g], s2[g]); } \
    }                                                                                    \
  }
#define BODY(Q) ONE((Q).x, (Q).y) ONE((Q).z, (Q).w)
  for (int s = 0; s < stages; ++s) {
    const float4* buf = ring + (warp * 4 + (s & 3)) * (kStageLen / 2);
    float4 q0 = buf[0];
    float4 q1 = buf[1];
#pragma unroll 1
    for (int p = 0; p < kStageLen / 2 - 4; p += 4) {
      const float4 a0 = buf[p + 2];
      const float4 a1 = buf[p + 3];
      BODY(q0)
      BODY(q1)
      q0 = buf[p + 4];
      q1 = buf[p + 5];
      BODY(a0)
      BODY(a1)
    }
    {
      const float4 a0 = buf[kStageLen / 2 - 2];
      const float4 a1 = buf[kStageLen / 2 - 1];
      BODY(q0)
      BODY(q1)
      BODY(a0)
      BODY(a1)
    }
  }
  float s = 0.f;
#pragma unroll
  for (int g = 0; g < NA; ++g) s += s2[g];
#pragma unroll
  for (int h = 0; h < NA / 2; ++h) {
    float a, b;
    unpack(s2p[h], a, b);
    s += a + b;
  }
  if (s == 123.456f) out[0] = s + (float)(junk[0] + junk[1] + junk[2] + junk[3]);
}

template <int V, int NA, int T>
void run(const char* name, float* d, int sms, int ctas_per_sm) {
  const int stages = 2048, grid = sms * ctas_per_sm;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  double best = 0;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0));
    k<V, NA, T><<<grid, T>>>(d, reinterpret_cast<const float4*>(d) + 64, stages);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double lanes = 2.0 * NA * kStageLen * (double)stages * (double)grid * T / (ms * 1e-3);
    if (rep && lanes > best) best = lanes;
  }
  printf("%-44s NA=%d  %2d warps/SM  %6.2f TFLOP/s-equiv\n", name, NA, ctas_per_sm * T / 32, 2 * best / 1e12);
}

// the in-library peak kernel's form, for the same clocks
__global__ void __launch_bounds__(256) peak(float* out, int iters, float a, float b) {
  float acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = (float)(threadIdx.x + k);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = fmaf(acc[k], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += acc[k];
  if (s == 123.456f) out[0] = s;
}

int main() {
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  float* d; CK(cudaMalloc(&d, 1 << 20)); CK(cudaMemset(d, 0, 1 << 20));
  {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
      CK(cudaEventRecord(e0));
      peak<<<sms * 8, 256>>>(d, 65536, 0.999f, 0.001f);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      double f = 2.0 * 16 * 65536.0 * sms * 8 * 256 / (ms * 1e-3);
      if (rep && f > best) best = f;
    }
    printf("peak FFMA (acc = fma(acc, a, b))  %6.2f TFLOP/s\n", best / 1e12);
  }
  run<0, 8, 256>("V0 FFMA.SAT + FFMA (product)", d, sms, 2);
  run<1, 8, 256>("V1 FFMA.SAT + FFMA2", d, sms, 2);
  run<2, 8, 256>("V2 FADD.SAT + FFMA2", d, sms, 2);
  run<3, 8, 256>("V3 FFMA2 only", d, sms, 2);
  run<4, 8, 256>("V4 FFMA only", d, sms, 2);
  run<5, 8, 256>("V5 FFMA2 + 1 int op each", d, sms, 2);
  run<1, 8, 128>("V1 FFMA.SAT + FFMA2", d, sms, 1);
  run<1, 6, 256>("V1 FFMA.SAT + FFMA2", d, sms, 2);
  run<0, 6, 256>("V0 FFMA.SAT + FFMA (product)", d, sms, 2);
  run<1, 4, 256>("V1 FFMA.SAT + FFMA2", d, sms, 2);
  run<1, 2, 256>("V1 FFMA.SAT + FFMA2", d, sms, 2);
  run<0, 2, 256>("V0 FFMA.SAT + FFMA (product)", d, sms, 2);
  return 0;
}
