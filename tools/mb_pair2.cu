// Micro-benchmark of the pair kernel's inner loop exactly as the product writes it
// (rgc_sync_pair.cu: RGC_PAIR_BODY, 8 particles per iteration, ping-pong LDS.128),
// on synthetic shared-memory data without the TMA ring: the ceiling of the loop alone,
// against the number of resident warps and the operand arrangement.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -o tools/_build/mb_pair2 tools/mb_pair2.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

constexpr int kStageLen = 64;

// V = 0: FFMA.SAT(fc, sgn, fap) + FFMA(w, r, s2)        (the product's form)
// V = 1: FADD.SAT(fc, fap)      + FFMA(w, r, s2)        (no sign operand)
// V = 2: FFMA only: s2 = fma(w, fap, s2) twice          (accumulate half alone)
// V = 3: as 0, whole stage unrolled (no inner loop branch)
template <int V, int NA, int T>
__global__ void __launch_bounds__(T) k(float* out, const float4* __restrict__ prt, int stages) {
  __shared__ float4 ring[4 * kStageLen / 2 * (T / 32)];
  for (int i = threadIdx.x; i < 4 * kStageLen / 2 * (T / 32); i += T) {
    ring[i] = prt[i & 255];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  float fap[NA], sgn[NA], s2[NA];
#pragma unroll
  for (int g = 0; g < NA; ++g) {
    fap[g] = (float)((threadIdx.x * NA + g) & 1023) * (1.0f / 1024.0f);
    sgn[g] = (threadIdx.x + g) & 64 ? 1.0f : -1.0f;
    s2[g]  = 0.f;
  }
#define ONE(FC, W)                                                                      \
  {                                                                                     \
    float r[NA];                                                                        \
    if (V == 0 || V == 3) {                                                             \
      _Pragma("unroll") for (int g = 0; g < NA; ++g) { r[g] = __saturatef(fmaf((FC), sgn[g], fap[g])); } \
    } else if (V == 1) {                                                                \
      _Pragma("unroll") for (int g = 0; g < NA; ++g) { r[g] = __saturatef((FC) + fap[g]); } \
    } else {                                                                            \
      _Pragma("unroll") for (int g = 0; g < NA; ++g) { r[g] = fap[g]; s2[g] = fmaf((FC), r[g], s2[g]); } \
    }                                                                                   \
    _Pragma("unroll") for (int g = 0; g < NA; ++g) { s2[g] = fmaf((W), r[g], s2[g]); }  \
  }
#define BODY(Q) ONE((Q).x, (Q).y) ONE((Q).z, (Q).w)
  for (int s = 0; s < stages; ++s) {
    const float4* buf = ring + (warp * 4 + (s & 3)) * (kStageLen / 2);
    if (V == 3) {
#pragma unroll
      for (int p = 0; p < kStageLen / 2; ++p) {
        const float4 q = buf[p];
        BODY(q)
      }
    } else {
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
  }
  float s = 0.f;
#pragma unroll
  for (int g = 0; g < NA; ++g) s += s2[g];
  if (s == 123.456f) out[0] = s;
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
    // FFMA-pipe lane-instructions: 2 per evaluation
    double lanes = 2.0 * NA * kStageLen * (double)stages * (double)grid * T / (ms * 1e-3);
    if (rep && lanes > best) best = lanes;
  }
  printf("%-34s NA=%d  %2d warps/SM  %6.2f TFLOP/s-equiv  %.3e FMA lanes/s\n", name, NA, ctas_per_sm * T / 32,
         2 * best / 1e12, best);
}

int main() {
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  float* d; CK(cudaMalloc(&d, 1 << 20)); CK(cudaMemset(d, 0, 1 << 20));
  run<0, 7, 256>("V0 FFMA.SAT+FFMA (product)", d, sms, 2);
  run<0, 7, 256>("V0 FFMA.SAT+FFMA (product)", d, sms, 1);
  run<0, 7, 128>("V0 FFMA.SAT+FFMA (product)", d, sms, 1);
  run<0, 7, 256>("V0 FFMA.SAT+FFMA (product)", d, sms, 3);
  run<0, 7, 256>("V0 FFMA.SAT+FFMA (product)", d, sms, 4);
  run<0, 4, 256>("V0 FFMA.SAT+FFMA (product)", d, sms, 2);
  run<0, 8, 256>("V0 FFMA.SAT+FFMA (product)", d, sms, 2);
  run<1, 7, 256>("V1 FADD.SAT+FFMA", d, sms, 2);
  run<1, 7, 128>("V1 FADD.SAT+FFMA", d, sms, 1);
  run<2, 7, 256>("V2 FFMA+FFMA", d, sms, 2);
  run<2, 7, 128>("V2 FFMA+FFMA", d, sms, 1);
  run<3, 7, 256>("V3 product, stage unrolled", d, sms, 2);
  run<3, 7, 128>("V3 product, stage unrolled", d, sms, 1);
  return 0;
}
