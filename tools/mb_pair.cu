// Micro-benchmark of candidate inner loops for the bucketed hinge pair kernel.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -o tools/_build/mb_pair tools/mb_pair.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

constexpr int T = 256;

template <int V, int G>
__global__ void __launch_bounds__(T) k(float* out, const float4* __restrict__ prt, int iters, float one) {
  __shared__ float4 sp[256];
  sp[threadIdx.x] = prt[threadIdx.x];
  __syncthreads();
  float fa[G], s2[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    fa[g] = (float)((threadIdx.x * G + g) & 1023) * (1.0f / 1024.0f);
    s2[g] = 0.f;
  }
  for (int it = 0; it < iters; ++it) {
    const float4 p = sp[it & 255];
    if (V == 0) {  // FADD.SAT + FFMA, 2 particles / iteration
#pragma unroll
      for (int g = 0; g < G; ++g) {
        s2[g] = fmaf(p.y, __saturatef(fa[g] + p.x), s2[g]);
        s2[g] = fmaf(p.w, __saturatef(fa[g] + p.z), s2[g]);
      }
    } else if (V == 1) {  // all r first, then all FFMA
      float r0[G], r1[G];
#pragma unroll
      for (int g = 0; g < G; ++g) { r0[g] = __saturatef(fa[g] + p.x); r1[g] = __saturatef(fa[g] + p.z); }
#pragma unroll
      for (int g = 0; g < G; ++g) { s2[g] = fmaf(p.y, r0[g], s2[g]); }
#pragma unroll
      for (int g = 0; g < G; ++g) { s2[g] = fmaf(p.w, r1[g], s2[g]); }
    } else if (V == 2) {  // FADD + FMNMX + FFMA
#pragma unroll
      for (int g = 0; g < G; ++g) {
        s2[g] = fmaf(p.y, fmaxf(fa[g] + p.x, 0.f), s2[g]);
        s2[g] = fmaf(p.w, fmaxf(fa[g] + p.z, 0.f), s2[g]);
      }
    } else if (V == 3) {  // FFMA.SAT + FFMA
#pragma unroll
      for (int g = 0; g < G; ++g) {
        s2[g] = fmaf(p.y, __saturatef(fmaf(fa[g], one, p.x)), s2[g]);
        s2[g] = fmaf(p.w, __saturatef(fmaf(fa[g], one, p.z)), s2[g]);
      }
    } else if (V == 4) {  // abs trick: FADD + FFMA(|.|)
#pragma unroll
      for (int g = 0; g < G; ++g) {
        s2[g] = fmaf(p.y, fabsf(fa[g] + p.x), s2[g]);
        s2[g] = fmaf(p.w, fabsf(fa[g] + p.z), s2[g]);
      }
    } else if (V == 5) {  // plain FADD + FFMA (no sat, no abs): pipe baseline
#pragma unroll
      for (int g = 0; g < G; ++g) {
        s2[g] = fmaf(p.y, fa[g] + p.x, s2[g]);
        s2[g] = fmaf(p.w, fa[g] + p.z, s2[g]);
      }
    } else if (V == 6) {  // packed: FADD2 + 2 FMNMX + FFMA2 over group pairs
      unsigned long long fc0, fc1, w0, w1;
      asm("mov.b64 %0, {%1, %1};" : "=l"(fc0) : "f"(p.x));
      asm("mov.b64 %0, {%1, %1};" : "=l"(fc1) : "f"(p.z));
      asm("mov.b64 %0, {%1, %1};" : "=l"(w0) : "f"(p.y));
      asm("mov.b64 %0, {%1, %1};" : "=l"(w1) : "f"(p.w));
#pragma unroll
      for (int g = 0; g < G; g += 2) {
        unsigned long long a, acc, u;
        float ux, uy;
        asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(fa[g]), "f"(fa[g + 1]));
        asm("mov.b64 %0, {%1, %2};" : "=l"(acc) : "f"(s2[g]), "f"(s2[g + 1]));
        asm("add.rn.f32x2 %0, %1, %2;" : "=l"(u) : "l"(a), "l"(fc0));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(ux), "=f"(uy) : "l"(u));
        ux = fmaxf(ux, 0.f); uy = fmaxf(uy, 0.f);
        asm("mov.b64 %0, {%1, %2};" : "=l"(u) : "f"(ux), "f"(uy));
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(acc) : "l"(w0), "l"(u), "l"(acc));
        asm("add.rn.f32x2 %0, %1, %2;" : "=l"(u) : "l"(a), "l"(fc1));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(ux), "=f"(uy) : "l"(u));
        ux = fmaxf(ux, 0.f); uy = fmaxf(uy, 0.f);
        asm("mov.b64 %0, {%1, %2};" : "=l"(u) : "f"(ux), "f"(uy));
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(acc) : "l"(w1), "l"(u), "l"(acc));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(s2[g]), "=f"(s2[g + 1]) : "l"(acc));
      }
    } else if (V == 7) {  // scalar FADD.SAT, packed FFMA2 accumulate
      unsigned long long w0, w1;
      asm("mov.b64 %0, {%1, %1};" : "=l"(w0) : "f"(p.y));
      asm("mov.b64 %0, {%1, %1};" : "=l"(w1) : "f"(p.w));
#pragma unroll
      for (int g = 0; g < G; g += 2) {
        unsigned long long acc, u;
        asm("mov.b64 %0, {%1, %2};" : "=l"(acc) : "f"(s2[g]), "f"(s2[g + 1]));
        asm("mov.b64 %0, {%1, %2};" : "=l"(u) : "f"(__saturatef(fa[g] + p.x)), "f"(__saturatef(fa[g + 1] + p.x)));
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(acc) : "l"(w0), "l"(u), "l"(acc));
        asm("mov.b64 %0, {%1, %2};" : "=l"(u) : "f"(__saturatef(fa[g] + p.z)), "f"(__saturatef(fa[g + 1] + p.z)));
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(acc) : "l"(w1), "l"(u), "l"(acc));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(s2[g]), "=f"(s2[g + 1]) : "l"(acc));
      }
    } else if (V == 8) {  // FFMA only x2 (pipe ceiling with this operand pattern)
#pragma unroll
      for (int g = 0; g < G; ++g) {
        s2[g] = fmaf(p.y, fa[g], s2[g]);
        s2[g] = fmaf(p.w, fa[g], s2[g]);
      }
    } else if (V == 9) {  // packed FFMA2 only (is FFMA2 2x FFMA per issue?)
      unsigned long long w0, w1;
      asm("mov.b64 %0, {%1, %1};" : "=l"(w0) : "f"(p.y));
      asm("mov.b64 %0, {%1, %1};" : "=l"(w1) : "f"(p.w));
#pragma unroll
      for (int g = 0; g < G; g += 2) {
        unsigned long long acc, a;
        asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(fa[g]), "f"(fa[g + 1]));
        asm("mov.b64 %0, {%1, %2};" : "=l"(acc) : "f"(s2[g]), "f"(s2[g + 1]));
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(acc) : "l"(w0), "l"(a), "l"(acc));
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(acc) : "l"(w1), "l"(a), "l"(acc));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(s2[g]), "=f"(s2[g + 1]) : "l"(acc));
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int g = 0; g < G; ++g) s += s2[g];
  if (s == 123.456f) out[0] = s;
}

template <int V, int G>
void run(const char* name, float* d, int sms) {
  const int iters = 8192, grid = sms * 8;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  double best = 0;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0));
    k<V, G><<<grid, T>>>(d, reinterpret_cast<const float4*>(d) + 64, iters, 1.0f);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double evals = 2.0 * G * iters * (double)grid * T / (ms * 1e-3);
    if (rep && evals > best) best = evals;
  }
  printf("%-40s G=%d  %.2f Tevals/s  %.1f evals/clk/SM @1965MHz\n", name, G, best / 1e12, best / sms / 1.965e9);
}

int main() {
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  float* d; CK(cudaMalloc(&d, 1 << 20)); CK(cudaMemset(d, 0, 1 << 20));
  run<0, 8>("V0 FADD.SAT+FFMA", d, sms);
  run<0, 7>("V0 FADD.SAT+FFMA", d, sms);
  run<0, 4>("V0 FADD.SAT+FFMA", d, sms);
  run<0, 16>("V0 FADD.SAT+FFMA", d, sms);
  run<1, 8>("V1 r first then FFMA", d, sms);
  run<2, 8>("V2 FADD+FMNMX+FFMA", d, sms);
  run<3, 8>("V3 FFMA.SAT+FFMA", d, sms);
  run<4, 8>("V4 FADD+FFMA|abs|", d, sms);
  run<5, 8>("V5 FADD+FFMA plain", d, sms);
  run<6, 8>("V6 FADD2+2FMNMX+FFMA2", d, sms);
  run<7, 8>("V7 2FADD.SAT+FFMA2", d, sms);
  run<8, 8>("V8 FFMA only (2/eval-pair)", d, sms);
  run<9, 8>("V9 FFMA2 only", d, sms);
  return 0;
}
