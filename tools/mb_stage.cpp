// Host-side micro-benchmark behind the pinned staging of pageable sources (rgc_runtime.cu: CopyPool):
// how fast can N threads move a pageable array into a (here: plain, page-touched) stage buffer?
//   g++ -O2 -pthread -mavx2 -o tools/_build/mb_stage tools/mb_stage.cpp && tools/_build/mb_stage
#include <immintrin.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static void copy_memcpy(char* d, const char* s, size_t n) { std::memcpy(d, s, n); }
static void copy_nt256(char* d, const char* s, size_t n) {
  for (size_t i = 0; i + 128 <= n; i += 128) {
    __m256i a = _mm256_loadu_si256((const __m256i*)(s + i)), b = _mm256_loadu_si256((const __m256i*)(s + i + 32));
    __m256i c = _mm256_loadu_si256((const __m256i*)(s + i + 64)), e = _mm256_loadu_si256((const __m256i*)(s + i + 96));
    _mm256_stream_si256((__m256i*)(d + i), a); _mm256_stream_si256((__m256i*)(d + i + 32), b);
    _mm256_stream_si256((__m256i*)(d + i + 64), c); _mm256_stream_si256((__m256i*)(d + i + 96), e);
  }
  _mm_sfence();
}
static void copy_movsb(char* d, const char* s, size_t n) { asm volatile("rep movsb" : "+D"(d), "+S"(s), "+c"(n) : : "memory"); }

int main() {
  const size_t total = size_t(1) << 31, stage = size_t(32) << 20; // 2 GiB source, 32 MiB stages (4 of them)
  char* src = (char*)aligned_alloc(4096, total);
  char* dst = (char*)aligned_alloc(4096, 4 * stage);
  std::memset(src, 1, total); std::memset(dst, 2, 4 * stage);
  struct V { const char* name; void (*fn)(char*, const char*, size_t); } vs[] = {
    {"memcpy", copy_memcpy}, {"nt-avx2", copy_nt256}, {"rep movsb", copy_movsb}};
  for (int nt : {4, 8, 12, 16, 24, 32}) {
    if (nt > (int)std::thread::hardware_concurrency()) break;
    for (auto& v : vs) {
      double best = 0;
      for (int rep = 0; rep < 3; ++rep) {
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back([&, t] {
          // every chunk of 32 MiB is split over the threads, like the pool does
          const size_t per = (stage / nt) & ~size_t(4095);
          for (size_t c = 0; c < total / stage; ++c)
            v.fn(dst + (c % 4) * stage + t * per, src + c * stage + t * per, per);
        });
        for (auto& x : th) x.join();
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        best = std::max(best, total / s / 1e9);
      }
      std::printf("%2d threads  %-10s %6.1f GB/s\n", nt, v.name, best);
    }
  }
  return 0;
}
