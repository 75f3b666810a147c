// Exhaustive check of ragnar_b200/csrc/rgc_glibc_log10f.cuh against the host libm:
//   g++ -O2 -fopenmp -ffp-contract=off -fno-builtin tools/check_log10f.cpp -o /tmp/check_log10f && /tmp/check_log10f [stride]
// Walks every float bit pattern (stride 1, ~10 s on 8 cores) or every stride-th one and
// counts results that are not bit-identical to log10f / logf (NaNs compare as equal).
// Exit code 0 when there are none.  tests/test_log10f_cpu.py runs it with a stride.
#include "../ragnar_b200/csrc/rgc_glibc_log10f.cuh"

#include <cmath>
#include <cstdio>
#include <cstdlib>

static const rgc::LogfEntry kTab[16] = RGC_LOGF_TAB_INIT;

int main(int argc, char** argv) {
  const unsigned long stride = argc > 1 ? std::strtoul(argv[1], nullptr, 10) : 1ul;
  long bad10 = 0, badln = 0;
#pragma omp parallel for reduction(+ : bad10, badln) schedule(static)
  for (long long u = 0; u <= 0xffffffffll; u += (long long)stride) {
    const float    x = rgc::u2f((std::uint32_t)u);
    volatile float xv = x;
    const float    want = log10f(xv), got = rgc::glibc_log10f(x, kTab);
    if (!(std::isnan(want) && std::isnan(got)) && rgc::f2u(want) != rgc::f2u(got)) {
      ++bad10;
    }
    if (u >= 0x00800000ll && u < 0x7f800000ll) { // positive normal: the logf core
      const float w2 = logf(xv), g2 = rgc::glibc_logf_normal(x, kTab);
      if (rgc::f2u(w2) != rgc::f2u(g2)) {
        ++badln;
      }
    }
  }
  std::printf("stride %lu: log10f mismatches %ld, logf mismatches %ld\n", stride, bad10, badln);
  return (bad10 || badln) ? 1 : 0;
}
