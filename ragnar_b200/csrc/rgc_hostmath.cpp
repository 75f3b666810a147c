// Host-side pieces of the hot path that must be bit-identical to what the
// reference computes on the host: bin edges, the F(x) table, generator values
// and the histogram index function (from which the device threshold table is
// derived).  Compiled by g++ with -ffp-contract=off and no -march flags, like
// the reference's default build, so float/double promotions and libm calls are
// the same ones the reference makes.
//
// real_t = float (reference src/utils/global.h:11).
#include "rgc_internal.hpp"

#include <cmath>
#include <cstring>
#include <limits>
#include <map>
#include <mutex>
#include <tuple>

namespace rgc {

  using real_t = float;

  // reference src/utils/snippets.cpp:21-37 — all-float arithmetic, the index and
  // (num - 1) are converted from size_t to float
  void host_linspace(float start, float stop, std::size_t num, float* out) {
    for (std::size_t i = 0; i < num; ++i) {
      out[i] = (num == 1) ? start : start + i * (stop - start) / (num - 1);
    }
  }

  // reference src/utils/snippets.cpp:39-62 — exponent in float (log10f), the
  // power itself through pow(int, float) -> double, rounded to float on store
  void host_logspace(float start, float stop, std::size_t num, float* out) {
    const real_t lg_start = std::log10(start);
    const real_t lg_ratio = std::log10(stop / start);
    const real_t denom    = static_cast<real_t>(num - 1);
    for (std::size_t i = 0; i < num; ++i) {
      if (num == 1) {
        out[i] = start;
      } else {
        const real_t expo = lg_start + static_cast<real_t>(i) * lg_ratio / denom;
        out[i]            = static_cast<real_t>(std::pow(10.0, static_cast<double>(expo)));
      }
    }
  }

  // reference src/physics/synchrotron.cpp:28-46: F(x) = x * int_x^inf K_{5/3},
  // asymptote below 1e-5, zero above 20, else a 100-point log-spaced trapezoid
  // on [x, 20] with float abscissae / ordinates and a float running sum fed by
  // double products.
  float host_ffunc_integrand(float x) {
    if (x < 1e-5) {
      const double pref = 4.0 * M_PI / (std::sqrt(3.0) * std::tgamma(0.3333333333333333));
      return static_cast<real_t>(pref * std::pow(0.5 * x, 1.0 / 3.0));
    }
    if (x > 20) {
      return 0.0f;
    }
    if (!(x < 20)) {
      return std::numeric_limits<real_t>::quiet_NaN(); // reference: Logspace(20, 20) throws
    }
    constexpr unsigned kPts = 100;
    real_t             grid[kPts], bessel[kPts];
    host_logspace(x, 20, kPts, grid);
    for (unsigned i = 0; i < kPts; ++i) {
      bessel[i] = static_cast<real_t>(
        std::cyl_bessel_k(1.6666666666666667, static_cast<double>(grid[i])));
    }
    real_t area = 0.0f;
    for (unsigned i = 0; i + 1 < kPts; ++i) {
      const double slab = 0.5 * (bessel[i] + bessel[i + 1]) * (grid[i + 1] - grid[i]);
      area              = static_cast<real_t>(area + slab);
    }
    return x * area;
  }

  // reference src/physics/synchrotron.cpp:48-65.  The reference's guard also
  // tests `ys[i] > 100.0` on a still-zero ys[i]; only `xs[i] < 1e-6` can fire,
  // which it does for i = 0 because float(1e-6) < 1e-6.  The reference rebuilds
  // this table on every spectrum call; here it is built once per (n, xmin, xmax).
  void host_tabulate_ffunc(std::size_t n, float xmin, float xmax, float* xs, float* ys) {
    using key_t = std::tuple<std::size_t, std::uint32_t, std::uint32_t>;
    static std::mutex                                        mtx;
    static std::map<key_t, std::vector<float>>               cache;
    std::uint32_t                                            bmin, bmax;
    std::memcpy(&bmin, &xmin, 4);
    std::memcpy(&bmax, &xmax, 4);
    const key_t                 key { n, bmin, bmax };
    std::lock_guard<std::mutex> lock(mtx);
    auto                        it = cache.find(key);
    if (it == cache.end()) {
      std::vector<float> tab(2 * n);
      host_logspace(xmin, xmax, n, tab.data());
      for (std::size_t i = 0; i < n; ++i) {
        const float xi = tab[i];
        tab[n + i]     = (xi < 1e-6) ? 0.0f : host_ffunc_integrand(xi);
      }
      it = cache.emplace(key, std::move(tab)).first;
    }
    std::memcpy(xs, it->second.data(), n * sizeof(float));
    std::memcpy(ys, it->second.data() + n, n * sizeof(float));
  }

  // reference src/containers/tabulation.hpp:19-53 (scalar host evaluation; the
  // device kernels use the fixed-point form derived in rgc_synchrotron.cu)
  float host_interpolate(bool loggrid, float x0, const float* x, const float* y,
                         std::size_t n, float yfill) {
    real_t xmin = std::numeric_limits<real_t>::max();
    real_t xmax = std::numeric_limits<real_t>::lowest();
    for (std::size_t i = 0; i < n; ++i) {
      xmin = x[i] < xmin ? x[i] : xmin;
      xmax = x[i] > xmax ? x[i] : xmax;
    }
    if (x0 < xmin or x0 >= xmax) {
      return yfill;
    }
    const real_t nm1 = static_cast<real_t>(n - 1);
    if (loggrid) {
      const auto cell = static_cast<std::size_t>(nm1 * std::abs(std::log10(x0 / xmin)) /
                                                 std::log10(xmax / xmin));
      if (cell >= n - 1) {
        return y[n - 1];
      }
      return (y[cell + 1] * std::log10(x0 / x[cell]) + y[cell] * std::log10(x[cell + 1] / x0)) /
             std::log10(x[cell + 1] / x[cell]);
    }
    const auto cell = static_cast<std::size_t>(nm1 * std::abs(x0 - xmin) / (xmax - xmin));
    if (cell >= n - 1) {
      return y[n - 1];
    }
    return (y[cell + 1] * (x0 - x[cell]) + y[cell] * (x[cell + 1] - x0)) /
           (x[cell + 1] - x[cell]);
  }

  // reference src/containers/distributions.cpp:37-130 (powf / logf throughout)
  int host_generator_eval(int kind, const float* prm, const float* energy, std::size_t n,
                          float* out) {
    if (kind == 0) { // PlawGenerator(p, emin, emax)
      const real_t p = prm[0], emin = prm[1], emax = prm[2];
      real_t       norm;
      if (emax == 0.0) {
        norm = -std::pow(emin, p + 1) / (p + 1);
      } else if (p != -1.0) {
        norm = (std::pow(emax, p + 1) - std::pow(emin, p + 1)) / (p + 1);
      } else {
        norm = std::log(emax / emin);
      }
      for (std::size_t i = 0; i < n; ++i) {
        const real_t e = energy[i];
        out[i] = (e < emin or (emax > 0.0 and e >= emax)) ? 0.0f : std::pow(e, p) / norm;
      }
      return RGC_OK;
    }
    if (kind == 1) { // BrokenPlawGenerator(e_break, p1, p2, emin, emax)
      const real_t eb = prm[0], p1 = prm[1], p2 = prm[2], emin = prm[3], emax = prm[4];
      real_t       below, above;
      if (eb <= emin) {
        below = 0.0;
      } else {
        below = (1 - std::pow(emin / eb, p1 + 1)) / (p1 + 1);
      }
      if (emax == 0.0) {
        above = 1 / (-p2 - 1);
      } else if (p2 == -1.0) {
        above = std::log(emax / eb);
      } else {
        above = (std::pow(emax / eb, p2 + 1) - 1) / (p2 + 1);
      }
      const real_t norm = (below + above) * eb;
      for (std::size_t i = 0; i < n; ++i) {
        const real_t e = energy[i];
        if (e < emin or (emax > 0.0 and e >= emax)) {
          out[i] = 0.0f;
        } else {
          out[i] = std::pow(e / eb, e < eb ? p1 : p2) / norm;
        }
      }
      return RGC_OK;
    }
    if (kind == 2) { // DeltaGenerator(energy0, denergy)
      const real_t e0 = prm[0], de = prm[1];
      const real_t norm = static_cast<real_t>(1.0 / de);
      for (std::size_t i = 0; i < n; ++i) {
        out[i] = (std::abs(energy[i] - e0) < de * 0.5) ? static_cast<real_t>(1.0 * norm) : 0.0f;
      }
      return RGC_OK;
    }
    return fail(RGC_ERR_INVALID, "unknown generator kind %d", kind);
  }

  // reference src/containers/particles.cpp:231 — the ternary's common type makes
  // `energy` a double: (double)sqrtf(Usqr) or sqrt(1.0 + (double)Usqr)
  double host_energy_from_usqr(float Usqr, bool fourvel) {
    return fourvel ? static_cast<double>(std::sqrt(Usqr)) : std::sqrt(1.0 + Usqr);
  }

  // reference src/containers/particles.cpp:233-245: float(n-1) * |log10(energy /
  // energy_min)| evaluated in double, divided by a float log10f, truncated.
  std::size_t host_energy_bin_index(float Usqr, bool fourvel, float energy_min,
                                    float energy_max, std::size_t n) {
    const double energy = host_energy_from_usqr(Usqr, fourvel);
    if (energy < energy_min) {
      return 0;
    }
    if (energy >= energy_max) {
      return n - 1;
    }
    const double scaled = static_cast<real_t>(n - 1) *
                          std::abs(std::log10(energy / energy_min)) /
                          std::log10(energy_max / energy_min);
    // NaN / out-of-range conversions are undefined in C++; x86 cvttsd2si yields
    // 0x8000000000000000, which the reference's clamp turns into n - 1.
    if (!(scaled >= 0.0) || scaled >= 9.2e18) {
      return n - 1;
    }
    const auto cell = static_cast<std::size_t>(scaled);
    return cell > n - 1 ? n - 1 : cell;
  }

} // namespace rgc

using namespace rgc;

extern "C" {

  int rgc_linspace(float start, float stop, size_t num, float* out) {
    if (start >= stop) {
      return fail(RGC_ERR_INVALID, "Linspace start must be < stop");
    }
    host_linspace(start, stop, num, out);
    return RGC_OK;
  }

  int rgc_logspace(float start, float stop, size_t num, float* out) {
    if (start <= 0.0 or stop <= 0.0) {
      return fail(RGC_ERR_INVALID, "Logspace start and stop must be strictly positive");
    }
    if (start >= stop) {
      return fail(RGC_ERR_INVALID, "Logspace start must be < stop");
    }
    host_logspace(start, stop, num, out);
    return RGC_OK;
  }

  int rgc_sync_ffunc_integrand(float x, float* out) {
    if (x == 20.0f) {
      return fail(RGC_ERR_INVALID, "Logspace start must be < stop");
    }
    *out = host_ffunc_integrand(x);
    return RGC_OK;
  }

  int rgc_sync_tabulate_ffunc(size_t npoints, float xmin, float xmax, float* xs, float* ys) {
    if (npoints < 2) {
      return fail(RGC_ERR_INVALID, "TabulateFfunc needs at least 2 points");
    }
    if (xmin <= 0.0 or xmax <= 0.0) {
      return fail(RGC_ERR_INVALID, "Logspace start and stop must be strictly positive");
    }
    if (xmin >= xmax) {
      return fail(RGC_ERR_INVALID, "Logspace start must be < stop");
    }
    host_tabulate_ffunc(npoints, xmin, xmax, xs, ys);
    return RGC_OK;
  }

  int rgc_interpolate(int loggrid, float x0, const float* x, const float* y, size_t n,
                      float yfill, float* out) {
    if (n < 2) {
      return fail(RGC_ERR_INVALID, "table needs at least 2 points");
    }
    *out = host_interpolate(loggrid != 0, x0, x, y, n, yfill);
    return RGC_OK;
  }

  int rgc_generator_eval(int kind, const float* params, const float* energy, size_t n,
                         float* out) {
    return host_generator_eval(kind, params, energy, n, out);
  }

} // extern "C"
