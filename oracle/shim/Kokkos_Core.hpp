// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// Minimal host-only stand-in for the subset of the Kokkos 4.x API that the
// reference (haykh/ragnar @ fceb6b08) uses outside io/ and plugins/.  It exists
// so that the reference's *unmodified* sources under /root/reference/src can be
// compiled into the CPU oracle `oracle/_ref/ragnar_ref*.so` (see
// oracle/build_ref.sh).  Nothing in ragnar_b200/ includes or links this file.
//
// Semantics restated here (the only non-reference arithmetic in the oracle):
//   * View<T*>, View<T*[N]>: zero-initialised, ref-counted, host LayoutRight
//     (row-major [i][c]); operator() is const and returns T&.
//   * subview / create_mirror_view / deep_copy on those two view types.
//   * parallel_for over a count, RangePolicy, MDRangePolicy<Rank<2|3>>:
//     outermost index OpenMP-parallel (static schedule), inner ones sequential.
//   * parallel_reduce with MinMax<T> (serial).
//   * Kokkos::sqrt/log10/log/abs/pow follow the <cmath> promotion rules
//     (float stays float, any double/integer operand promotes to double).
//
// API usage sites in the reference: src/utils/snippets.cpp:21-62,
// src/containers/{array,particles,tabulation,distributions}.cpp,
// src/physics/{synchrotron,ic}.{hpp,cpp}.
#ifndef RAGNAR_ORACLE_KOKKOS_CORE_SHIM_HPP
#define RAGNAR_ORACLE_KOKKOS_CORE_SHIM_HPP

#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <limits>
#include <map>
#include <memory>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#if defined(_OPENMP)
  #include <omp.h>
#endif

#define KOKKOS_INLINE_FUNCTION inline
#define KOKKOS_FUNCTION
#define KOKKOS_LAMBDA [=]

namespace Kokkos {

  // ------------------------------------------------------------------ runtime
  namespace shim_detail {
    inline bool& initialized_flag() {
      static bool flag = false;
      return flag;
    }
  } // namespace shim_detail

  inline void initialize() { shim_detail::initialized_flag() = true; }
  inline void finalize() { shim_detail::initialized_flag() = false; }
  inline bool is_initialized() { return shim_detail::initialized_flag(); }
  inline void fence() {}

  // --------------------------------------------------------------------- math
  using std::abs;
  using std::log;
  using std::log10;
  using std::pow;
  using std::sqrt;

  // -------------------------------------------------------------------- views
  struct ALL_t {};
  inline constexpr ALL_t ALL {};

  namespace shim_detail {
    template <class DT>
    struct view_traits;

    template <class T>
    struct view_traits<T*> {
      using value_type = T;
      static constexpr std::size_t rank  = 1;
      static constexpr std::size_t ncols = 1;
    };

    template <class T, std::size_t N>
    struct view_traits<T* [N]> {
      using value_type = T;
      static constexpr std::size_t rank  = 2;
      static constexpr std::size_t ncols = N;
    };
  } // namespace shim_detail

  template <class DT>
  class View {
    using traits = shim_detail::view_traits<DT>;

  public:
    using value_type = typename traits::value_type;

  private:
    std::shared_ptr<value_type[]> m_store;
    value_type*                   m_base { nullptr };
    std::size_t                   m_ext[2] { 0, traits::ncols };
    std::size_t                   m_str[2] { traits::ncols, 1 };

  public:
    View() = default;

    View(const std::string&, std::size_t n)
      : m_store { new value_type[n * traits::ncols + 1]() }
      , m_base { m_store.get() } {
      m_ext[0] = n;
    }

    // used by subview()
    View(std::shared_ptr<value_type[]> store,
         value_type*                   base,
         std::size_t                   n0,
         std::size_t                   s0,
         std::size_t                   s1 = 1)
      : m_store { std::move(store) }
      , m_base { base } {
      m_ext[0] = n0;
      m_str[0] = s0;
      m_str[1] = s1;
    }

    static constexpr std::size_t rank() { return traits::rank; }

    std::size_t extent(std::size_t d) const { return d < traits::rank ? m_ext[d] : 1; }
    std::size_t stride(std::size_t d) const { return m_str[d]; }
    value_type* data() const { return m_base; }
    std::size_t size() const { return m_ext[0] * (traits::rank == 2 ? m_ext[1] : 1); }
    const std::shared_ptr<value_type[]>& store() const { return m_store; }

    value_type& operator()(std::size_t i) const {
      static_assert(traits::rank == 1);
      return m_base[i * m_str[0]];
    }

    value_type& operator()(std::size_t i, std::size_t c) const {
      static_assert(traits::rank == 2);
      return m_base[i * m_str[0] + c * m_str[1]];
    }
  };

  template <class T, class I0, class I1>
  auto subview(const View<T*>& v, const std::pair<I0, I1>& r) -> View<T*> {
    return View<T*> { v.store(),
                      v.data() + (std::size_t)r.first * v.stride(0),
                      (std::size_t)(r.second - r.first),
                      v.stride(0) };
  }

  template <class T, std::size_t N, class I0, class I1>
  auto subview(const View<T* [N]>& v, const std::pair<I0, I1>& r, ALL_t)
    -> View<T* [N]> {
    return View<T* [N]> { v.store(),
                          v.data() + (std::size_t)r.first * v.stride(0),
                          (std::size_t)(r.second - r.first),
                          v.stride(0),
                          v.stride(1) };
  }

  template <class T, std::size_t N, class I0, class I1, class C>
  auto subview(const View<T* [N]>& v, const std::pair<I0, I1>& r, C comp)
    -> std::enable_if_t<std::is_integral_v<C> || std::is_enum_v<C>, View<T*>> {
    return View<T*> { v.store(),
                      v.data() + (std::size_t)r.first * v.stride(0) +
                        (std::size_t)comp * v.stride(1),
                      (std::size_t)(r.second - r.first),
                      v.stride(0) };
  }

  template <class T, std::size_t N, class C>
  auto subview(const View<T* [N]>& v, ALL_t, C comp)
    -> std::enable_if_t<std::is_integral_v<C> || std::is_enum_v<C>, View<T*>> {
    return View<T*> { v.store(),
                      v.data() + (std::size_t)comp * v.stride(1),
                      v.extent(0),
                      v.stride(0) };
  }

  template <class DT>
  auto create_mirror_view(const View<DT>& v) -> View<DT> {
    return v; // host space: the mirror is the view itself
  }

  template <class T>
  void deep_copy(const View<T*>& dst, const View<T*>& src) {
    if (dst.data() == src.data() && dst.stride(0) == src.stride(0)) {
      return;
    }
    const auto n = dst.extent(0) < src.extent(0) ? dst.extent(0) : src.extent(0);
    for (std::size_t i = 0; i < n; ++i) {
      dst(i) = src(i);
    }
  }

  template <class T, std::size_t N>
  void deep_copy(const View<T* [N]>& dst, const View<T* [N]>& src) {
    if (dst.data() == src.data() && dst.stride(0) == src.stride(0)) {
      return;
    }
    const auto n = dst.extent(0) < src.extent(0) ? dst.extent(0) : src.extent(0);
    for (std::size_t i = 0; i < n; ++i) {
      for (std::size_t c = 0; c < N; ++c) {
        dst(i, c) = src(i, c);
      }
    }
  }

  // ----------------------------------------------------------------- policies
  template <class... Props>
  struct RangePolicy {
    std::size_t b { 0 }, e { 0 };
    RangePolicy() = default;
    RangePolicy(std::size_t b_, std::size_t e_) : b { b_ }, e { e_ } {}
  };

  template <unsigned R>
  struct Rank {
    static constexpr unsigned value = R;
  };

  template <class RankT>
  struct MDRangePolicy {
    static constexpr unsigned R = RankT::value;
    std::size_t               lo[R], hi[R];

    template <class LT, class UT>
    MDRangePolicy(const LT (&lower)[R], const UT (&upper)[R]) {
      for (unsigned d = 0; d < R; ++d) {
        lo[d] = (std::size_t)lower[d];
        hi[d] = (std::size_t)upper[d];
      }
    }
  };

  // ------------------------------------------------------------- parallel_for
  template <class F>
  void parallel_for(const std::string&, std::size_t n, const F& f) {
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n; ++i) {
      f((std::size_t)i);
    }
  }

  template <class F, class... P>
  void parallel_for(const std::string&, const RangePolicy<P...>& pol, const F& f) {
#pragma omp parallel for schedule(static)
    for (long long i = (long long)pol.b; i < (long long)pol.e; ++i) {
      f((std::size_t)i);
    }
  }

  template <class F>
  void parallel_for(const std::string&, const MDRangePolicy<Rank<2>>& pol, const F& f) {
#pragma omp parallel for schedule(static)
    for (long long i = (long long)pol.lo[0]; i < (long long)pol.hi[0]; ++i) {
      for (std::size_t j = pol.lo[1]; j < pol.hi[1]; ++j) {
        f((std::size_t)i, j);
      }
    }
  }

  template <class F>
  void parallel_for(const std::string&, const MDRangePolicy<Rank<3>>& pol, const F& f) {
#pragma omp parallel for schedule(static)
    for (long long i = (long long)pol.lo[0]; i < (long long)pol.hi[0]; ++i) {
      for (std::size_t j = pol.lo[1]; j < pol.hi[1]; ++j) {
        for (std::size_t k = pol.lo[2]; k < pol.hi[2]; ++k) {
          f((std::size_t)i, j, k);
        }
      }
    }
  }

  // ---------------------------------------------------------- parallel_reduce
  template <class T>
  struct MinMaxScalar {
    T min_val { std::numeric_limits<T>::max() };
    T max_val { std::numeric_limits<T>::lowest() };
  };

  template <class T>
  struct MinMax {
    MinMaxScalar<T>& ref;
    MinMax(MinMaxScalar<T>& r) : ref { r } {}
  };

  template <class F, class T>
  void parallel_reduce(const std::string&, std::size_t n, const F& f, MinMax<T> red) {
    red.ref = MinMaxScalar<T> {};
    for (std::size_t i = 0; i < n; ++i) {
      f(i, red.ref);
    }
  }

} // namespace Kokkos

#endif // RAGNAR_ORACLE_KOKKOS_CORE_SHIM_HPP
