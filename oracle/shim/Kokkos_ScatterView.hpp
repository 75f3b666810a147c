// TEST INFRASTRUCTURE — NOT PRODUCT CODE.  See Kokkos_Core.hpp in this directory.
//
// Kokkos::Experimental::ScatterView stand-in.  Kokkos' OpenMP default is
// "duplicated, non-atomic": one private copy of the target per thread, summed
// into the target by contribute().  Two build variants:
//
//   default                    duplicates hold T (float): the reference's own
//                              accumulation arithmetic, thread copies summed
//                              0..T-1.  Run with OMP_NUM_THREADS=1 for a
//                              deterministic, Kokkos-Serial-equivalent order.
//   -DRAGNAR_ORACLE_SCATTER64  duplicates hold double and are rounded to T once
//                              in contribute(): every per-pair term is still the
//                              reference's float value, only the summation is
//                              wide ("ragnar_ref64").
//
// Call sites in the reference: src/physics/synchrotron.hpp:47,88-93,114,169-170,
// src/physics/synchrotron.cpp:79-80,100,129-130,140,
// src/containers/particles.cpp:218-219,247-255, src/physics/ic.cpp:36,41.
#ifndef RAGNAR_ORACLE_KOKKOS_SCATTERVIEW_SHIM_HPP
#define RAGNAR_ORACLE_KOKKOS_SCATTERVIEW_SHIM_HPP

#include <Kokkos_Core.hpp>

#include <memory>
#include <vector>

namespace Kokkos::Experimental {

#if defined(RAGNAR_ORACLE_SCATTER64)
  template <class T>
  using scatter_acc_t = double;
#else
  template <class T>
  using scatter_acc_t = T;
#endif

  template <class DT>
  class ScatterView;

  template <class T>
  class ScatterView<T*> {
    using acc_t = scatter_acc_t<T>;

    struct Storage {
      std::size_t        n { 0 };
      int                nthreads { 1 };
      std::vector<acc_t> dup;
    };

    std::shared_ptr<Storage> m_s;

  public:
    // Kokkos' ScatterValue::operator+= takes the right-hand side as the view's
    // value_type, i.e. a double term such as `1.0 / energy`
    // (src/containers/particles.cpp:249) is rounded to T *before* it is added.
    struct Value {
      acc_t& slot;

      void operator+=(T rhs) const { slot += rhs; }
    };

    struct Accessor {
      acc_t* row;

      Value operator()(std::size_t i) const { return Value { row[i] }; }
    };

    ScatterView() = default;

    explicit ScatterView(const View<T*>& target) : m_s { std::make_shared<Storage>() } {
      m_s->n = target.extent(0);
#if defined(_OPENMP)
      m_s->nthreads = omp_get_max_threads();
#endif
      m_s->dup.assign((std::size_t)m_s->nthreads * m_s->n, acc_t(0));
    }

    Accessor access() const {
      int tid = 0;
#if defined(_OPENMP)
      tid = omp_get_thread_num();
#endif
      return Accessor { m_s->dup.data() + (std::size_t)tid * m_s->n };
    }

    void contribute_into(const View<T*>& target) const {
      for (std::size_t i = 0; i < m_s->n; ++i) {
        acc_t sum = acc_t(target(i));
        for (int t = 0; t < m_s->nthreads; ++t) {
          sum += m_s->dup[(std::size_t)t * m_s->n + i];
        }
        target(i) = static_cast<T>(sum);
      }
    }
  };

  template <class T>
  auto create_scatter_view(const View<T*>& target) -> ScatterView<T*> {
    return ScatterView<T*> { target };
  }

  template <class T>
  void contribute(const View<T*>& target, const ScatterView<T*>& scatter) {
    scatter.contribute_into(target);
  }

} // namespace Kokkos::Experimental

#endif // RAGNAR_ORACLE_KOKKOS_SCATTERVIEW_SHIM_HPP
