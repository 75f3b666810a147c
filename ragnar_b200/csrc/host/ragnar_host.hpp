// Host side of the `ragnar` Python module: the reference's container / physics
// interface (same class names, arguments and error behaviour) implemented on
// top of the C-ABI in include/ragnar_cuda.h.  No Kokkos, no CUDA headers, no CPU
// compute path: every array lives in a device buffer owned by libragnar_cuda.
//
// Namespace is `rgb` (not the reference's `rgnr`) so this module and the
// reference-built oracle can be imported in one process.
#ifndef RGB_RAGNAR_HOST_HPP
#define RGB_RAGNAR_HOST_HPP

#include "ragnar_cuda.h"

#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>

#include <cstddef>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

namespace py = pybind11;

namespace rgb {

  using real_t = float; // reference src/utils/global.h:11
  using dim_t  = unsigned short;

  // reference src/utils/global.h:21-27
  struct EnergyUnits {
    inline static const std::string eV   = "eV";
    inline static const std::string MeV  = "MeV";
    inline static const std::string GeV  = "GeV";
    inline static const std::string mec2 = "mec2";
    inline static const std::string mpc2 = "mpc2";
  };

  // C-ABI status -> the exception the reference throws at the same point
  void check(int rc);
  // "2.00·10^10" — reference src/utils/snippets.cpp:64-110 (USE_POW10 branch)
  std::string human_readable(double value);

  template <class T>
  constexpr int dtype_of() {
    if constexpr (std::is_same_v<T, int>) {
      return RGC_I32;
    } else if constexpr (std::is_same_v<T, float>) {
      return RGC_F32;
    } else {
      static_assert(std::is_same_v<T, double>);
      return RGC_F64;
    }
  }

  // A device buffer plus a lazily fetched host mirror.  The Python API never
  // mutates an array in place, so the mirror stays valid once filled.
  struct DeviceStorage {
    rgc_buf_t*                dev { nullptr };
    mutable std::vector<char> mirror;
    mutable bool              mirrored { false };

    DeviceStorage() = default;
    DeviceStorage(const DeviceStorage&)            = delete;
    DeviceStorage& operator=(const DeviceStorage&) = delete;
    ~DeviceStorage();
  };

  // reference src/containers/array.hpp:15-42: shallow, ref-counted 1-D device array
  template <class T>
  class Array1D {
    std::shared_ptr<DeviceStorage> m_store;

  public:
    Array1D() = default;
    explicit Array1D(const py::array_t<T, py::array::c_style | py::array::forcecast>& arr);
    explicit Array1D(const std::vector<T>& host);
    // adopts a buffer handle returned by the C-ABI (takes over its reference)
    static Array1D adopt(rgc_buf_t* buf);

    void        head(std::size_t n = 10, std::size_t start = 0) const;
    std::string repr() const;
    py::array_t<T> as_array() const;
    const T*    host_data() const; // cached host mirror (extent(0) elements)
    std::size_t extent(unsigned short d = 0) const;
    // the C-ABI handle behind this array (nullptr for an empty array)
    rgc_buf_t* handle() const { return m_store ? m_store->dev : nullptr; }
  };

  // reference src/containers/bins.hpp:17-33
  struct Bins : Array1D<real_t> {
    bool        log_spaced { false };
    std::string unit;

    Bins(const Array1D<real_t>& arr, const std::string& unit_ = "")
      : Array1D<real_t> { arr }
      , unit { unit_ } {}

    explicit Bins(const std::string& unit_ = "") : unit { unit_ } {}

    Bins(const py::array_t<real_t, py::array::c_style | py::array::forcecast>& arr,
         const std::string& unit_ = "")
      : Array1D<real_t> { arr }
      , unit { unit_ } {}
  };

  Array1D<real_t> Linspace(real_t start, real_t stop, std::size_t num);
  Array1D<real_t> Logspace(real_t start, real_t stop, std::size_t num);
  Bins Linbins(real_t start, real_t stop, std::size_t num, const std::string& unit = "");
  Bins Logbins(real_t start, real_t stop, std::size_t num, const std::string& unit = "");

  // reference src/containers/tabulation.hpp:55-116
  template <bool LG>
  class TabulatedFunction {
    Array1D<real_t> m_x, m_y;
    real_t          m_yfill;
    std::size_t     m_n;
    real_t          m_xmin, m_xmax;

    void finish(); // min/max + the reference's verify()

  public:
    TabulatedFunction(const Array1D<real_t>& x, const Array1D<real_t>& y, real_t yfill = 0.0);
    TabulatedFunction(const py::array_t<real_t, py::array::c_style | py::array::forcecast>& x,
                      const py::array_t<real_t, py::array::c_style | py::array::forcecast>& y,
                      real_t yfill = 0.0);

    const Array1D<real_t>& xArr() const { return m_x; }
    const Array1D<real_t>& yArr() const { return m_y; }
    std::size_t            nPoints() const { return m_n; }
    real_t                 yFill() const { return m_yfill; }
    real_t                 xMin() const { return m_xmin; }
    real_t                 xMax() const { return m_xmax; }
  };

  // reference src/containers/distributions.hpp:15-59.  Values come from
  // rgc_generator_eval (host-exact powf/logf arithmetic).
  struct PlawGenerator {
    const real_t p, emin, emax;
    PlawGenerator(real_t p, real_t emin = 0.0, real_t emax = 0.0);
    Array1D<real_t> compute(const Bins& energy_bins) const;
  };

  struct BrokenPlawGenerator {
    const real_t e_break, emin, emax, p1, p2;
    BrokenPlawGenerator(real_t e_break, real_t p1, real_t p2, real_t emin = 0.0,
                        real_t emax = 0.0);
    Array1D<real_t> compute(const Bins& energy_bins) const;
  };

  struct DeltaGenerator {
    const real_t energy0, denergy;
    DeltaGenerator(real_t energy0, real_t denergy);
    Array1D<real_t> compute(const Bins& energy_bins) const;
  };

  // reference src/containers/distributions.hpp:61-78
  class TabulatedDistribution {
    Bins            m_e_bins;
    Array1D<real_t> m_f;

  public:
    TabulatedDistribution(const Bins& e_bins, const Array1D<real_t>& f);
    TabulatedDistribution(const Bins& e_bins, const PlawGenerator& g);
    TabulatedDistribution(const Bins& e_bins, const BrokenPlawGenerator& g);
    TabulatedDistribution(const Bins& e_bins, const DeltaGenerator& g);

    const Bins&            EnergyBins() const { return m_e_bins; }
    const Array1D<real_t>& F() const { return m_f; }
    std::size_t            extent() const { return m_f.extent(0); }
    bool                   log_spaced() const { return m_e_bins.log_spaced; }
  };

  // reference src/containers/particles.hpp:18-72; the SoA columns live behind
  // an rgc_particles_t handle
  struct ParticleStorage {
    rgc_particles_t* handle { nullptr };
    ParticleStorage()                                  = default;
    ParticleStorage(const ParticleStorage&)            = delete;
    ParticleStorage& operator=(const ParticleStorage&) = delete;
    ~ParticleStorage();
  };

  using column_dict_t =
    std::map<std::string, py::array_t<real_t, py::array::c_style | py::array::forcecast>>;

  template <dim_t D>
  class Particles {
    std::shared_ptr<ParticleStorage> m_store;
    bool                             m_is_allocated { false };
    bool                             m_coords_ignored { false };
    std::size_t                      m_nactive { 0 };
    std::size_t                      m_nalloc { 0 };
    std::string                      m_label;

    Array1D<real_t> column(int quantity, std::size_t d) const;

  public:
    explicit Particles(const std::string& label);
    // wraps a container filled through the C-ABI (used by the plugins)
    static Particles adopt(const std::string& label, rgc_particles_t* handle,
                           std::size_t nparticles, bool coords_ignored);

    void fromArrays(const column_dict_t& arrays, bool append = false);
    void allocate(std::size_t nalloc);
    void reallocate(std::size_t nalloc);
    void setNactive(std::size_t nactive);
    void setIgnoreCoords(bool ignore) { m_coords_ignored = ignore; }

    bool               is_allocated() const { return m_is_allocated; }
    bool               coords_ignored() const { return m_coords_ignored; }
    std::size_t        nalloc() const { return m_nalloc; }
    std::size_t        nactive() const { return m_nactive; }
    const std::string& label() const { return m_label; }
    rgc_particles_t*   handle() const { return m_store ? m_store->handle : nullptr; }

    void        printHead(std::size_t number = 5, std::size_t start = 0) const;
    std::string repr() const;

    TabulatedDistribution energyDistribution(const Bins& energy_bins, bool fourvel = true) const;

    Array1D<real_t> Xarr(std::size_t d) const;
    Array1D<real_t> Uarr(std::size_t d) const;
    Array1D<real_t> Earr(std::size_t d) const;
    Array1D<real_t> Barr(std::size_t d) const;
  };

  // reference src/physics/synchrotron.{hpp,cpp}
  real_t Ffunc_integrand(real_t x);
  TabulatedFunction<true> TabulateFfunc(std::size_t npoints = 200,
                                        real_t      xmin    = static_cast<real_t>(1e-6),
                                        real_t      xmax    = static_cast<real_t>(100));
  Array1D<real_t> SynchrotronSpectrumFromDist(const TabulatedDistribution& dist_prtls,
                                              const Bins& bins_e_syn, real_t g_syn,
                                              real_t e_syn_at_g_syn);
  template <dim_t D>
  Array1D<real_t> SynchrotronSpectrum(const Particles<D>& prtls, const Bins& bins_e_syn,
                                      real_t B0, real_t g_syn, real_t e_syn_at_g_syn);

  // reference src/plugins/tristan-v2.hpp:16-43
  template <dim_t D>
  class TristanV2 {
    std::string m_path;
    std::size_t m_step { 0 };
    bool        is_path_set { false };
    bool        is_step_set { false };

  public:
    std::string label() const { return "Tristan V2"; }
    void        setPath(const std::string& path);
    void        setStep(std::size_t step);
    std::string getPath() const;
    std::size_t getStep() const;
    Particles<D> readParticles(const std::string& label, unsigned short sp,
                               std::size_t start = 0, std::size_t size = 0,
                               std::size_t stride = 1, bool ignore_coordinates = false) const;
  };

  // pybind11 registration, one function per reference pyDefine*
  void define_units(py::module& m);
  void define_spaces(py::module& m);
  void define_arrays_and_bins(py::module& m);
  void define_tabulated_functions(py::module& m);
  void define_generators(py::module& m);
  void define_particles(py::module& m);
  void define_synchrotron(py::module& m);
  void define_h5(py::module& m);      // H5read1DArray_* / H5write1DArray_*
  void define_ic(py::module& m); // ICSpectrum (reference src/physics/ic.cpp)
  void define_tristan(py::module& m);

} // namespace rgb

#endif // RGB_RAGNAR_HOST_HPP
