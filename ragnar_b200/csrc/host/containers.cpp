// Array1D / Bins / spaces / TabulatedFunction / generators / TabulatedDistribution /
// Particles — host-side mirror of the reference's src/containers and src/utils
// on top of the C-ABI.  Error types and messages follow the reference so that
// pybind11 raises the same Python exceptions.
#include "docstrings.hpp"
#include "ragnar_host.hpp"

#include <pybind11/stl.h>

#include <algorithm>
#include <array>
#include <cstdio>
#include <cstring>
#include <limits>

using namespace pybind11::literals;

namespace rgb {

  // ------------------------------------------------------------------ helpers
  void check(int rc) {
    if (rc == RGC_OK) {
      return;
    }
    const std::string msg = rgc_last_error();
    throw std::runtime_error(msg.empty() ? "ragnar_cuda error " + std::to_string(rc) : msg);
  }

  std::string human_readable(double value) {
    // reference snippets.cpp:97-110; zero is guarded (the reference loops forever on it)
    const bool negative = value < 0;
    double     v        = negative ? -value : value;
    int        pow      = 0;
    if (v > 0 && std::isfinite(v)) {
      while (v < 0.1 or v >= 10) {
        if (v < 0.1) {
          v *= 10;
          --pow;
        } else {
          v /= 10;
          ++pow;
        }
      }
    }
    char buf[64];
    std::snprintf(buf, sizeof(buf), "%.2f", v);
    return (negative ? "-" : "") + std::string(buf) + "·10^" + std::to_string(pow);
  }

  DeviceStorage::~DeviceStorage() {
    if (dev) {
      rgc_buf_release(dev);
    }
  }

  ParticleStorage::~ParticleStorage() {
    if (handle) {
      rgc_particles_release(handle);
    }
  }

  // ------------------------------------------------------------------ Array1D
  template <class T>
  Array1D<T>::Array1D(const py::array_t<T, py::array::c_style | py::array::forcecast>& arr)
    : m_store { std::make_shared<DeviceStorage>() } {
    check(rgc_buf_from_host(dtype_of<T>(), arr.data(), (std::size_t)arr.size(), &m_store->dev));
  }

  template <class T>
  Array1D<T>::Array1D(const std::vector<T>& host) : m_store { std::make_shared<DeviceStorage>() } {
    check(rgc_buf_from_host(dtype_of<T>(), host.data(), host.size(), &m_store->dev));
    m_store->mirror.resize(host.size() * sizeof(T));
    std::memcpy(m_store->mirror.data(), host.data(), host.size() * sizeof(T));
    m_store->mirrored = true;
  }

  template <class T>
  Array1D<T> Array1D<T>::adopt(rgc_buf_t* buf) {
    Array1D<T> a;
    a.m_store      = std::make_shared<DeviceStorage>();
    a.m_store->dev = buf;
    return a;
  }

  template <class T>
  std::size_t Array1D<T>::extent(unsigned short d) const {
    if (d != 0) {
      return 1; // a rank-1 Kokkos::View reports 1 beyond its rank
    }
    return m_store ? rgc_buf_size(m_store->dev) : 0;
  }

  template <class T>
  const T* Array1D<T>::host_data() const {
    static const T none {};
    if (!m_store || extent(0) == 0) {
      return &none;
    }
    if (!m_store->mirrored) {
      m_store->mirror.resize(extent(0) * sizeof(T));
      check(rgc_buf_to_host(m_store->dev, 0, extent(0), m_store->mirror.data()));
      m_store->mirrored = true;
    }
    return reinterpret_cast<const T*>(m_store->mirror.data());
  }

  template <class T>
  void Array1D<T>::head(std::size_t n, std::size_t start) const {
    if (start + n > extent(0)) {
      throw std::range_error("Array::head: n > data.extent(0)");
    }
    std::vector<T> part(n);
    if (n > 0) {
      check(rgc_buf_to_host(m_store->dev, start, n, part.data()));
    }
    for (std::size_t i = 0; i < n; ++i) {
      py::print(part[i], " ", "end"_a = "");
    }
    py::print();
  }

  template <class T>
  std::string Array1D<T>::repr() const {
    return "1D Array [ size: " + std::to_string(extent(0)) + " ]";
  }

  template <class T>
  py::array_t<T> Array1D<T>::as_array() const {
    py::array_t<T> out((py::ssize_t)extent(0));
    if (extent(0) > 0) {
      if (m_store->mirrored) {
        std::memcpy(out.mutable_data(), m_store->mirror.data(), extent(0) * sizeof(T));
      } else {
        check(rgc_buf_to_host(m_store->dev, 0, extent(0), out.mutable_data()));
      }
    }
    return out;
  }

  template class Array1D<int>;
  template class Array1D<float>;
  template class Array1D<double>;

  // ------------------------------------------------------------------- spaces
  // Grids beyond kDeviceGridMin points are built on the device (bit-identical, see
  // rgc_spaces.cu) and never staged through a host vector; smaller ones — bin edges, F
  // tables — keep their host mirror, which the kernels' host-side planning reads.
  constexpr std::size_t kDeviceGridMin = std::size_t(1) << 16;

  Array1D<real_t> Linspace(real_t start, real_t stop, std::size_t num) {
    if (num > kDeviceGridMin) {
      rgc_buf_t* buf = nullptr;
      check(rgc_linspace_device(start, stop, num, &buf));
      return Array1D<real_t>::adopt(buf);
    }
    std::vector<real_t> edges(num);
    check(rgc_linspace(start, stop, num, edges.data()));
    return Array1D<real_t> { edges };
  }

  Array1D<real_t> Logspace(real_t start, real_t stop, std::size_t num) {
    if (num > kDeviceGridMin) {
      rgc_buf_t* buf = nullptr;
      check(rgc_logspace_device(start, stop, num, &buf));
      return Array1D<real_t>::adopt(buf);
    }
    std::vector<real_t> edges(num);
    check(rgc_logspace(start, stop, num, edges.data()));
    return Array1D<real_t> { edges };
  }

  Bins Linbins(real_t start, real_t stop, std::size_t num, const std::string& unit) {
    Bins b { Linspace(start, stop, num), unit };
    b.log_spaced = false;
    return b;
  }

  Bins Logbins(real_t start, real_t stop, std::size_t num, const std::string& unit) {
    Bins b { Logspace(start, stop, num), unit };
    b.log_spaced = true;
    return b;
  }

  // -------------------------------------------------------- TabulatedFunction
  template <bool LG>
  void TabulatedFunction<LG>::finish() {
    // reference tabulation.cpp:84-117
    m_xmin = std::numeric_limits<real_t>::max();
    m_xmax = std::numeric_limits<real_t>::lowest();
    if (m_n > kDeviceGridMin) { // large tables: the MinMax reduction runs where the data is
      check(rgc_buf_minmax(m_x.handle(), &m_xmin, &m_xmax));
    } else {
      const real_t* x = m_x.host_data();
      for (std::size_t i = 0; i < m_n; ++i) {
        m_xmin = x[i] < m_xmin ? x[i] : m_xmin;
        m_xmax = x[i] > m_xmax ? x[i] : m_xmax;
      }
    }
    if (m_y.extent(0) != m_n) {
      throw std::range_error("y.size != x.size in TabulatedFunction");
    }
    if (m_xmin >= m_xmax) {
      throw std::range_error("xmin >= xmax in TabulatedFunction");
    }
    if (LG and m_xmin <= 0.0) {
      throw std::range_error("xmin <= 0.0 in Logspace TabulatedFunction");
    }
  }

  template <bool LG>
  TabulatedFunction<LG>::TabulatedFunction(const Array1D<real_t>& x, const Array1D<real_t>& y,
                                           real_t yfill)
    : m_x { x }
    , m_y { y }
    , m_yfill { yfill }
    , m_n { x.extent(0) } {
    finish();
  }

  template <bool LG>
  TabulatedFunction<LG>::TabulatedFunction(
    const py::array_t<real_t, py::array::c_style | py::array::forcecast>& x,
    const py::array_t<real_t, py::array::c_style | py::array::forcecast>& y, real_t yfill)
    : m_x { x }
    , m_y { y }
    , m_yfill { yfill }
    , m_n { (std::size_t)x.size() } {
    finish();
  }

  template class TabulatedFunction<true>;
  template class TabulatedFunction<false>;

  // --------------------------------------------------------------- generators
  static Array1D<real_t> eval_generator(int kind, const std::vector<real_t>& params,
                                        const Bins& bins) {
    const std::size_t   n = bins.extent(0);
    std::vector<real_t> f(n);
    check(rgc_generator_eval(kind, params.data(), bins.host_data(), n, f.data()));
    return Array1D<real_t> { f };
  }

  // validation: reference distributions.cpp:42-51
  PlawGenerator::PlawGenerator(real_t p_, real_t emin_, real_t emax_)
    : p { p_ }
    , emin { emin_ }
    , emax { emax_ } {
    if (emin < 0.0) {
      throw std::runtime_error("emin < 0.0");
    } else if (emin > emax and emax > 0.0) {
      throw std::runtime_error("emin > emax");
    } else if (emin == 0.0 and p <= -1.0) {
      throw std::runtime_error("p <= -1 and emin = 0.0 : normalization diverges");
    } else if (emax == 0.0 and p >= -1.0) {
      throw std::runtime_error("p >= -1 and emax = 0.0 (infinity) : normalization diverges");
    }
  }

  Array1D<real_t> PlawGenerator::compute(const Bins& bins) const {
    return eval_generator(0, { p, emin, emax }, bins);
  }

  // validation: reference distributions.cpp:80-100
  BrokenPlawGenerator::BrokenPlawGenerator(real_t e_break_, real_t p1_, real_t p2_, real_t emin_,
                                           real_t emax_)
    : e_break { e_break_ }
    , emin { emin_ }
    , emax { emax_ }
    , p1 { p1_ }
    , p2 { p2_ } {
    if (e_break <= 0.0) {
      throw std::runtime_error("e_break <= 0.0");
    }
    if (emin < 0.0) {
      throw std::runtime_error("emin < 0.0");
    }
    if (emin > emax and emax > 0.0) {
      throw std::runtime_error("emin > emax");
    }
    if (emin == 0.0 and p1 <= -1.0) {
      throw std::runtime_error("p1 <= -1 and emin = 0.0 : normalization diverges");
    }
    if (emax == 0.0 and p2 >= -1.0) {
      throw std::runtime_error("p2 >= -1 and emax = 0.0 (infinity) : normalization diverges");
    }
  }

  Array1D<real_t> BrokenPlawGenerator::compute(const Bins& bins) const {
    return eval_generator(1, { e_break, p1, p2, emin, emax }, bins);
  }

  DeltaGenerator::DeltaGenerator(real_t energy0_, real_t denergy_)
    : energy0 { energy0_ }
    , denergy { denergy_ } {}

  Array1D<real_t> DeltaGenerator::compute(const Bins& bins) const {
    return eval_generator(2, { energy0, denergy }, bins);
  }

  TabulatedDistribution::TabulatedDistribution(const Bins& e_bins, const Array1D<real_t>& f)
    : m_e_bins { e_bins }
    , m_f { f } {
    if (m_e_bins.extent() != m_f.extent()) {
      throw std::runtime_error("e_bins.extent() != f.extent()");
    }
  }

  TabulatedDistribution::TabulatedDistribution(const Bins& e_bins, const PlawGenerator& g)
    : m_e_bins { e_bins }
    , m_f { g.compute(e_bins) } {}

  TabulatedDistribution::TabulatedDistribution(const Bins& e_bins, const BrokenPlawGenerator& g)
    : m_e_bins { e_bins }
    , m_f { g.compute(e_bins) } {}

  TabulatedDistribution::TabulatedDistribution(const Bins& e_bins, const DeltaGenerator& g)
    : m_e_bins { e_bins }
    , m_f { g.compute(e_bins) } {}

  // ---------------------------------------------------------------- Particles
  template <dim_t D>
  Particles<D>::Particles(const std::string& label) : m_label { label } {}

  template <dim_t D>
  Particles<D> Particles<D>::adopt(const std::string& label, rgc_particles_t* handle,
                                   std::size_t nparticles, bool coords_ignored) {
    Particles<D> p { label };
    p.m_store            = std::make_shared<ParticleStorage>();
    p.m_store->handle    = handle;
    p.m_is_allocated     = true;
    p.m_coords_ignored   = coords_ignored;
    p.m_nalloc           = rgc_particles_nalloc(handle);
    p.m_nactive          = nparticles;
    return p;
  }

  template <dim_t D>
  void Particles<D>::allocate(std::size_t nalloc) {
    if (is_allocated()) {
      throw std::runtime_error("Particles already allocated");
    }
    auto store = std::make_shared<ParticleStorage>();
    check(rgc_particles_create(D, &store->handle));
    check(rgc_particles_allocate(store->handle, nalloc, m_coords_ignored ? 0 : 1));
    m_store        = std::move(store);
    m_is_allocated = true;
    m_nalloc       = nalloc;
  }

  template <dim_t D>
  void Particles<D>::reallocate(std::size_t nalloc) {
    if (!is_allocated()) {
      throw std::runtime_error(
        "Particles not allocated, if you want to allocate, call `allocate` instead");
    }
    if (nalloc <= m_nalloc) {
      throw std::runtime_error("New allocation size must be greater than the current one");
    }
    check(rgc_particles_reallocate(m_store->handle, nalloc));
    m_nalloc = nalloc;
  }

  template <dim_t D>
  void Particles<D>::setNactive(std::size_t nactive) {
    if (!is_allocated()) {
      throw std::runtime_error("Particles not allocated");
    }
    if (nactive > m_nalloc) {
      // the reference does not check; its kernels would then read out of bounds
      throw std::runtime_error("nactive exceeds the allocated number of particles");
    }
    m_nactive = nactive;
  }

  // reference particles.cpp:25-100
  template <dim_t D>
  void Particles<D>::fromArrays(const column_dict_t& arrays, bool append) {
    if ((not append) and is_allocated()) {
      throw std::runtime_error("Particles already allocated, if you want to "
                               "append, specify `append = True`");
    }
    std::size_t nprtls = 0;
    for (const auto& [name, arr] : arrays) {
      if (arr.ndim() < 1) {
        throw std::runtime_error("Inconsistent number of particles");
      }
      if (nprtls == 0) {
        nprtls = (std::size_t)arr.shape(0);
      } else if (nprtls != (std::size_t)arr.shape(0)) {
        throw std::runtime_error("Inconsistent number of particles");
      }
    }
    if (nprtls == 0) {
      throw std::runtime_error("No particles provided");
    }
    std::size_t start = 0;
    if (is_allocated()) {
      start = m_nactive;
      if (start + nprtls > m_nalloc) {
        reallocate(m_nactive + nprtls);
      }
    } else {
      allocate(nprtls);
    }
    bool has_coords = not m_coords_ignored;
    for (auto d = 0u; d < D; ++d) {
      if (arrays.find("X" + std::to_string(d + 1)) != arrays.end()) {
        has_coords = true;
      }
    }
    setIgnoreCoords(not has_coords);
    if (has_coords and not rgc_particles_has_coords(m_store->handle)) {
      check(rgc_particles_enable_coords(m_store->handle));
    }
    static const std::array<std::pair<const char*, int>, 4> quantities {
      { { "X", RGC_Q_X }, { "U", RGC_Q_U }, { "E", RGC_Q_E }, { "B", RGC_Q_B } }
    };
    for (const auto& [qname, qid] : quantities) {
      const unsigned ncomp = qid == RGC_Q_X ? D : 3u;
      for (auto d = 0u; d < ncomp; ++d) {
        const auto it = arrays.find(std::string(qname) + std::to_string(d + 1));
        if (it != arrays.end()) {
          check(rgc_particles_write(m_store->handle, qid, (int)d, start, it->second.data(),
                                    nprtls));
        }
      }
    }
    check(rgc_synchronize()); // the numpy sources may go away after we return
    setNactive(start + nprtls);
  }

  template <dim_t D>
  void Particles<D>::printHead(std::size_t number, std::size_t start) const {
    py::print("Particles:", m_label);
    if (!is_allocated()) {
      py::print(" [ not allocated ]");
      return;
    }
    if (start + number > m_nalloc) {
      throw std::runtime_error("Number of particles to print exceeds allocated space");
    }
    py::print(" [", m_nactive, "/", m_nalloc, "]");
    py::print(" showing", start, "to", start + number);
    std::vector<real_t> vals(number);
    auto print_row = [&](int qid, unsigned d, const std::string& name, bool nactive_rule) {
      check(rgc_particles_read(m_store->handle, qid, (int)d, start, number, vals.data()));
      if (qid == RGC_Q_X) {
        py::print(" " + name, ":", "end"_a = "");
      } else {
        py::print(" ", name, ":", "end"_a = "");
      }
      if (start > 0) {
        py::print("...", "end"_a = "");
      }
      for (std::size_t i = 0; i < number; ++i) {
        if (qid == RGC_Q_X) {
          py::print(vals[i], "end"_a = "");
        } else {
          py::print(vals[i], " ", "end"_a = "");
        }
      }
      const bool more = nactive_rule ? m_nactive > number : m_nalloc > start + number;
      if (more) {
        py::print("...");
      } else {
        py::print();
      }
    };
    if (not m_coords_ignored and rgc_particles_has_coords(m_store->handle)) {
      for (auto d = 0u; d < D; ++d) {
        print_row(RGC_Q_X, d, "X_" + std::to_string(d + 1), true);
      }
    }
    const char* names[3] = { "U", "E", "B" };
    const int   qids[3]  = { RGC_Q_U, RGC_Q_E, RGC_Q_B };
    for (int q = 0; q < 3; ++q) {
      for (auto d = 0u; d < 3u; ++d) {
        print_row(qids[q], d, std::string(names[q]) + "_" + std::to_string(d + 1), false);
      }
    }
  }

  template <dim_t D>
  std::string Particles<D>::repr() const {
    return "Particles<" + std::to_string(D) + "D> (" + m_label + ") : " +
           (is_allocated() ? human_readable((double)nactive()) : "not allocated");
  }

  // reference particles.cpp:189-260
  template <dim_t D>
  TabulatedDistribution Particles<D>::energyDistribution(const Bins& energy_bins,
                                                         bool        fourvel) const {
    py::print("Computing energy distribution for", label(), "...", "end"_a = "", "flush"_a = true);
    if (!is_allocated()) {
      throw std::runtime_error("Particles not allocated");
    }
    const std::size_t   n = energy_bins.extent(0);
    std::vector<real_t> hist(n, 0.0f);
    check(rgc_energy_histogram(m_store->handle, m_nactive, energy_bins.host_data(), n,
                               energy_bins.log_spaced ? 1 : 0, fourvel ? 1 : 0, hist.data(),
                               nullptr, nullptr));
    py::print(": OK", "flush"_a = true);
    return TabulatedDistribution { energy_bins, Array1D<real_t> { hist } };
  }

  // reference particles.cpp:346-383 (getSubview): copy of the active range
  template <dim_t D>
  Array1D<real_t> Particles<D>::column(int quantity, std::size_t d) const {
    const std::size_t ncomp = quantity == RGC_Q_X ? D : 3;
    if (d - 1 >= ncomp) { // d is 1-based; d = 0 wraps like the reference's size_t
      throw std::out_of_range("Invalid component");
    }
    if (!is_allocated()) {
      return Array1D<real_t> {};
    }
    rgc_buf_t* buf = nullptr;
    check(rgc_particles_column(m_store->handle, quantity, (int)(d - 1), m_nactive, &buf));
    return Array1D<real_t>::adopt(buf);
  }

  template <dim_t D>
  Array1D<real_t> Particles<D>::Xarr(std::size_t d) const {
    if (m_coords_ignored) {
      throw std::runtime_error("Particle coordinates ignored");
    }
    return column(RGC_Q_X, d);
  }

  template <dim_t D>
  Array1D<real_t> Particles<D>::Uarr(std::size_t d) const {
    return column(RGC_Q_U, d);
  }

  template <dim_t D>
  Array1D<real_t> Particles<D>::Earr(std::size_t d) const {
    return column(RGC_Q_E, d);
  }

  template <dim_t D>
  Array1D<real_t> Particles<D>::Barr(std::size_t d) const {
    return column(RGC_Q_B, d);
  }

  template class Particles<1>;
  template class Particles<2>;
  template class Particles<3>;

  // ------------------------------------------------------------------ bindings
  void define_units(py::module& m) {
    py::class_<EnergyUnits>(m, "EnergyUnits")
      .def_readonly_static("eV", &EnergyUnits::eV)
      .def_readonly_static("MeV", &EnergyUnits::MeV)
      .def_readonly_static("GeV", &EnergyUnits::GeV)
      .def_readonly_static("mec2", &EnergyUnits::mec2)
      .def_readonly_static("mpc2", &EnergyUnits::mpc2);
  }

  void define_spaces(py::module& m) {
    m.def("Linspace", &Linspace, "start"_a, "stop"_a, "num"_a, doc::Linspace);
    m.def("Logspace", &Logspace, "start"_a, "stop"_a, "num"_a, doc::Logspace);
  }

  template <class T>
  static void define_array(py::module& m, const char* suffix) {
    using np_t = py::array_t<T, py::array::c_style | py::array::forcecast>;
    py::class_<Array1D<T>>(m, (std::string("Array1D_") + suffix).c_str())
      .def(py::init<>())
      .def(py::init<const np_t&>(), "arr"_a)
      .def("head", &Array1D<T>::head, "n"_a = 10, "start"_a = 0, doc::Array1D_head)
      .def("__repr__", &Array1D<T>::repr)
      .def("as_array", &Array1D<T>::as_array, doc::Array1D_as_array)
      .def("extent", &Array1D<T>::extent, "d"_a = 0, doc::Array1D_extent)
      .doc() = doc::Array1D_class;
  }

  void define_arrays_and_bins(py::module& m) {
    // the reference names the classes after typeid(T).name(): i, f, d
    define_array<int>(m, "i");
    define_array<float>(m, "f");
    define_array<double>(m, "d");
    using np_t = py::array_t<real_t, py::array::c_style | py::array::forcecast>;
    py::class_<Bins, Array1D<real_t>>(m, "Bins")
      .def(py::init<const Array1D<real_t>&, const std::string&>(), "arr"_a, "unit"_a = "")
      .def(py::init<const std::string&>(), "unit"_a = "")
      .def(py::init<const np_t&, const std::string&>(), "arr"_a, "unit"_a = "")
      .def_readwrite("log_spaced", &Bins::log_spaced)
      .def_readwrite("unit", &Bins::unit)
      .doc() = doc::Bins_class;
    m.def("Linbins", &Linbins, "start"_a, "stop"_a, "num"_a, "unit"_a = "", doc::Linbins);
    m.def("Logbins", &Logbins, "start"_a, "stop"_a, "num"_a, "unit"_a = "", doc::Logbins);
  }

  template <bool LG>
  static void define_tabulated_function(py::module& m) {
    using np_t = py::array_t<real_t, py::array::c_style | py::array::forcecast>;
    py::class_<TabulatedFunction<LG>>(m, LG ? "TabulatedFunction_log" : "TabulatedFunction")
      .def(py::init<const Array1D<real_t>&, const Array1D<real_t>&, real_t>(), "x"_a, "y"_a,
           "yfill"_a = 0.0)
      .def(py::init<const np_t&, const np_t&, real_t>(), "x"_a, "y"_a, "yfill"_a = 0.0)
      .def("xArr", &TabulatedFunction<LG>::xArr)
      .def("yArr", &TabulatedFunction<LG>::yArr)
      .def("nPoints", &TabulatedFunction<LG>::nPoints)
      .def("yFill", &TabulatedFunction<LG>::yFill)
      .def("xMin", &TabulatedFunction<LG>::xMin)
      .def("xMax", &TabulatedFunction<LG>::xMax)
      .doc() = doc::TabulatedFunction_class;
  }

  void define_tabulated_functions(py::module& m) {
    define_tabulated_function<true>(m);
    define_tabulated_function<false>(m);
  }

  void define_generators(py::module& m) {
    py::class_<PlawGenerator>(m, "PlawGenerator")
      .def(py::init<real_t, real_t, real_t>(), "p"_a, "emin"_a = 0.0, "emax"_a = 0.0)
      .def_readonly("p", &PlawGenerator::p)
      .def_readonly("emin", &PlawGenerator::emin)
      .def_readonly("emax", &PlawGenerator::emax)
      .def("compute", &PlawGenerator::compute, "energy_bins"_a, doc::Plaw_compute)
      .doc() = doc::Plaw_class;
    py::class_<BrokenPlawGenerator>(m, "BrokenPlawGenerator")
      .def(py::init<real_t, real_t, real_t, real_t, real_t>(), "e_break"_a, "p1"_a, "p2"_a,
           "emin"_a = 0.0, "emax"_a = 0.0)
      .def_readonly("e_break", &BrokenPlawGenerator::e_break)
      .def_readonly("p1", &BrokenPlawGenerator::p1)
      .def_readonly("p2", &BrokenPlawGenerator::p2)
      .def_readonly("emin", &BrokenPlawGenerator::emin)
      .def_readonly("emax", &BrokenPlawGenerator::emax)
      .def("compute", &BrokenPlawGenerator::compute, "energy_bins"_a, doc::BrokenPlaw_compute)
      .doc() = doc::BrokenPlaw_class;
    py::class_<DeltaGenerator>(m, "DeltaGenerator")
      .def(py::init<real_t, real_t>(), "energy0"_a, "denergy"_a)
      .def_readonly("energy0", &DeltaGenerator::energy0)
      .def_readonly("denergy", &DeltaGenerator::denergy)
      .def("compute", &DeltaGenerator::compute, "energy_bins"_a, doc::Delta_compute)
      .doc() = doc::Delta_class;
    py::class_<TabulatedDistribution>(m, "TabulatedDistribution")
      .def(py::init<const Bins&, const Array1D<real_t>&>(), "bins_energy"_a, "f"_a)
      .def(py::init<const Bins&, const PlawGenerator&>(), "bins_energy"_a, "generator"_a)
      .def(py::init<const Bins&, const BrokenPlawGenerator&>(), "bins_energy"_a, "generator"_a)
      .def(py::init<const Bins&, const DeltaGenerator&>(), "bins_energy"_a, "generator"_a)
      .def("extent", &TabulatedDistribution::extent, doc::TabDist_extent)
      .def("log_spaced", &TabulatedDistribution::log_spaced, doc::TabDist_log_spaced)
      .def("EnergyBins", &TabulatedDistribution::EnergyBins, doc::TabDist_EnergyBins)
      .def("F", &TabulatedDistribution::F, doc::TabDist_F)
      .doc() = doc::TabDist_class;
  }

  template <dim_t D>
  static void define_particles_d(py::module& m) {
    py::class_<Particles<D>>(m, ("Particles_" + std::to_string(D) + "D").c_str())
      .def(py::init<const std::string&>())
      .def("__repr__", &Particles<D>::repr)
      .def("__len__", &Particles<D>::nactive)
      .def("fromArrays", &Particles<D>::fromArrays, "arrays"_a, "append"_a = false,
           doc::Particles_fromArrays)
      .def("setNactive", &Particles<D>::setNactive)
      .def("setIgnoreCoords", &Particles<D>::setIgnoreCoords)
      .def("allocate", &Particles<D>::allocate)
      .def("printHead", &Particles<D>::printHead, "number"_a = 5, "start"_a = 0,
           doc::Particles_printHead)
      .def("is_allocated", &Particles<D>::is_allocated)
      .def("nactive", &Particles<D>::nactive)
      .def("nalloc", &Particles<D>::nalloc)
      .def("label", &Particles<D>::label)
      .def("energyDistribution", &Particles<D>::energyDistribution, "energy_bins"_a,
           "fourvel"_a = true, doc::Particles_energyDistribution)
      .def("X", &Particles<D>::Xarr, "d"_a, doc::Particles_X)
      .def("U", &Particles<D>::Uarr, "d"_a, doc::Particles_U)
      .def("E", &Particles<D>::Earr, "d"_a, doc::Particles_E)
      .def("B", &Particles<D>::Barr, "d"_a, doc::Particles_B)
      .doc() = doc::Particles_class;
  }

  void define_particles(py::module& m) {
    define_particles_d<1>(m);
    define_particles_d<2>(m);
    define_particles_d<3>(m);
  }

} // namespace rgb
