// Tristan-v2 plugin — host-side mirror of the reference's
// src/plugins/tristan-v2.cpp.  The HDF5 parsing and the disk -> pinned host ->
// device streaming live behind rgc_tristan_read_particles.
#include "docstrings.hpp"
#include "ragnar_host.hpp"

using namespace pybind11::literals;

namespace rgb {

  template <dim_t D>
  void TristanV2<D>::setPath(const std::string& path) {
    m_path      = path;
    is_path_set = true;
  }

  template <dim_t D>
  void TristanV2<D>::setStep(std::size_t step) {
    m_step      = step;
    is_step_set = true;
  }

  template <dim_t D>
  std::string TristanV2<D>::getPath() const {
    if (!is_path_set) {
      throw std::runtime_error("Path not set");
    }
    return m_path;
  }

  template <dim_t D>
  std::size_t TristanV2<D>::getStep() const {
    if (!is_step_set) {
      throw std::runtime_error("Step not set");
    }
    return m_step;
  }

  // reference tristan-v2.cpp:95-188
  template <dim_t D>
  Particles<D> TristanV2<D>::readParticles(const std::string& label, unsigned short sp,
                                           std::size_t start, std::size_t size,
                                           std::size_t stride, bool ignore_coordinates) const {
    if (stride == 0) {
      throw std::runtime_error("Stride must be greater than 0");
    } else if (stride != 1 and size != 0) {
      throw std::runtime_error("Size must be determined automatically (0) when stride != 1");
    }
    const auto        step  = std::to_string(getStep());
    const std::string fname = getPath() + "/output/prtl/prtl.tot." +
                              std::string(step.length() < 5 ? 5 - step.length() : 0, '0') + step;
    py::print("Reading particles #", sp, "from", fname, "...", "flush"_a = true);
    rgc_particles_t* handle = nullptr;
    std::size_t      ntotal = 0, nread = 0;
    check(rgc_tristan_read_particles(getPath().c_str(), getStep(), sp, start, size, stride,
                                     ignore_coordinates ? 1 : 0, D, &handle, &ntotal, &nread));
    py::print(" found", human_readable((double)ntotal), "particles, reading",
              human_readable((double)nread), "starting from", start, "flush"_a = true);
    const char* coord[3] = { "x", "y", "z" };
    const char* vel[3]   = { "u", "v", "w" };
    if (not ignore_coordinates) {
      for (auto d = 0u; d < D; ++d) {
        py::print(" ", coord[d], ": OK", "flush"_a = true);
      }
    }
    for (auto d = 0u; d < 3u; ++d) {
      py::print(" ", vel[d], ": OK", "flush"_a = true);
      py::print(" ", std::string("e") + coord[d], ": OK", "flush"_a = true);
      py::print(" ", std::string("b") + coord[d], ": OK", "flush"_a = true);
    }
    return Particles<D>::adopt(label, handle, nread, ignore_coordinates);
  }

  template class TristanV2<1>;
  template class TristanV2<2>;
  template class TristanV2<3>;

  template <dim_t D>
  static void define_tristan_d(py::module& m) {
    py::class_<TristanV2<D>>(m, ("TristanV2_" + std::to_string(D) + "D").c_str())
      .def(py::init<>())
      .def("label", &TristanV2<D>::label)
      .def("setPath", &TristanV2<D>::setPath)
      .def("setStep", &TristanV2<D>::setStep)
      .def("getPath", &TristanV2<D>::getPath)
      .def("getStep", &TristanV2<D>::getStep)
      .def("readParticles", &TristanV2<D>::readParticles, "label"_a, "sp"_a, "start"_a = 0,
           "size"_a = 0, "stride"_a = 1, "ignore_coordinates"_a = false,
           doc::Tristan_readParticles)
      .doc() = doc::Tristan_class;
  }

  void define_tristan(py::module& m) {
    define_tristan_d<1>(m);
    define_tristan_d<2>(m);
    define_tristan_d<3>(m);
  }

} // namespace rgb
