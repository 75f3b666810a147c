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

  // ------------------------------------------------------------------ io::h5
  // reference src/io/h5.cpp:16-48
  template <class T>
  static Array1D<T> H5Read1DArray(const std::string& filename, const std::string& dsetname,
                                  std::size_t size, std::size_t stride) {
    { // the reference validates before it prints (h5.cpp:21-35)
      rgc_h5_t* f = nullptr;
      check(rgc_h5_open(filename.c_str(), RGC_H5_READONLY, &f));
      int       rank = 0;
      uint64_t  dims[8] {};
      const int rc = rgc_h5_dataset_info(f, dsetname.c_str(), &rank, dims, 8, nullptr, nullptr,
                                         nullptr);
      const std::string msg = rc == RGC_OK ? "" : rgc_last_error();
      rgc_h5_close(f);
      if (rc != RGC_OK) {
        throw std::runtime_error(msg);
      }
      if (rank != 1) {
        throw std::runtime_error("Dataset is not 1D");
      }
      if (stride == 0) {
        throw std::runtime_error("Stride must be greater than 0");
      } else if (size != 0 and dims[0] / stride > size) {
        throw std::runtime_error("Number of read quantity exceeds allocated space");
      }
    }
    py::print("Reading", dsetname, "from", filename, "...", "end"_a = "", "flush"_a = true);
    rgc_buf_t* buf = nullptr;
    if (rgc_h5_read_array(filename.c_str(), dsetname.c_str(), dtype_of<T>(), size, stride, &buf) !=
        RGC_OK) {
      const std::string what = rgc_last_error();
      py::print("Error reading", dsetname, "from", filename, ":", what);
      throw std::runtime_error(what);
    }
    py::print(": OK", "flush"_a = true);
    return Array1D<T>::adopt(buf);
  }

  // reference src/io/h5.cpp:50-68
  template <class T>
  static void H5Write1DArray(const std::string& filename, const std::string& dsetname,
                             const Array1D<T>& array) {
    py::print("Writing", dsetname, "to", filename, "...", "end"_a = "", "flush"_a = true);
    int rc = RGC_OK;
    if (array.handle() != nullptr) {
      rc = rgc_h5_write_array(filename.c_str(), dsetname.c_str(), array.handle());
    } else { // an empty Array1D: a dataset of extent 0
      rgc_h5_t* f = nullptr;
      rc          = rgc_h5_open(filename.c_str(), RGC_H5_READWRITE, &f);
      if (rc == RGC_OK) {
        rc = rgc_h5_create_dataset(f, dsetname.c_str(), dtype_of<T>(), 0);
        const std::string keep = rc == RGC_OK ? "" : rgc_last_error();
        const int         rc2  = rgc_h5_close(f);
        if (rc != RGC_OK) {
          py::print("Error writing", dsetname, "to", filename, ":", keep);
          throw std::runtime_error(keep);
        }
        rc = rc2;
      }
    }
    if (rc != RGC_OK) {
      const std::string what = rgc_last_error();
      py::print("Error writing", dsetname, "to", filename, ":", what);
      throw std::runtime_error(what);
    }
    py::print(": OK", "flush"_a = true);
  }

  template <class T>
  static void define_h5_t(py::module& m, const char* suffix) {
    m.def((std::string("H5read1DArray_") + suffix).c_str(), &H5Read1DArray<T>, "filename"_a,
          "dsetname"_a, "size"_a = 0, "stride"_a = 1, doc::H5read1DArray);
    m.def((std::string("H5write1DArray_") + suffix).c_str(), &H5Write1DArray<T>, "filename"_a,
          "dsetname"_a, "array"_a, doc::H5write1DArray);
  }

  // suffixes are typeid(T).name() in the reference (h5.cpp:72,96): i, f, d
  void define_h5(py::module& m) {
    define_h5_t<int>(m, "i");
    define_h5_t<real_t>(m, "f");
    define_h5_t<double>(m, "d");
  }

  void define_tristan(py::module& m) {
    define_tristan_d<1>(m);
    define_tristan_d<2>(m);
    define_tristan_d<3>(m);
  }

} // namespace rgb
