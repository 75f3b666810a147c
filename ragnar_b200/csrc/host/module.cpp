// PYBIND11_MODULE(ragnar): same module name, function / class names, argument
// names and defaults as the reference's src/pyinterface.cpp:26-91, so
// `import ragnar as rg` scripts and the reference's own tests run unchanged.
#include "ragnar_host.hpp"

using namespace pybind11::literals;

PYBIND11_MODULE(ragnar, m) {
  m.doc() = "Ragnar: A simple module for radiative post-processing of PIC data";

  // reference pyinterface.cpp:30-49 (Kokkos::initialize / finalize); here: select
  // the CUDA device ($LOCAL_RANK under torchrun, else 0), create streams.
  // Raises RuntimeError without a B200-class device: there is no CPU fallback.
  m.def(
    "Initialize",
    []() {
      if (!rgc_is_initialized()) {
        rgb::check(rgc_init(-1));
      } else {
        py::print("Kokkos is already initialized");
      }
    },
    "Initialize Kokkos");
  m.def(
    "Finalize",
    []() {
      if (rgc_is_initialized()) {
        rgb::check(rgc_finalize());
      } else {
        py::print("Kokkos is not initialized");
      }
    },
    "Finalize Kokkos");

  // registration order as in the reference (Array1D_f before Bins)
  rgb::define_units(m);
  rgb::define_spaces(m);
  rgb::define_tabulated_functions(m);
  rgb::define_arrays_and_bins(m);
  rgb::define_particles(m);
  rgb::define_h5(m);
  rgb::define_tristan(m);
  rgb::define_generators(m);
  rgb::define_synchrotron(m);
  rgb::define_ic(m);
}
