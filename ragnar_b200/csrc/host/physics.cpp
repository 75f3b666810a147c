// Synchrotron spectrum entry points — host-side mirror of the reference's
// src/physics/synchrotron.cpp drivers: same banners, same argument checks, the
// compute itself is one C-ABI call into the sm_100a kernels.
#include "docstrings.hpp"
#include "ragnar_host.hpp"

using namespace pybind11::literals;

namespace rgb {

  real_t Ffunc_integrand(real_t x) {
    real_t out = 0.0f;
    check(rgc_sync_ffunc_integrand(x, &out));
    return out;
  }

  // reference synchrotron.cpp:48-65 — built on the host (libstdc++ cyl_bessel_k)
  // once per (npoints, xmin, xmax) and cached inside libragnar_cuda
  TabulatedFunction<true> TabulateFfunc(std::size_t npoints, real_t xmin, real_t xmax) {
    std::vector<real_t> xs(npoints), ys(npoints);
    check(rgc_sync_tabulate_ffunc(npoints, xmin, xmax, xs.data(), ys.data()));
    return TabulatedFunction<true> { Array1D<real_t> { xs }, Array1D<real_t> { ys } };
  }

  namespace {
    struct FTable {
      std::vector<real_t> x, y;
    };

    const FTable& default_ftable() {
      static FTable tab = [] {
        FTable t;
        t.x.resize(200);
        t.y.resize(200);
        check(rgc_sync_tabulate_ffunc(200, static_cast<real_t>(1e-6), static_cast<real_t>(100),
                                      t.x.data(), t.y.data()));
        return t;
      }();
      return tab;
    }
  } // namespace

  // reference synchrotron.cpp:69-105
  Array1D<real_t> SynchrotronSpectrumFromDist(const TabulatedDistribution& dist_prtls,
                                              const Bins& bins_e_syn, real_t g_syn,
                                              real_t e_syn_at_g_syn) {
    py::print("Computing synchrotron spectrum from a distribution", "flush"_a = true);
    const auto& tab   = default_ftable();
    const auto  nbins = bins_e_syn.extent(0);
    py::print(" Launching", human_readable((double)(dist_prtls.extent() * nbins)), "threads",
              "end"_a = "", "flush"_a = true);
    std::vector<real_t> spec(nbins, 0.0f);
    check(rgc_sync_spectrum_dist(dist_prtls.EnergyBins().host_data(), dist_prtls.F().host_data(),
                                 dist_prtls.extent(), dist_prtls.log_spaced() ? 1 : 0,
                                 bins_e_syn.host_data(), nbins, tab.x.data(), tab.y.data(),
                                 tab.x.size(), g_syn, e_syn_at_g_syn, spec.data(), nullptr));
    py::print(": OK", "flush"_a = true);
    return Array1D<real_t> { spec };
  }

  // reference synchrotron.cpp:107-145; the unit check of sync::Kernel's constructor
  // (synchrotron.hpp:139-142) fires after the " Launching" line there too
  template <dim_t D>
  Array1D<real_t> SynchrotronSpectrum(const Particles<D>& prtls, const Bins& bins_e_syn,
                                      real_t B0, real_t g_syn, real_t e_syn_at_g_syn) {
    py::print("Computing synchrotron spectrum for", prtls.label(), "flush"_a = true);
    const auto& tab   = default_ftable();
    const auto  nbins = bins_e_syn.extent(0);
    py::print(" Launching", human_readable((double)(prtls.nactive() * nbins)), "threads",
              "end"_a = "", "flush"_a = true);
    if (bins_e_syn.unit != EnergyUnits::mec2 and bins_e_syn.unit != EnergyUnits::mpc2) {
      throw std::runtime_error("bins_e_syn must be in units of mc^2");
    }
    std::vector<real_t> spec(nbins, 0.0f);
    // unallocated particles are an empty shard: the call still joins a multi-rank all-reduce
    const bool alloc = prtls.is_allocated();
    check(rgc_sync_spectrum_particles(alloc ? prtls.handle() : nullptr, alloc ? prtls.nactive() : 0,
                                      bins_e_syn.host_data(), nbins, tab.x.data(), tab.y.data(),
                                      tab.x.size(), B0, g_syn, e_syn_at_g_syn, spec.data(), nullptr));
    py::print(": OK", "flush"_a = true);
    return Array1D<real_t> { spec };
  }

  template Array1D<real_t> SynchrotronSpectrum(const Particles<1>&, const Bins&, real_t, real_t,
                                               real_t);
  template Array1D<real_t> SynchrotronSpectrum(const Particles<2>&, const Bins&, real_t, real_t,
                                               real_t);
  template Array1D<real_t> SynchrotronSpectrum(const Particles<3>&, const Bins&, real_t, real_t,
                                               real_t);

  void define_synchrotron(py::module& m) {
    m.def("SynchrotronSpectrum_1D", &SynchrotronSpectrum<1>, "prtls"_a, "bins_e_syn"_a, "B0"_a,
          "g_syn"_a, "e_syn_at_g_syn"_a, doc::SynchrotronSpectrum);
    m.def("SynchrotronSpectrum_2D", &SynchrotronSpectrum<2>, "prtls"_a, "bins_e_syn"_a, "B0"_a,
          "g_syn"_a, "e_syn_at_g_syn"_a, doc::SynchrotronSpectrum);
    m.def("SynchrotronSpectrum_3D", &SynchrotronSpectrum<3>, "prtls"_a, "bins_e_syn"_a, "B0"_a,
          "g_syn"_a, "e_syn_at_g_syn"_a, doc::SynchrotronSpectrum);
    m.def("Ffunc_integrand", &Ffunc_integrand, "x"_a);
    m.def("SynchrotronSpectrumFromDist", &SynchrotronSpectrumFromDist, "dist_prtls"_a,
          "bins_e_syn"_a, "g_syn"_a, "e_syn_at_g_syn"_a, doc::SynchrotronSpectrumFromDist);
  }

  // reference src/physics/ic.cpp:15-46.  ic::Kernel's constructor (ic.hpp:48-55) is
  // evaluated as an argument of parallel_for, i.e. after the " Launching" line: the
  // unit errors surface there too.
  Array1D<real_t> ICSpectrum(const TabulatedDistribution& dist_prtls,
                             const TabulatedDistribution& dist_soft_photons,
                             const Bins&                  bins_e_ic) {
    py::print("Computing IC spectrum ...", "flush"_a = true);
    const auto nbins_prtls        = dist_prtls.extent();
    const auto nbins_soft_photons = dist_soft_photons.extent();
    const auto nbins_ic           = bins_e_ic.extent(0);
    py::print(" Launching",
              human_readable((double)(nbins_prtls * nbins_soft_photons * nbins_ic)), "threads",
              "end"_a = "", "flush"_a = true);
    if (dist_soft_photons.EnergyBins().unit != EnergyUnits::mec2) {
      throw std::runtime_error("Soft photons energy bins must be in units of mec^2");
    }
    if (bins_e_ic.unit != EnergyUnits::mec2) {
      throw std::runtime_error("E_ic must be in units of mec^2");
    }
    std::vector<real_t> spec(nbins_ic, 0.0f);
    check(rgc_ic_spectrum(dist_prtls.EnergyBins().host_data(), dist_prtls.F().host_data(),
                          nbins_prtls, dist_prtls.log_spaced() ? 1 : 0,
                          dist_soft_photons.EnergyBins().host_data(),
                          dist_soft_photons.F().host_data(), nbins_soft_photons,
                          bins_e_ic.host_data(), nbins_ic, spec.data(), nullptr));
    py::print(": OK", "flush"_a = true);
    return Array1D<real_t> { spec };
  }

  void define_ic(py::module& m) {
    m.def("ICSpectrum", &ICSpectrum, "dist_prtls"_a, "dist_soft_photons"_a, "bins_e_ic"_a,
          doc::ICSpectrum);
  }

} // namespace rgb
