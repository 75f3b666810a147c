"""Generates tests/golden/ic.npz from the reference's OWN code (oracle/_ref: the
unmodified sources of haykh/ragnar @ fceb6b08 on the Kokkos-subset shim).  Run here,
where /root/reference exists; the fixture travels to the GPU box.

    OMP_NUM_THREADS=1 python tests/golden/make_golden_ic.py

ICSpectrum (src/physics/ic.cpp:15-46) is only well defined in the reference when the
soft-photon and IC grids have the same length (ic.cpp:31-34 vs ic.hpp:58), so every
case here has nsoft == nic."""
import contextlib
import io
import os
import sys
from pathlib import Path

os.environ.setdefault("OMP_NUM_THREADS", "1")
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

import oracle  # noqa: E402

OUT = Path(__file__).resolve().parent
rg = oracle.ref()
rg64 = oracle.ref64()


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def case(mod, name):
    if name == "test_ic_log":  # src/tests/ic.py:14-50
        dp = mod.TabulatedDistribution(mod.Logbins(1e3, 1e7, 200), mod.PlawGenerator(-1.5, 1e3, 1e7))
        ds = mod.TabulatedDistribution(mod.Logbins(1e-11, 1e-7, 200, mod.EnergyUnits.mec2),
                                       mod.DeltaGenerator(1e-8, 1e-9))
        b = mod.Bins(mod.Logspace(1e3, 1e7, 200), mod.EnergyUnits.mec2)
    elif name == "plaw_soft_log":  # broad soft-photon field, deep Klein-Nishina included
        dp = mod.TabulatedDistribution(mod.Logbins(1.5, 1e6, 160), mod.PlawGenerator(-2.2, 1.5, 1e6))
        ds = mod.TabulatedDistribution(mod.Logbins(1e-9, 1e-2, 120, mod.EnergyUnits.mec2),
                                       mod.BrokenPlawGenerator(1e-6, 1.0, -2.0, 1e-9, 1e-2))
        b = mod.Logbins(1e-6, 1e6, 120, mod.EnergyUnits.mec2)
    elif name == "lin_prtls":  # linear particle bins: the 1/g^2 branch (ic.hpp:80-83)
        dp = mod.TabulatedDistribution(mod.Linbins(2, 5e3, 300), mod.PlawGenerator(-2.5, 2, 5e3))
        ds = mod.TabulatedDistribution(mod.Logbins(1e-10, 1e-5, 64, mod.EnergyUnits.mec2),
                                       mod.PlawGenerator(-1.0, 1e-10, 1e-5))
        b = mod.Logbins(1e-8, 1e4, 64, mod.EnergyUnits.mec2)
    else:
        raise KeyError(name)
    return dp, ds, b


ic = {}
for name in ("test_ic_log", "plaw_soft_log", "lin_prtls"):
    for mod, tag in ((rg, "f32"), (rg64, "f64")):
        dp, ds, b = case(mod, name)
        ic[f"{name}_spec_{tag}"] = quiet(mod.ICSpectrum, dp, ds, b).as_array()
    ic[f"{name}_g"] = dp.EnergyBins().as_array()
    ic[f"{name}_f"] = dp.F().as_array()
    ic[f"{name}_islog"] = np.array(dp.log_spaced())
    ic[f"{name}_es"] = ds.EnergyBins().as_array()
    ic[f"{name}_fs"] = ds.F().as_array()
    ic[f"{name}_bins"] = b.as_array()
np.savez_compressed(OUT / "ic.npz", **ic)
print("written", OUT / "ic.npz", (OUT / "ic.npz").stat().st_size, "B")
