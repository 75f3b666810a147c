"""Generates tests/golden/*.npz from the reference's OWN code (oracle/_ref: the
unmodified sources of haykh/ragnar @ fceb6b08 compiled on the Kokkos-subset shim
by oracle/build_ref.sh).  Run here, where /root/reference exists; the fixtures
travel to the GPU box, the reference does not.

    OMP_NUM_THREADS=1 python tests/golden/make_golden.py

OMP_NUM_THREADS=1 makes the reference's float ScatterView accumulation
deterministic (== Kokkos Serial order)."""
import contextlib
import io
import json
import os
import sys
from pathlib import Path

os.environ.setdefault("OMP_NUM_THREADS", "1")
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

import oracle  # noqa: E402
from tests import synth  # noqa: E402

OUT = Path(__file__).resolve().parent
rg = oracle.ref()
rg64 = oracle.ref64()


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def prtls_of(mod, U, E, B):
    p = mod.Particles_3D("golden")
    p.fromArrays({f"{q}{d + 1}": a[d] for q, a in (("U", U), ("E", E), ("B", B)) for d in range(3)})
    return p


# ---- spaces, F(x), table, generators, interpolation --------------------------
spaces = {}
SPACE_CASES = [(1e-2, 1e3, 200), (1e-6, 100, 200), (0.01, 1e7, 200), (1, 100, 200),
               (10**-2.5, 10**2.5, 213), (0.01, 1e5, 200), (1e-3, 1e6, 1000), (1, 1e3, 22),
               (55.0, 56.0, 123), (3.0, 7.0, 1)]
for i, (a, b, n) in enumerate(SPACE_CASES):
    spaces[f"log_{i}"] = rg.Logspace(a, b, n).as_array()
    spaces[f"lin_{i}"] = rg.Linspace(a, b, n).as_array()
spaces["cases"] = np.array(SPACE_CASES, np.float64)
np.savez_compressed(OUT / "spaces.npz", **spaces)

xs = np.concatenate([np.logspace(-7, 1.29, 60), [1.0, 0.29, 1e-6, 25.0, 1e-5, 0.9999e-5, 19.99,
                                                 20.5, 3.3e-3, 7.7]]).astype(np.float32)
fvals = np.array([rg.Ffunc_integrand(float(x)) for x in xs], np.float32)
# the table the reference rebuilds on every spectrum call (TabulateFfunc is not bound to
# Python; its nodes are Logspace(1e-6, 100, 200) and its values Ffunc_integrand, y[0] forced 0)
tab_x = rg.Logspace(1e-6, 100, 200).as_array()
tab_y = np.array([0.0 if float(x) < 1e-6 else rg.Ffunc_integrand(float(x)) for x in tab_x],
                 np.float32)
tf = rg.TabulatedFunction_log(tab_x, tab_y)
np.savez_compressed(OUT / "ffunc.npz", x=xs, f=fvals, tab_x=tab_x, tab_y=tab_y,
                    xmin=np.float32(tf.xMin()), xmax=np.float32(tf.xMax()))

gen = {}
e_log = rg.Logbins(1e-3, 10, 500)
e_lin = rg.Linbins(1e-3, 10, 500)
for name, bins in (("log", e_log), ("lin", e_lin)):
    gen[f"bins_{name}"] = bins.as_array()
    gen[f"plaw_{name}"] = rg.PlawGenerator(-1.2, 1e-2, 1).compute(bins).as_array()
    gen[f"plaw_inf_{name}"] = rg.PlawGenerator(-2.5, 0.5).compute(bins).as_array()
    gen[f"plaw_m1_{name}"] = rg.PlawGenerator(-1.0, 1e-2, 5).compute(bins).as_array()
    gen[f"broken_{name}"] = rg.BrokenPlawGenerator(0.3, 0.23, -1.0, 1e-2, 2).compute(bins).as_array()
    gen[f"broken_inf_{name}"] = rg.BrokenPlawGenerator(0.3, 1.5, -2.2).compute(bins).as_array()
    gen[f"delta_{name}"] = rg.DeltaGenerator(2e-2, 0.01).compute(bins).as_array()
np.savez_compressed(OUT / "generators.npz", **gen)

# ---- SynchrotronSpectrumFromDist ------------------------------------------------
fd = {}
FD_CASES = {
    "config1": (("log", 1, 100, 200), (-2, 1, 100), (0.01, 1e7, 200)),
    "sync_log": (("log", 1, 1000, 200), (-2.23, 1, 1000), (0.01, 1e7, 200)),
    "sync_lin": (("lin", 1, 1000, 10000), (-2.5, 1, 1000), (0.01, 1e6, 500)),
}
for name, ((kind, lo, hi, n), plaw, (blo, bhi, bn)) in FD_CASES.items():
    for mod, tag in ((rg, "f32"), (rg64, "f64")):
        pb = (mod.Logbins if kind == "log" else mod.Linbins)(lo, hi, n)
        dist = mod.TabulatedDistribution(pb, mod.PlawGenerator(*plaw))
        bins = mod.Logbins(blo, bhi, bn, "mec2")
        fd[f"{name}_spec_{tag}"] = quiet(mod.SynchrotronSpectrumFromDist, dist, bins, 1, 1).as_array()
    fd[f"{name}_gbeta"] = dist.EnergyBins().as_array()
    fd[f"{name}_f"] = dist.F().as_array()
    fd[f"{name}_bins"] = bins.as_array()
    fd[f"{name}_islog"] = np.array(kind == "log")
np.savez_compressed(OUT / "fromdist.npz", **fd)

# ---- SynchrotronSpectrum_3D and energyDistribution on small seeded populations ----
pr = {}
N = 4000
POPS = {"config3": synth.config3(N, seed=123), "full3d": synth.full3d(N, seed=321)}
CONSTS = {"config3": (1.0, 1.0, 1.0), "full3d": (1.3, 2.0, 0.7)}
for name, (U, E, B) in POPS.items():
    for q, arr in (("U", U), ("E", E), ("B", B)):
        for d in range(3):
            pr[f"{name}_{q}{d + 1}"] = arr[d]
    pr[f"{name}_consts"] = np.array(CONSTS[name], np.float32)
    for mod, tag in ((rg, "f32"), (rg64, "f64")):
        p = prtls_of(mod, U, E, B)
        bins = mod.Logbins(0.01, 1e5, 200, "mec2")
        pr[f"{name}_spec_{tag}"] = quiet(mod.SynchrotronSpectrum_3D, p, bins, *CONSTS[name]).as_array()
        for fourvel in (True, False):
            for logsp in (True, False):
                gb = mod.Logbins(1e-2, 1e3, 200)
                gb.log_spaced = logsp
                key = f"{name}_hist_{'u' if fourvel else 'g'}_{'log' if logsp else 'lin'}_{tag}"
                pr[key] = quiet(p.energyDistribution, gb, fourvel).F().as_array()
    pr[f"{name}_bins"] = bins.as_array()
    pr[f"{name}_gbins"] = gb.as_array()
# the survey's probe vector around the clamps
u = np.array([0.001, 0.5, 1, 5, 999, 1000, 5000, 0.0099999], np.float32)
p = rg.Particles_3D("probe")
p.fromArrays({"U1": u})
b6 = rg.Logbins(1e-2, 1e3, 6)
pr["probe_u"] = u
pr["probe_bins"] = b6.as_array()
pr["probe_hist_log"] = quiet(p.energyDistribution, b6).F().as_array()
b6.log_spaced = False
pr["probe_hist_lin"] = quiet(p.energyDistribution, b6).F().as_array()
np.savez_compressed(OUT / "particles.npz", **pr)

# ---- API surface: names and docstrings of the reference module -------------------
api, docs = {}, {}
for name in sorted(n for n in dir(rg) if not n.startswith("_")):
    obj = getattr(rg, name)
    doc = (obj.__doc__ or "").replace("ragnar_ref.", "ragnar.")
    api[name] = {"kind": "class" if isinstance(obj, type) else "function"}
    docs[name] = {"doc": doc}
    if isinstance(obj, type):
        members = [m for m in sorted(dir(obj)) if not m.startswith("__")]
        api[name]["members"] = members
        docs[name]["members"] = {
            m: (getattr(obj, m).__doc__ or "").replace("ragnar_ref.", "ragnar.") for m in members
            if callable(getattr(obj, m))}
(OUT / "api_surface.json").write_text(json.dumps(api, indent=1, sort_keys=True))
(OUT / "api_docs.json").write_text(json.dumps(docs, indent=1, sort_keys=True))
print("golden fixtures written to", OUT)
for f in sorted(OUT.glob("*.np*")) + sorted(OUT.glob("*.json")):
    print(f"  {f.name:20s} {f.stat().st_size / 1024:8.1f} KiB")
