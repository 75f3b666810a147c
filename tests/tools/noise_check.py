"""Who is closer to the exact value of the reference FORMULA near a zero of the table?
FromDist with photon bins whose x0 = e_syn / e_peak falls into the table's first cells
(F rises linearly from the forced F(xmin) = 0): compares (a) this library, (b) the
reference's float arithmetic (oracle port, bit-exact to the reference) against (c) the
same formula evaluated in float64 on the same float table."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import numpy as np

import oracle
from ragnar_b200 import cabi

cabi.init(0)
port = oracle.port
gb = cabi.logspace(1, 100, 200)
fd = cabi.generator_eval(0, [-2.0, 1.0, 100.0], gb)
bins = cabi.logspace(7.75e-6, 1.04e-4, 2033)
g_syn, e_at = np.float32(1.0175), np.float32(1.2758)
tx, ty = cabi.tabulate_ffunc()
ours = cabi.sync_spectrum_dist(gb, fd, True, bins, float(g_syn), float(e_at))[1]
_, ref = port.sync_spectrum_dist(gb, fd, True, bins, float(g_syn), float(e_at))
# exact evaluation of tabulation.hpp:29-41 / synchrotron.hpp:78-93 in float64
x = tx.astype(np.float64)
y = ty.astype(np.float64)
n = len(x)
e_peak = (e_at * gb * gb / (g_syn * g_syn)).astype(np.float64)  # float32 like the reference, then exact
exact = np.zeros(len(bins))
for g in range(len(gb)):
    if not e_peak[g] > 0:
        continue
    x0 = bins.astype(np.float64) / e_peak[g]
    inside = (x0 >= x[0]) & (x0 < x[-1])
    xi = np.floor((n - 1) * np.abs(np.log10(x0 / x[0])) / np.log10(x[-1] / x[0])).astype(int)
    xi = np.clip(xi, 0, n - 2)
    F = (y[xi + 1] * np.log10(x0 / x[xi]) + y[xi] * np.log10(x[xi + 1] / x0)) / np.log10(x[xi + 1] / x[xi])
    F = np.where(inside, F, 0.0)
    exact += float(fd[g]) * bins.astype(np.float64) * float(gb[g]) * F
big = exact >= 1e-6 * exact.max()
e_ours = np.max(np.abs(ours[big] - exact[big]) / exact[big])
e_ref = np.max(np.abs(ref[big] - exact[big]) / exact[big])
e_pair = np.max(np.abs(ours[big] - ref[big]) / ref[big])
print(f"[noise_check] FromDist, x0 in the table's first cells: max rel dev from the float64 evaluation of "
      f"the reference formula: this library {e_ours:.2e}, the reference's float arithmetic {e_ref:.2e}; "
      f"library vs reference {e_pair:.2e}")

# ---- the same question for the particle path with a mono-energetic population
n = 20000
U = [np.full(n, 30.0, np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)]
E = [np.zeros(n, np.float32)] * 3
B = [np.zeros(n, np.float32), np.full(n, 1.0, np.float32), np.zeros(n, np.float32)]
consts = (2.0847, 1.0092, 59.44)
pbins = cabi.logspace(0.1737, 453805.2, 500)
p = cabi.Particles(3).from_columns(U=U, E=E, B=B)
ours = cabi.sync_spectrum_particles(p, pbins, *consts)[1]
_, ref = port.sync_spectrum_particles(U, E, B, pbins, *consts)
ep, ch = port.sync_epeak_chir([u[:1] for u in U], [e[:1] for e in E], [b[:1] for b in B], *consts)
x0 = pbins.astype(np.float64) / float(ep[0])
inside = (x0 >= x[0]) & (x0 < x[-1])
with np.errstate(all="ignore"):
    xi = np.clip(np.floor((n_tab := len(x)) - 1) * 0 + np.floor((len(x) - 1) * np.abs(np.log10(x0 / x[0])) / np.log10(x[-1] / x[0])), 0, len(x) - 2).astype(int)
    F = (y[xi + 1] * np.log10(x0 / x[xi]) + y[xi] * np.log10(x[xi + 1] / x0)) / np.log10(x[xi + 1] / x[xi])
exact = n * pbins.astype(np.float64) * float(ch[0]) * np.where(inside, F, 0.0)
big = exact >= 1e-6 * exact.max()
print(f"[noise_check] {n} identical particles, 500 bins: max rel dev from the float64 evaluation of the reference "
      f"formula: this library {np.max(np.abs(ours[big] - exact[big]) / exact[big]):.2e}, the reference's float "
      f"arithmetic {np.max(np.abs(ref[big] - exact[big]) / exact[big]):.2e}; library vs reference "
      f"{np.max(np.abs(ours[big] - ref[big]) / ref[big]):.2e}")
