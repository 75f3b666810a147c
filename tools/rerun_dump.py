"""Re-run a case dumped by tests/tools/fuzz_parity.py (gpurun_out/fuzz_fail_*.npz)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from ragnar_b200 import cabi
cabi.init(0)
for f in sys.argv[1:]:
    z = np.load(f)
    U, E, B = list(z["U"]), list(z["E"]), list(z["B"])
    p = cabi.Particles(3).from_columns(U=U, E=E, B=B)
    got = cabi.sync_spectrum_particles(p, z["bins"], *[float(x) for x in z["consts"]])[1]
    want = z["want"]
    fin = np.isfinite(want)
    big = fin & (np.abs(want) >= 1e-6 * np.max(np.abs(want[fin])))
    rel = np.abs(got - want) / np.where(want == 0, 1, np.abs(want))
    print(f, "max rel err", rel[big].max(), "median", np.median(rel[big]), "p99", np.quantile(rel[big], 0.99))
