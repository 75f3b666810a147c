"""Randomised parity sweep of the CUDA path against the CPU oracle (tests/fuzzlib.py);
tests/test_gpu_fuzz.py runs two short seeds of it under `pytest -m gpu`.

    python tests/tools/fuzz_parity.py [cases] [seed]
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import oracle
from ragnar_b200 import cabi
from tests import fuzzlib

ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 150
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cabi.init(0)
fails, stats = fuzzlib.run(cabi, oracle.port, ncases, seed, dump_dir=str(ROOT / "gpurun_out"),
                           log=lambda *a: print(*a, flush=True))
print(f"[fuzz_parity] seed={seed} cases={ncases}: {len(fails)} failures; {stats}", flush=True)
sys.exit(1 if fails else 0)
