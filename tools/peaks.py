#!/usr/bin/env python
"""Prints the on-device roofline denominators (rgc_measure_peak) as one JSON line."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ragnar_b200 import cabi  # noqa: E402

cabi.init(0)
sm = cabi.device_info()[1]
out = {"sm_count": sm}
for name, kind in (("ffma_gflops", cabi.PEAK_FFMA), ("pair_gevals", cabi.PEAK_PAIR),
                   ("lds64_gbs", cabi.PEAK_LDS64), ("hbm_read_gbs", cabi.PEAK_HBM_READ),
                   ("int_ginstr", cabi.PEAK_INT)):
    out[name] = [round(cabi.measure_peak(kind), 1) for _ in range(3)]
# per SM per clock at the nominal 1965 MHz, for orientation
clk = 1965e6
out["ffma_per_clk_sm"] = max(out["ffma_gflops"]) * 1e9 / 2 / sm / clk
out["pair_evals_per_clk_sm"] = max(out["pair_gevals"]) * 1e9 / sm / clk
out["lds_bytes_per_clk_sm"] = max(out["lds64_gbs"]) * 1e9 / sm / clk
out["int_instr_per_clk_sm"] = max(out["int_ginstr"]) * 1e9 / sm / clk
print(json.dumps(out))
