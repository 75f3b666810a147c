"""rgc_glibc_log10f.cuh — the device restatement of glibc's log10f that the literal pair
kernel relies on — compiled for the host and compared with the host libm, bit for bit,
on every 251st float bit pattern (tools/check_log10f.cpp; stride 1 = all 2^32, ~20 s)."""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_log10f_restatement_matches_libm(tmp_path):
    exe = tmp_path / "check_log10f"
    subprocess.run(["g++", "-O2", "-fopenmp", "-ffp-contract=off", "-fno-builtin",
                    str(ROOT / "tools" / "check_log10f.cpp"), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe), "251"], capture_output=True, text=True)
    sys.stdout.write(out.stdout)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "log10f mismatches 0, logf mismatches 0" in out.stdout
