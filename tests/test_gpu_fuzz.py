"""Driver-run slice of the randomised parity sweep (tests/fuzzlib.py) plus the fixed
degenerate populations the round-1 review asked for: mono-energetic, 1 % spread,
n in {2, 100, 4095}, dirty inputs, and the committed fuzz fixtures."""
from pathlib import Path

import numpy as np
import pytest

from tests import fuzzlib, synth

pytestmark = pytest.mark.gpu

GOLDEN = Path(__file__).resolve().parent / "golden" / "fuzz"


@pytest.mark.parametrize("seed", [101, 102])
def test_fuzz_sweep(cabi, port, seed):
    fails, stats = fuzzlib.run(cabi, port, 40, seed)
    assert not fails, fails
    assert stats["literal_cases"] > 0 and stats["hinge_cases"] > 0
    assert stats["worst_literal"] < fuzzlib.LITERAL_RTOL
    assert stats["worst_hinge_stat"] < 1e-5


@pytest.mark.parametrize("kind", ["mono", "narrow", "dirty", "config3", "full3d"])
@pytest.mark.parametrize("n", [2, 100, 4095, 65_537])
def test_degenerate_classes_are_exact(cabi, port, kind, n):
    """small or degenerate populations take the literal path: the reference's float term
    per pair, so every non-zero bin agrees to fp64 summation order"""
    rng = np.random.default_rng(n * 7 + len(kind))
    U, E, B = fuzzlib.make_population(rng, n, kind)
    for M, lo, hi in ((200, 0.01, 1e5), (2033, 1e-4, 1e7), (5, 1.0, 30.0)):
        bins = cabi.logspace(lo, hi, M)
        p = cabi.Particles(3).from_columns(U=U, E=E, B=B)
        _, got = cabi.sync_spectrum_particles(p, bins, 0.7, 3.0, 2.0)
        _, want = port.sync_spectrum_particles(U, E, B, bins, 0.7, 3.0, 2.0)
        p.release()
        why, _ = fuzzlib.check_spectrum(got, want, literal=True, degenerate=True)
        assert not why, (kind, n, M, why)


def test_committed_fuzz_fixtures(cabi, port, monkeypatch):
    """failures of earlier sweeps, kept as fixtures: exact on the literal path; the hinge
    pipeline forced onto them holds its degenerate bar"""
    files = sorted(GOLDEN.glob("*.npz"))
    assert files
    for f in files:
        d = np.load(f)
        U, E, B = list(d["U"]), list(d["E"]), list(d["B"])
        consts = tuple(float(x) for x in d["consts"])
        _, want = port.sync_spectrum_particles(U, E, B, d["bins"], *consts)
        assert np.array_equal(want, d["want"], equal_nan=True), "oracle drifted from the stored reference values"
        p = cabi.Particles(3).from_columns(U=U, E=E, B=B)
        _, got = cabi.sync_spectrum_particles(p, d["bins"], *consts)
        why, _ = fuzzlib.check_spectrum(got, want, literal=len(U[0]) <= fuzzlib.LITERAL_MAX_N, degenerate=True)
        assert not why, (f.name, why)
        monkeypatch.setenv("RGC_LITERAL_MAX_N", "0")
        _, got_h = cabi.sync_spectrum_particles(p, d["bins"], *consts)
        monkeypatch.delenv("RGC_LITERAL_MAX_N")
        why, _ = fuzzlib.check_spectrum(got_h, want, literal=False, degenerate=True)
        assert not why, (f.name, "hinge", why)
        p.release()
