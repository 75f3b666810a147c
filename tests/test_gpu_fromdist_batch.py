"""SURVEY 8f row f3, second half: a batch of distributions on shared bins (steps x species of
a run) through SynchrotronSpectrumFromDist — literal terms per item, or the kernel matrix
built once and contracted with the batch in fp64 (reference src/physics/synchrotron.hpp:72-95,
legacy/simulation.cpp.bak:67-219)."""
import numpy as np
import pytest

from tests import synth

pytestmark = pytest.mark.gpu


def _batch(cabi, nbatch, G, seed=0):
    rng = np.random.default_rng(seed)
    gb = cabi.logspace(1.0, 1e3, G)
    rows = []
    for b in range(nbatch):
        f = cabi.generator_eval(0, [float(rng.uniform(-3.5, -1.2)), 1.0, float(rng.uniform(50, 1e3))], gb)
        rows.append(f * np.float32(rng.uniform(0.1, 10)))
    return gb, np.stack(rows)


@pytest.mark.parametrize("islog", [True, False])
@pytest.mark.parametrize("nbatch,G,M", [(1, 200, 200), (16, 200, 200), (70, 333, 517), (256, 64, 1000)])
def test_batch_matches_oracle_per_item(cabi, port, islog, nbatch, G, M):
    gb, fb = _batch(cabi, nbatch, G, seed=nbatch)
    bins = cabi.logspace(0.01, 1e7, M)
    _, lit = cabi.sync_spectrum_dist_batch(gb, fb, islog, bins, 2.0, 0.7, mode=0)
    _, con = cabi.sync_spectrum_dist_batch(gb, fb, islog, bins, 2.0, 0.7, mode=1)
    for b in range(nbatch):
        _, want = port.sync_spectrum_dist(gb, fb[b], islog, bins, 2.0, 0.7)
        # literal: the reference's float terms, fp64 sums in distribution order on both sides
        assert synth.rel_err(lit[b], want, floor_frac=0.0) < 1e-11
        assert np.array_equal(lit[b] == 0, want == 0)
        # contraction: each term within three float roundings of the reference's
        assert synth.rel_err(con[b], want, floor_frac=0.0) < 5e-7
        assert np.array_equal(con[b] == 0, want == 0)
    # one item of the batch == the single-distribution entry point (same terms; the sources are
    # cut into slices by the size of the launch, so only the fp64 summation grouping differs)
    _, single = cabi.sync_spectrum_dist(gb, fb[nbatch // 2], islog, bins, 2.0, 0.7)
    assert synth.rel_err(single, lit[nbatch // 2], floor_frac=0.0) < 1e-13
    assert np.array_equal(single == 0, lit[nbatch // 2] == 0)


def test_batch_auto_mode_and_skipped_sources(cabi, port):
    gb, fb = _batch(cabi, 80, 100, seed=3)
    gb = gb.copy()
    gb[5] = 0.0       # e_peak = 0 fails `e_peak > 0`: the source is skipped ...
    fb[:, 5] = np.inf  # ... whatever its weight holds
    gb[9] = np.nan
    bins = cabi.logspace(0.01, 1e6, 128)
    _, auto = cabi.sync_spectrum_dist_batch(gb, fb, True, bins, 1.0, 1.0)
    _, con = cabi.sync_spectrum_dist_batch(gb, fb, True, bins, 1.0, 1.0, mode=1)
    assert np.array_equal(auto, con), "80 distributions: automatic mode takes the contraction"
    for b in (0, 17, 79):
        _, want = port.sync_spectrum_dist(gb, fb[b], True, bins, 1.0, 1.0)
        assert np.all(np.isfinite(want)) and synth.rel_err(con[b], want, floor_frac=0.0) < 5e-7
