"""Batched driver (ragnar_b200/pipeline.py): steps x species,
read -> energyDistribution -> SynchrotronSpectrum_3D -> write, with the next species'
read overlapped with the current one's compute."""
import numpy as np
import pytest

from tests import synth
from tests.test_gpu_tristan import _write_step

pytestmark = pytest.mark.gpu


def test_pipeline_matches_oracle_and_sequential(cabi, port, tmp_path):
    from ragnar_b200 import pipeline

    written = {st: _write_step(cabi, tmp_path, st, 120_000 + 1000 * st, 70_000, seed=20 + st)
               for st in (3, 4)}
    pbins = cabi.logspace(1e-3, 1e3, 200)   # legacy/simulation.cpp.bak:136
    gbins = cabi.logspace(1e-1, 200, 200)   # :128
    consts = (0.45 * 0.45 * np.sqrt(100.0) / 2.0, 50.0, (27.0 / 8.0) * 0.1 * 137.0)  # :110-118
    species = [("e-", 1), ("e+", 2)]
    out = tmp_path / "spec.h5"
    rep = pipeline.process_steps(str(tmp_path), [3, 4], species, pbins, gbins, *consts,
                                 out_file=str(out), prefetch=True)
    seq = pipeline.process_steps(str(tmp_path), [3, 4], species, pbins, gbins, *consts,
                                 prefetch=False)
    assert len(rep.results) == 4 and rep.wall_s > 0
    for a, b in zip(rep.results, seq.results):
        assert (a.step, a.label, a.nparticles) == (b.step, b.label, b.nparticles)
        assert np.array_equal(a.spectrum64, b.spectrum64) and np.array_equal(a.distribution, b.distribution)
    for st in (3, 4):
        for (label, sp), cols, off in zip(species, written[st], (3, 0)):
            U, E, B = cols[off:off + 3], cols[off + 3:off + 6], cols[off + 6:off + 9]
            r = rep.by(st, label)
            assert r.nparticles == len(U[0])
            _, want = port.sync_spectrum_particles(U, E, B, pbins, *consts)
            assert synth.rel_err(r.spectrum64, want) < 1e-5
            _, want_h, _ = port.energy_distribution(*U, gbins, True, True)
            nz = want_h > 0
            assert np.max(np.abs(r.distribution[nz] - want_h[nz]) / want_h[nz]) < 1e-5
            # the batched SynchrotronSpectrumFromDist of that distribution (4 items: literal terms)
            _, want_d = port.sync_spectrum_dist(gbins, r.distribution, True, pbins, consts[1], consts[2])
            assert synth.rel_err(r.spectrum_from_dist, want_d.astype(np.float32), floor_frac=0.0) < 1.2e-7
    with cabi.H5File(str(out), "r") as f:
        names = set(f.list("/"))
        assert {"sync_photon_energy_mec2", "sync_intensity_e-_3", "distribution_e+_4",
                "gammaM1_e-_4", "sync_intensity_dist_e-_3"} <= names
        assert np.array_equal(f.read("sync_intensity_dist_e+_4"), rep.by(4, "e+").spectrum_from_dist)
        assert np.array_equal(f.read("sync_intensity_e+_4"), rep.by(4, "e+").spectrum)
        assert np.array_equal(f.read("sync_photon_energy_mec2"), pbins)


def test_range_read_covers_the_tail(cabi, tmp_path):
    """rgc_tristan_read_range: a rank's share may end at the last particle, which the
    reference's selection rejects (tristan-v2.cpp:126-128)"""
    from ragnar_b200 import dist as rdist

    sp1, _ = _write_step(cabi, tmp_path, 1, 10_007, 10, seed=1)
    got = []
    for rank in range(3):
        off, cnt = rdist.shard_range(10_007, rank, 3)
        p, ntotal = cabi.tristan_read_range(str(tmp_path), 1, 1, off, cnt, ignore_coords=True)
        assert ntotal == 10_007 and p.n == cnt
        got.append(p.read(cabi.Q_U, 0, 0, cnt))
    assert np.array_equal(np.concatenate(got), sp1[3])
    with pytest.raises(cabi.RagnarCudaError):
        cabi.tristan_read_range(str(tmp_path), 1, 1, 10_000, 8)
