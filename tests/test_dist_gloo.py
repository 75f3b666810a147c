"""N > 1 host logic on CPU: world_size-2 `gloo` process group.  Covers the
particle-range sharding, the NCCL-unique-id bootstrap (with a recording stand-in
for the C-ABI) and the additivity the single exchange step relies on: per-rank
partial spectra / counts summed over ranks equal the unsharded result.  The
per-rank compute here is the CPU oracle (test infrastructure); on the GPU box
the same plumbing drives libragnar_cuda's own ncclAllReduce."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from ragnar_b200.dist import shard_range  # noqa: E402


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 100, 4_000_000_003):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0
            assert sum(c for _, c in spans) == n
            for (o0, c0), (o1, _) in zip(spans, spans[1:]):
                assert o0 + c0 == o1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


class _RecordingCabi:
    COMM_ID_BYTES = 128

    def __init__(self, rank):
        self.rank = rank
        self.got = None

    def comm_unique_id(self):
        assert self.rank == 0, "only rank 0 asks NCCL for the id"
        return bytes((7 * i + 3) % 256 for i in range(128))

    def comm_init(self, uid, rank, nranks):
        self.got = (uid, rank, nranks)


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from ragnar_b200 import dist as rdist
        from tests import synth

        fake = _RecordingCabi(rank)
        assert rdist.install_communicator(fake, dist) == (rank, world)
        uid, r, n = fake.got
        assert (r, n) == (rank, world) and uid == bytes((7 * i + 3) % 256 for i in range(128))

        n_total = 30_001
        U, E, B = synth.full3d(n_total, seed=5)
        bins = oracle.port.logspace(0.01, 1e5, 200)
        gbins = oracle.port.logspace(1e-2, 1e3, 200)
        off, cnt = rdist.shard_range(n_total, rank, world)
        sl = slice(off, off + cnt)
        _, part = oracle.port.sync_spectrum_particles([u[sl] for u in U], [e[sl] for e in E],
                                                      [b[sl] for b in B], bins, 1.3, 2.0, 0.7)
        _, _, counts = oracle.port.energy_distribution(*[u[sl] for u in U], gbins, False, True)
        total = rdist.allreduce_sum_host(dist, part)
        total_counts = rdist.allreduce_sum_host(dist, counts.astype(np.int64))
        if rank == 0:
            _, whole = oracle.port.sync_spectrum_particles(U, E, B, bins, 1.3, 2.0, 0.7)
            _, _, whole_counts = oracle.port.energy_distribution(*U, gbins, False, True)
            assert np.array_equal(total_counts, whole_counts.astype(np.int64))
            assert total_counts.sum() == n_total
            assert np.allclose(total, whole, rtol=1e-12, atol=0)
        Path(out_dir, f"ok{rank}").write_text("ok")
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
