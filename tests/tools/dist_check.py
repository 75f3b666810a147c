"""Multi-GPU parity check (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/tools/dist_check.py

Every rank owns shard_range(n, rank, world) of one seeded host population; the NCCL
all-reduced spectrum / histogram of the sharded run must match the CPU oracle on the
whole population (spectrum <= 1e-5 per bin, counts bit-exact and identical for every
world size) and a one-rank run of the same library (<= 1e-5: the hinge sums are float
per run of a <= 4096-entry piece, the pieces depend on the partition, and shards of up to
2^19 particles take the literal path)."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import numpy as np
import torch
import torch.distributed as dist

import oracle
from ragnar_b200 import cabi
from ragnar_b200 import dist as rdist
from tests import synth

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cabi.init(local)
n = 1_000_003
U, E, B = synth.full3d(n, seed=17)
bins = cabi.logspace(1e-3, 1e6, 1000)
gbins = cabi.logspace(1e-2, 1e3, 200)
consts = (1.3, 2.0, 0.7)

# one-rank reference run of the same library, before the communicator exists
if rank == 0:
    p_all = cabi.Particles(3).from_columns(U=U, E=E, B=B)
    solo_spec = cabi.sync_spectrum_particles(p_all, bins, *consts)[1]
    solo_hist = cabi.energy_histogram(p_all, gbins, log_spaced=True, fourvel=True)
    p_all.release()

rdist.install_communicator(cabi, dist)
lo, cnt = rdist.shard_range(n, rank, world)
hi = lo + cnt
p = cabi.Particles(3).from_columns(U=[u[lo:hi] for u in U], E=[e[lo:hi] for e in E],
                                   B=[b[lo:hi] for b in B])
spec = cabi.sync_spectrum_particles(p, bins, *consts)[1]
hist, counts, h64 = cabi.energy_histogram(p, gbins, log_spaced=True, fourvel=True)
# a rank that owns no particles still joins the exchange
empty = cabi.Particles(3).allocate(16)
empty.n = 0
spec0 = cabi.sync_spectrum_particles(p if rank == 0 else empty, bins, *consts)[1]
ok = True
if rank == 0:
    _, want = oracle.port.sync_spectrum_particles(U, E, B, bins, *consts)
    big = want >= 1e-6 * want.max()
    err = float(np.max(np.abs(spec[big] - want[big]) / want[big]))
    _, want_h64, want_c = oracle.port.energy_distribution(*U, gbins, True, True)
    nz = want_h64 > 0
    herr = float(np.max(np.abs(h64[nz] - want_h64[nz]) / want_h64[nz]))
    add = float(np.max(np.abs(spec[big] - solo_spec[big]) / solo_spec[big]))
    # shards of <= 2^19 particles take the literal path while the one-rank run of the whole
    # population takes the hinge pipeline: then the two differ like the pipeline and the oracle
    same_path = (cnt <= (1 << 19)) == (n <= (1 << 19))
    ok = (err < 1e-5 and herr < 1e-5 and np.array_equal(counts, want_c)
          and np.array_equal(counts, solo_hist[1]) and add < (1e-6 if same_path else 1e-5)
          and np.array_equal(spec == 0, want == 0) and np.all(np.isfinite(spec0)))
    print(f"[dist_check] world={world} exchange: {cabi.comm_exchange_kind()}", flush=True)
    print(f"[dist_check] world={world} n={n}: spectrum rel err vs oracle {err:.2e}, "
          f"weighted hist {herr:.2e}, counts bit-exact={np.array_equal(counts, want_c)}, "
          f"sharded vs one-rank spectrum {add:.2e} -> {'OK' if ok else 'FAIL'}", flush=True)
# ---- the batched driver with sharded range reads (ragnar_b200/pipeline.py)
import shutil
import tempfile

from ragnar_b200 import pipeline

root = Path(tempfile.gettempdir()) / "rgc_dist_check_pipeline"
if rank == 0:
    shutil.rmtree(root, ignore_errors=True)
    (root / "output" / "prtl").mkdir(parents=True)
    m = 200_003
    cabi.tristan_write_species(str(root), 7, 1, [c[:m] for q in (U, E, B) for c in q], with_coords=False,
                               append=False)
dist.barrier()
m = 200_003
rep = pipeline.process_steps(str(root), [7], [("e-", 1)], bins, gbins, *consts, rank=rank, world=world)
if rank == 0:
    r = rep.results[0]
    _, want_p = oracle.port.sync_spectrum_particles([u[:m] for u in U], [e[:m] for e in E], [b[:m] for b in B],
                                                    bins, *consts)
    bigp = want_p >= 1e-6 * want_p.max()
    perr = float(np.max(np.abs(r.spectrum64[bigp] - want_p[bigp]) / want_p[bigp]))
    _, want_ph, _ = oracle.port.energy_distribution(*[u[:m] for u in U], gbins, True, True)
    nzp = want_ph > 0
    pherr = float(np.max(np.abs(r.distribution[nzp] - want_ph[nzp]) / want_ph[nzp]))
    lo_r, cnt_r = rdist.shard_range(m, rank, world)
    okp = perr < 1e-5 and pherr < 1e-5 and r.nparticles == cnt_r
    ok = ok and okp
    print(f"[dist_check] world={world} pipeline (sharded reads of {m} particles): spectrum rel err {perr:.2e}, "
          f"distribution {pherr:.2e}, rank-0 share {r.nparticles} -> {'OK' if okp else 'FAIL'}", flush=True)
dist.barrier()
if rank == 0:
    shutil.rmtree(root, ignore_errors=True)
# ---- a starved exchange must fail loudly, never return a partial sum: the last rank
# arrives 3 s late at a call whose exchange waits 300 ms at most
if cabi.comm_exchange_kind().startswith("peer-store"):
    import time

    os.environ["RGC_XCHG_TIMEOUT_MS"] = "300"
    dist.barrier()
    late = rank == world - 1
    if late:
        time.sleep(3.0)
    try:
        got = cabi.sync_spectrum_particles(p, bins, *consts)[1]
        raised = False
    except RuntimeError as e:
        raised = "timed out" in str(e)
        got = None
    # early ranks must raise; the late rank finds every peer's data in place and gets the full sum
    ok_t = raised if not late else (got is not None and np.array_equal(got, spec))
    t = torch.tensor([0 if ok_t else 1], device="cuda")
    dist.all_reduce(t)
    # afterwards the exchange refuses further calls until the communicator is re-created
    refused = True
    if not late:
        try:
            cabi.sync_spectrum_particles(p, bins, *consts)
            refused = False
        except RuntimeError:
            pass
    dist.barrier()
    os.environ.pop("RGC_XCHG_TIMEOUT_MS")
    cabi.comm_destroy()
    rdist.install_communicator(cabi, dist)
    again = cabi.sync_spectrum_particles(p, bins, *consts)[1]
    ok_t2 = int(t.item()) == 0 and refused and np.array_equal(again, spec)
    if rank == 0:
        print(f"[dist_check] world={world} starved exchange: early ranks raised RGC_ERR_NCCL, late rank "
              f"summed, exchange refused until re-created, result after re-creation bit-identical -> "
              f"{'OK' if ok_t2 else 'FAIL'}", flush=True)
    ok = ok and ok_t2
flag = torch.tensor([0 if ok else 1], device="cuda")
dist.all_reduce(flag)
dist.barrier()
dist.destroy_process_group()
sys.exit(int(flag.item() != 0))
