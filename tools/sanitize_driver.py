"""Small pass over every kernel family of the library, for compute-sanitizer:
    compute-sanitizer --tool racecheck python tools/sanitize_driver.py
    compute-sanitizer --tool memcheck  python tools/sanitize_driver.py
(and under torchrun on 2 GPUs for the peer-store exchange kernel).  Results are checked
against loose invariants only; parity is the tests' job."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np

from ragnar_b200 import cabi

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
cabi.init(local)
if world > 1:
    import torch
    import torch.distributed as dist

    from ragnar_b200 import dist as rdist

    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rdist.install_communicator(cabi, dist)
    print(f"[sanitize] rank {rank}/{world}: exchange = {cabi.comm_exchange_kind()}", flush=True)

n = 150_000
p = cabi.Particles(3).allocate(n).generate(1, 7, rank * n, 0, n, 0.05, 500.0)
gb = cabi.logspace(1e-2, 1e3, 200)
hist, counts, _ = cabi.energy_histogram(p, gb, log_spaced=True, fourvel=True)       # energy_hist_kernel
assert counts.sum() == world * n
os.environ["RGC_LITERAL_MAX_N"] = "0"                                                # hinge pipeline
for M, lo, hi in ((200, 0.01, 1e5), (1000, 1e-3, 1e6)):
    s = cabi.sync_spectrum_particles(p, cabi.logspace(lo, hi, M), 1.0, 1.0, 1.0)[1]  # prologue, colscan, sort,
    assert np.all(np.isfinite(s)) and s.max() > 0                                   # pair, moments, final
os.environ["RGC_SORT_RANK"] = "ballot"
s = cabi.sync_spectrum_particles(p, cabi.logspace(0.01, 1e5, 200), 1.0, 1.0, 1.0)[1]  # sync_sort_kernel<false>
os.environ.pop("RGC_SORT_RANK")
os.environ["RGC_SPECTRUM_PATH"] = "gather"
s = cabi.sync_spectrum_particles(p, cabi.logspace(0.01, 1e5, 200), 1.0, 1.0, 1.0)[1]  # gather fallback kernel
os.environ.pop("RGC_SPECTRUM_PATH")
os.environ.pop("RGC_LITERAL_MAX_N")
s = cabi.sync_spectrum_particles(p, cabi.logspace(0.01, 1e5, 200), 1.0, 1.0, 1.0, nactive=5000)[1]  # literal
if rank == 0:
    g = cabi.logspace(1, 100, 200)
    f = np.tile(cabi.generator_eval(0, [-2.0, 1.0, 100.0], g), (70, 1))
    for mode in (0, 1):
        cabi.sync_spectrum_dist_batch(g, f, True, cabi.logspace(0.01, 1e7, 200), 1.0, 1.0, mode=mode)
    cabi.ic_spectrum(g, f[0], True, np.geomspace(1e-9, 1e-4, 64).astype(np.float32),
                     np.ones(64, np.float32), np.geomspace(1e-3, 1e3, 128).astype(np.float32))
    x = cabi.logspace_device(1e-3, 1e3, 300_000)
    y = cabi.linspace_device(0.0, 1.0, 300_000)
    cabi.tabulated_eval(True, x, y, cabi.DeviceArray.from_host(np.geomspace(1e-4, 1e4, 10_000))).to_host()
    x.minmax()
cabi.synchronize()
print(f"[sanitize] rank {rank}: {cabi.launch_count()} kernel launches, all entry points returned", flush=True)
if world > 1:
    dist.barrier()
    cabi.comm_destroy()
    dist.destroy_process_group()
