"""Batched driver = the caller of the hot path: loop over simulation
steps and species, read -> energy histogram -> synchrotron spectrum -> write, as the
reference's retired driver did (legacy/simulation.cpp.bak:24-41,67-219: per species
`readParticles`, `computeSyncSpectrum` -> dataset `sync_intensity_<label>`, the particle
distribution -> `gammaM1_<label>` / `distribution_<label>`, the photon grid ->
`sync_photon_energy_mec2`).

B200 shape of it: everything goes through the C-ABI (`ragnar_b200.cabi`, ctypes drops
the GIL), so while species k is being reduced on the compute stream, species k+1 is
already streaming disk -> pinned lanes -> device on the reader's own I/O streams
(`rgc_tristan_read_particles` touches no compute scratch; see rgc_tristan.cpp).  Bins,
F table and the kernels' plans are built once and reused for every species and step
(plan caches in rgc_histogram.cu / rgc_sync_pair.cu).  With a communicator installed
(`ragnar_b200.dist.install_communicator`) every rank reads its own particle range of
each species and the results are the all-reduced sums.
"""
from __future__ import annotations

import threading
import time
from dataclasses import dataclass, field

import numpy as np

from . import cabi
from . import dist as rdist


@dataclass
class SpeciesResult:
    step: int
    label: str
    species: int
    nparticles: int
    distribution: np.ndarray  # float32 [len(gamma_bins)]
    spectrum: np.ndarray      # float32 [len(photon_bins)]
    spectrum64: np.ndarray
    spectrum_from_dist: np.ndarray | None = None  # SynchrotronSpectrumFromDist of `distribution`
    read_s: float = 0.0
    compute_s: float = 0.0


@dataclass
class PipelineReport:
    results: list = field(default_factory=list)
    wall_s: float = 0.0
    read_s: float = 0.0     # sum of the readers' wall times (overlapped when prefetching)
    compute_s: float = 0.0  # sum of histogram + spectrum wall times

    def by(self, step: int, label: str) -> SpeciesResult:
        for r in self.results:
            if r.step == step and r.label == label:
                return r
        raise KeyError((step, label))


class _Reader(threading.Thread):
    """Reads one species in the background; `.get()` joins and returns it."""

    def __init__(self, path, step, sp, dim, ignore_coords, rank, world):
        super().__init__(daemon=True)
        self.args = (path, step, sp, dim, ignore_coords, rank, world)
        self.out = None
        self.err = None
        self.seconds = 0.0

    def run(self):
        path, step, sp, dim, ignore_coords, rank, world = self.args
        t0 = time.perf_counter()
        try:
            self.out = read_species(path, step, sp, dim, ignore_coords, rank, world)
        except BaseException as e:  # re-raised in the caller's thread
            self.err = e
        self.seconds = time.perf_counter() - t0

    def get(self):
        if self.is_alive() or self.ident is not None:
            self.join()
        if self.err is not None:
            raise self.err
        return self.out


def read_species(path: str, step: int, sp: int, dim: int = 3, ignore_coords: bool = True,
                 rank: int = 0, world: int = 1):
    """This rank's particle range of one species (whole species for world == 1)."""
    if world == 1:
        p, _ = cabi.tristan_read_particles(path, step, sp, ignore_coords=ignore_coords, dim=dim)
        return p
    # size the species first (datasets are 1-D, all of one length: tristan-v2.cpp:119-130)
    fname = f"{path.rstrip('/')}/output/prtl/prtl.tot.{step:05d}"
    with cabi.H5File(fname, "r") as f:
        ntotal = int(f.info(f"u_{sp}")["dims"][0])
    off, cnt = rdist.shard_range(ntotal, rank, world)
    if cnt == 0:
        p = cabi.Particles(dim).allocate(16)
        p.n = 0
        return p
    p, _ = cabi.tristan_read_range(path, step, sp, off, cnt, ignore_coords=ignore_coords, dim=dim)
    return p


def process_steps(path: str, steps, species, photon_bins, gamma_bins, B0: float, g_syn: float,
                  e_syn_at_g_syn: float, *, dim: int = 3, ignore_coordinates: bool = True,
                  fourvel: bool = True, gamma_bins_log_spaced: bool = True,
                  out_file: str | None = None, prefetch: bool = False, rank: int = 0,
                  world: int = 1, from_dist: bool = True) -> PipelineReport:
    """steps: iterable of step numbers; species: [(label, sp), ...] as in
    `TristanV2.readParticles(label, sp)`.  Returns per (step, species) the energy
    distribution (Particles.energyDistribution) and the synchrotron spectrum
    (SynchrotronSpectrum_<D>D) and, with `out_file`, writes them with the legacy driver's
    dataset names (`<name>_<label>` gets a `_<step>` suffix when several steps are given).
    `from_dist=True` also evaluates SynchrotronSpectrumFromDist of every energy distribution:
    the (steps x species) distributions share their bins, so they go through ONE batched
    call (`rgc_sync_spectrum_dist_batch`: the kernel matrix is built once and contracted with
    the whole batch) -> dataset `sync_intensity_dist_<label>`.
    `prefetch=True` reads species k+1 in a background thread while species k is reduced;
    measured on B200 (profiles/r1_pipeline_3steps_1e8_v4.json) the reduction is 3.5 % of
    the page-cache read time (36 ms vs 1 s per 1e9 particles), so the overlap buys nothing
    there and is off by default — it is for many-bin spectra on fast storage."""
    photon_bins = np.ascontiguousarray(photon_bins, np.float32)
    gamma_bins = np.ascontiguousarray(gamma_bins, np.float32)
    steps = list(steps)
    work = [(st, label, sp) for st in steps for (label, sp) in species]
    table = cabi.tabulate_ffunc()
    report = PipelineReport()
    t_start = time.perf_counter()

    def start(i):
        st, _, sp = work[i]
        r = _Reader(path, st, sp, dim, ignore_coordinates, rank, world)
        if prefetch:
            r.start()
        else:
            r.run()
        return r

    nxt = start(0) if work else None
    for i, (st, label, sp) in enumerate(work):
        reader = nxt
        prtls = reader.get()
        nxt = start(i + 1) if i + 1 < len(work) else None  # overlaps the compute below
        t0 = time.perf_counter()
        # one call: histogram and spectrum kernels back to back on the stream, one wait
        hist, _, s32, s64 = cabi.hist_and_spectrum(prtls, gamma_bins, gamma_bins_log_spaced, fourvel,
                                                   photon_bins, B0, g_syn, e_syn_at_g_syn, table=table)
        dt = time.perf_counter() - t0
        report.results.append(SpeciesResult(st, label, sp, prtls.n, hist, s32, s64,
                                            reader.seconds, dt))
        report.read_s += reader.seconds
        report.compute_s += dt
        prtls.release()
    if from_dist and report.results:
        # all (step, species) distributions on the shared gamma-beta bins in one batched call
        fb = np.stack([r.distribution for r in report.results])
        d32, _ = cabi.sync_spectrum_dist_batch(gamma_bins, fb, gamma_bins_log_spaced, photon_bins, g_syn,
                                               e_syn_at_g_syn, table=table)
        for r, row in zip(report.results, d32):
            r.spectrum_from_dist = row
    report.wall_s = time.perf_counter() - t_start
    if out_file is not None and rank == 0:
        write_results(out_file, report, photon_bins, gamma_bins, multi_step=len(steps) > 1)
    return report


def write_results(out_file: str, report: PipelineReport, photon_bins, gamma_bins,
                  multi_step: bool = False) -> None:
    """legacy/simulation.cpp.bak:39,60-64,156-159 dataset names"""
    with cabi.H5File(out_file, "w") as f:
        def put(name, arr):
            arr = np.ascontiguousarray(arr, np.float32)
            f.create_dataset(name, np.float32, arr.size)
            f.write(name, arr)

        put("sync_photon_energy_mec2", photon_bins)
        for r in report.results:
            tag = f"{r.label}_{r.step}" if multi_step else r.label
            put(f"gammaM1_{tag}", gamma_bins)
            put(f"distribution_{tag}", r.distribution)
            put(f"sync_intensity_{tag}", r.spectrum)
            if r.spectrum_from_dist is not None:
                put(f"sync_intensity_dist_{tag}", r.spectrum_from_dist)
