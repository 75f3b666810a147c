#!/usr/bin/env python
"""bench.py — ragnar's radiation hot path on B200, the metric BASELINE.json names:
particle x photon-bin synchrotron evaluations per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[2], "SynchrotronSpectrum_3D: 1e8 synthetic
electrons in a random-angle uniform B field, 200 photon bins, 1 GPU"): per GPU,
1e8 electrons with U1 ~ u^-2 on [1,100], E = 0, B an isotropic unit vector;
photon bins Logbins(0.01, 1e5, 200, mec2); B0 = g_syn = e_syn_at_g_syn = 1.
One *step* = one pass of the hot path over that batch:
Particles.energyDistribution(Logbins(1e-2, 1e3, 200)) + SynchrotronSpectrum_3D, issued as the
batched driver issues them (one C-ABI call, rgc_hist_and_spectrum: the histogram's kernels are
enqueued ahead of the spectrum pipeline and both are collected after one wait; same kernels,
bit-identical results); the same step as two separate calls is reported beside it
(step_roofline.ms_per_step_as_two_separate_calls), and `e2e` goes through the module's two calls.
Evaluations per step = particles x photon bins (every pair the reference functor
runs for).  N > 1: one process per GPU (torchrun), each rank owns
1e8 particles of one global Philox stream (weak scaling); the per-rank spectrum
and histogram partials are summed by one NCCL all-reduce each inside the call.

`value` is timed with inputs resident in HBM; `e2e` is the same step through the pybind11
module `ragnar` from pageable NumPy columns (H2D of all nine columns and D2H of the results
inside the timed region), with the pinned C-ABI variant and the bare H2D ceiling beside it.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "synchrotron_particle_x_photon_bin_evals_per_sec"
UNIT = "evals/s"
N_PER_GPU = 100_000_000
NBINS = 200
PHOTON_BINS = (0.01, 1e5)
GAMMA_BINS = (1e-2, 1e3, 200)
CONSTS = (1.0, 1.0, 1.0)  # B0, g_syn, e_syn_at_g_syn
SEED = 123
CPU_SAMPLE = 10_000_000  # particles per CPU-baseline step (x 200 bins = 2e9 evals, ~10 s on 16 cores)
# dram__bytes_read.sum + dram__bytes_write.sum per launch at the default workload, from the
# committed `ncu --set full` capture profiles/r1_ncu_full_v7_summary.json (not measurable
# live: a number printed under a profiler is never a bench value)
NCU_TRAFFIC_BYTES = {
    "sync_pair_kernel": 0.800168e9 + 5.542e6,
    "sync_prologue_kernel": 3.600554e9 + 944.69e6,
    "sync_sort_kernel": 1.010174e9 + 763.34e6,
    "energy_hist_kernel": 1.200057e9 + 3.21e6,
}


def workload_config(n_per_gpu: int, nbins: int, ngpus: int, population: str = "config3") -> dict:
    if population != "config3":
        return {
            "workload": "SynchrotronSpectrum_3D + Particles.energyDistribution, BASELINE configs[4] "
                        "(at scale: full-3D population)",
            "particles_per_gpu": n_per_gpu, "photon_bins": nbins,
            "photon_bin_range_mec2": list(PHOTON_BINS), "gamma_beta_bins": list(GAMMA_BINS),
            "population": "isotropic U, |U|~u^-2 on [1,100], |B| in [0.5,2] isotropic, "
                          "E = 0.1 B x random (device Philox4x32-10, seed 123)",
            "sharding": f"particles/{ngpus} ranks, NCCL all-reduce of spectra" if ngpus > 1 else "none",
            "cache": "inputs_larger_than_L2",
        }
    return {
        "workload": "SynchrotronSpectrum_3D + Particles.energyDistribution, BASELINE configs[2]",
        "particles_per_gpu": n_per_gpu,
        "photon_bins": nbins,
        "photon_bin_range_mec2": list(PHOTON_BINS),
        "gamma_beta_bins": list(GAMMA_BINS),
        "population": "U1~u^-2 on [1,100], E=0, B isotropic unit (device Philox4x32-10, seed 123)",
        "sharding": f"particles/{ngpus} ranks, all-reduce of the per-rank spectra and histograms "
                    f"(peer-store exchange over NVLink, NCCL fallback)" if ngpus > 1 else "none",
        "cache": "inputs_larger_than_L2 (3.6 GB of particle columns per pass vs 126 MB L2)",
    }


# ------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.rows = []
        self._stop = threading.Event()
        self._thread = None

    def _run(self):
        # NVML (sub-millisecond per sample) when available: the timed region of the default
        # run is ~35 ms; nvidia-smi (tens of ms per call) otherwise
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.device)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40),
                    ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))
            while not self._stop.is_set():
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                r = int(get_reasons(h))
                power = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                self.rows.append([str(sm), str(mx), str(power)] +
                                 ["Active" if r & b else "Not Active" for _, b in bits])
                self._stop.wait(0.004)
            return
        except Exception as e:  # fall back to nvidia-smi
            print(f"bench.py: NVML clock sampling unavailable ({type(e).__name__}: {e})", file=sys.stderr)
        while not self._stop.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                     "-i", str(self.device)], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=6)

    def summary(self) -> dict:
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for name, flag in zip(names, r[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# -------------------------------------------------------------- reference arm
def _ref_module():
    import oracle

    if oracle.ref_available():
        return oracle.ref(), "reference"
    return None, "port"


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_step(sample_cols, bins, gbins):
    """One bounded CPU sample of the step through the reference's own code path
    (oracle/_ref = the reference sources on the Kokkos-subset OpenMP shim) or,
    where that was never built, the C++ port.  Returns (seconds, kind)."""
    import contextlib
    import io

    import oracle

    U, E, B = sample_cols
    mod, kind = _ref_module()
    t0 = time.perf_counter()
    if mod is not None:
        with contextlib.redirect_stdout(io.StringIO()):
            p = mod.Particles_3D("e-")
            p.fromArrays({f"{q}{d + 1}": a[d] for q, a in (("U", U), ("E", E), ("B", B))
                          for d in range(3)})
            t0 = time.perf_counter()  # the reference's fromArrays is not part of the metric
            p.energyDistribution(mod.Logbins(*GAMMA_BINS))
            mod.SynchrotronSpectrum_3D(p, mod.Logbins(*PHOTON_BINS, len(bins), "mec2"), *CONSTS)
    else:
        oracle.port.energy_distribution(*U, gbins, True, True)
        oracle.port.sync_spectrum_particles(U, E, B, bins, *CONSTS)
    return time.perf_counter() - t0, kind


REFERENCE_BUDGET_S = 150.0  # wall-clock budget of all (warmup + steps) reference passes


def run_reference(args) -> None:
    """The reference's own CPU implementation on all host threads.  torchrun exports
    OMP_NUM_THREADS=1: the width is set explicitly.  Every step is a bounded sample of the
    workload (the path is linear in the particle count), sized from a short probe so that
    warmup + steps passes fit REFERENCE_BUDGET_S."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from tests import synth

    import oracle

    cores = oracle.port.set_num_threads(host_threads())
    nbins = args.bins
    bins = oracle.port.logspace(*PHOTON_BINS, nbins)
    gbins = oracle.port.logspace(*GAMMA_BINS)
    probe_n = min(500_000, args.particles)
    probe = synth.config3(probe_n, seed=SEED)
    cpu_step(probe, bins, gbins)  # also pays the F-table build and the first-touch costs
    dt, kind = cpu_step(probe, bins, gbins)
    per_particle = dt / probe_n
    passes = max(1, args.steps + args.warmup)
    sample = int(REFERENCE_BUDGET_S / passes / per_particle)
    sample = max(100_000, min(sample, args.cpu_sample, args.particles))
    sample -= sample % 100_000 if sample > 100_000 else 0
    cols = synth.config3(sample, seed=SEED)
    for _ in range(args.warmup):
        _, kind = cpu_step(cols, bins, gbins)
    times = []
    for _ in range(args.steps):
        dt, kind = cpu_step(cols, bins, gbins)
        times.append(dt)
    total = sum(times)
    value = sample * nbins * args.steps / total
    sample_desc = (f"{sample} particles x {nbins} bins per step (first {sample} of the workload's "
                   f"population, host-generated; sized from a {probe_n}-particle probe so that "
                   f"{passes} passes fit {REFERENCE_BUDGET_S:.0f} s), {cores} OpenMP threads")
    cfg = workload_config(args.particles, nbins, args.gpus)
    cfg["timed_sample_particles_per_step"] = sample
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 terms (reference arithmetic)", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": sample_desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# -------------------------------------------------------------------- our arm
def _ncu_traffic() -> dict:
    """dram bytes per launch from the committed `ncu --set full` capture, valid only while the
    kernel's source file still hashes to what was profiled (profiles/ncu_traffic.json, written
    by tools/ncu_traffic.py next to the capture it summarises)."""
    import hashlib

    path = ROOT / "profiles" / "ncu_traffic.json"
    if not path.exists():
        return {}
    rec = json.loads(path.read_text())
    out = {}
    for kernel, ent in rec.get("kernels", {}).items():
        src = ROOT / ent["source"]
        if src.exists() and hashlib.sha256(src.read_bytes()).hexdigest()[:16] == ent["source_sha16"] \
                and ent.get("particles") == N_PER_GPU and ent.get("bins") == NBINS:
            out[kernel] = {"bytes": ent["dram_bytes"], "from": rec.get("capture")}
    return out


def run_ours(args) -> None:
    import torch

    from ragnar_b200 import cabi
    from ragnar_b200 import dist as rdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    all_cpus = os.sched_getaffinity(0)
    numa = rdist.bind_to_gpu_numa(local_rank)  # before any pinned allocation / worker thread
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cabi.init(local_rank)
    rdist.install_communicator(cabi, dist)
    stream = torch.cuda.ExternalStream(cabi.stream_handle(), device=local_rank)

    def barrier():
        cabi.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(a: np.ndarray) -> np.ndarray:
        if dist is None:
            return a
        t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
        dist.all_reduce(t)
        return t.cpu().numpy()

    table = cabi.tabulate_ffunc()
    gbins = cabi.logspace(*GAMMA_BINS)
    peaks = {}
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peaks = json.loads(peaks_path.read_text())
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"

    def timed_workload(prtls, n, bins, steps, warmup):
        """`steps` passes of the hot path over resident particles: CUDA events on the library's
        stream, barrier + synchronize on both sides, max over ranks."""
        km = {"hist": [], "spec": [], "prologue": [], "sort": [], "issued": [], "ontable": []}

        def step(record=True):
            # Particles.energyDistribution + SynchrotronSpectrum_3D of the batch in ONE call, as the
            # batched driver (ragnar_b200/pipeline.py) issues it: histogram and spectrum kernels
            # back to back on the stream, one wait (rgc_hist_and_spectrum)
            h32, h64, s32, s64 = cabi.hist_and_spectrum(prtls, gbins, True, True, bins, *CONSTS, table=table)
            if record:
                times = cabi.last_kernel_times()
                km["spec"].append(times[1])
                km["prologue"].append(times[2])
                km["sort"].append(times[3])
                km["issued"].append(cabi.last_pair_lane_evals())
                km["ontable"].append(cabi.last_pair_ontable_evals())
            return (h32, None, h64), (s32, s64)

        for _ in range(warmup):
            step(False)
        barrier()
        launches0 = cabi.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(stream)
        for _ in range(steps):
            hist, spec = step()
        ev1.record(stream)
        barrier()
        launches = cabi.launch_count() - launches0
        ms_per_step = max_over_ranks(ev0.elapsed_time(ev1)) / steps
        # beside it, untimed for the metric: the same step as the two separate calls of the
        # reference API (their sum is what a script calling energyDistribution and
        # SynchrotronSpectrum_3D one after the other sees), and the stand-alone histogram kernel
        barrier()
        ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev2.record(stream)
        for _ in range(3):
            cabi.energy_histogram(prtls, gbins, True, True, want_counts=False)
            km["hist"].append(cabi.last_kernel_ms()[1])
            cabi.sync_spectrum_particles(prtls, bins, *CONSTS, table=table)
        ev3.record(stream)
        barrier()
        km["separate_calls_ms"] = [max_over_ranks(ev2.elapsed_time(ev3)) / 3]
        return ms_per_step, launches, {k: statistics.mean(v) for k, v in km.items()}, hist, spec

    def rooflines(n, nbins, km, ffma_peak_tflops, traffic):
        issued, ontable = km["issued"], km["ontable"]
        spec_ms, pro_ms, sort_ms, hist_ms = km["spec"], km["prologue"], km["sort"], km["hist"]
        achieved = issued * 4 / (spec_ms * 1e-3) / 1e12

        def tr(kernel):
            return traffic.get(kernel, {}).get("bytes")
        pair = {
            "kernel": "sync_pair_kernel", "bound": "fp32",
            "achieved": achieved, "peak": ffma_peak_tflops, "unit": "TFLOP/s",
            "frac": achieved / ffma_peak_tflops, "traffic": tr("sync_pair_kernel"),
            "traffic_unit": "bytes per launch (ncu dram read + write of the committed capture named in "
                            "profiles/ncu_traffic.json; null when the kernel's source changed since, or "
                            "for a non-default workload); algorithmic: 8 B per particle",
            "peak_source": "FFMA issue rate measured on this device by rgc_measure_peak(0) in this "
                           "run (of measured); achieved = issued hinge evaluations x 4 flop "
                           "(FADD.SAT + FFMA) / kernel time.  The sub-bucket decomposition leaves "
                           "about one pair in eight for the pair loop (issued_over_total); the rest of "
                           "the kernel's time is the CTA-wide counting sort of every 4096-entry piece, "
                           "so at one lane group per sub-bucket (200 bins) the kernel is not FP32-bound",
            "evals_issued_per_launch": issued, "evals_on_table_per_launch": ontable,
            "evals_per_launch": n * nbins, "issued_over_total": issued / (n * nbins),
            "on_table_over_total": ontable / (n * nbins),
            "evals_per_s": n * nbins / (spec_ms * 1e-3), "ms_per_launch": spec_ms,
            "hbm_gbs_of_this_kernel": n * 8 / (spec_ms * 1e-3) / 1e9,
        }
        pro = {
            "kernel": "sync_prologue_kernel", "bound": "hbm",
            "achieved": n * 46 / (pro_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
            "frac": n * 46 / (pro_ms * 1e-3) / 1e9 / hbm_peak, "traffic": tr("sync_prologue_kernel"),
            "peak_source": hbm_src, "ms_per_launch": pro_ms,
            "bytes_per_particle": "36 read (U,E,B) + 10 written (fc, w, bucket)",
        }
        sort = {
            "kernel": "sync_sort_kernel (+ pair_colscan_kernel)", "bound": "hbm",
            "achieved": n * 18 / (sort_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
            "frac": n * 18 / (sort_ms * 1e-3) / 1e9 / hbm_peak, "traffic": tr("sync_sort_kernel"),
            "peak_source": hbm_src, "ms_per_launch": sort_ms,
            "bytes_per_particle": "10 read (fc, w, bucket) + 8 written (fc, w in global bucket order)",
        }
        hist = {
            "kernel": "energy_hist_kernel", "bound": "hbm",
            "achieved": n * 12 / (hist_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
            "frac": n * 12 / (hist_ms * 1e-3) / 1e9 / hbm_peak, "traffic": tr("energy_hist_kernel"),
            "peak_source": hbm_src, "particles_per_s": n / (hist_ms * 1e-3), "ms_per_launch": hist_ms,
        }
        return pair, pro, sort, hist

    def dominant(rs):
        """the kernel with the longest launch is the line's `roofline`; the others keep their keys"""
        return max(rs, key=lambda r: r["ms_per_launch"])

    def step_roofline(n, km, ms_per_step, ffma_peak_tflops):
        """whole step against its own floor: every particle column read once (36 B; the
        histogram's U is among them) and the on-table pairs at 4 flop each"""
        t_hbm = n * 36 / (hbm_peak * 1e9) * 1e3
        t_fp = km["issued"] * 4 / (ffma_peak_tflops * 1e12) * 1e3
        return {"ideal_ms": max(t_hbm, t_fp), "hbm_floor_ms": t_hbm, "fp32_floor_ms": t_fp,
                "ms_per_step": ms_per_step, "step_frac": max(t_hbm, t_fp) / ms_per_step,
                "ms_per_step_as_two_separate_calls": km["separate_calls_ms"],
                "moved_bytes_per_particle": 84,
                "definition": "max(36 B x particles / hbm_gbs, pair evaluations the decomposition needs x "
                              "4 flop / measured FFMA peak) / ms_per_step; the pipeline moves 84 B per "
                              "particle through HBM (12 histogram + 46 prologue + 18 sort + 8 pair)"}

    def parity_check(prtls, bins, sample, nthreads):
        """the code that was just timed against the oracle: this rank's first `sample`
        particles (hinge pipeline: sample > 2^19); with N ranks the library all-reduces the N
        slices, the oracle results are summed over ranks the same way; u64 counts must be
        identical (SURVEY 8d config-5 checks i-iii)"""
        import oracle

        oracle.port.set_num_threads(nthreads)
        cols = [[prtls.read(q, d, 0, sample) for d in range(3)] for q in (cabi.Q_U, cabi.Q_E, cabi.Q_B)]
        got = cabi.sync_spectrum_particles(prtls, bins, *CONSTS, table=table, nactive=sample)[1]
        _, counts, _ = cabi.energy_histogram(prtls, gbins, log_spaced=False, fourvel=True, nactive=sample)
        _, want = oracle.port.sync_spectrum_particles(*cols, bins, *CONSTS)
        _, _, want_c = oracle.port.energy_distribution(*cols[0], gbins, False, True)
        want = sum_over_ranks(want)
        want_c = sum_over_ranks(want_c.astype(np.int64))
        big = want >= 1e-6 * want.max()
        err = float(np.max(np.abs(got[big] - want[big]) / want[big]))
        return {"spectrum_rel_err": err, "spectrum_bar": 1e-5,
                "counts_equal": bool(np.array_equal(counts.astype(np.int64), want_c)),
                "counts_total": int(counts.sum()), "zero_bins_equal": bool(np.array_equal(got == 0, want == 0)),
                "sample": f"first {sample} particles of every rank x {len(bins)} bins, all-reduced over "
                          f"{world} rank(s); oracle = C++ port (pinned to the reference), fp64 sums",
                "ok": bool(err < 1e-5 and np.array_equal(counts.astype(np.int64), want_c))}

    # ================================================= the headline workload (configs[2])
    n, nbins = args.particles, args.bins
    bins = cabi.logspace(*PHOTON_BINS, nbins)
    prtls = cabi.Particles(3).allocate(n)
    # rank r owns global particle indices [r*n, (r+1)*n) of one Philox stream
    prtls.generate(0 if args.population == "config3" else 1, SEED, rank * n, 0, n, 1.0, 100.0)
    cabi.synchronize()
    # clocks / throttle reasons are sampled through all timed regions of this process
    clocks = ClockSampler(local_rank)
    clocks.__enter__()
    ms_per_step, launches, km, hist, spec = timed_workload(prtls, n, bins, args.steps, args.warmup)
    value = world * n * nbins / (ms_per_step * 1e-3)
    cpu_share = max(1, len(all_cpus) // world)
    parity = None
    if not args.no_parity:
        os.sched_setaffinity(0, all_cpus)
        parity = parity_check(prtls, bins, min(args.parity_sample, n), cpu_share)
        rdist.bind_to_gpu_numa(local_rank)  # back to the GPU's node for the pinned buffers below

    # ---- end to end: host buffers in, results out, every step
    e2e = None
    if not args.no_e2e:
        import ragnar_b200

        rg = ragnar_b200.load()
        rg.Initialize()
        names = [f"{q}{d + 1}" for q in "UEB" for d in range(3)]
        qd = [(q, d) for q in (cabi.Q_U, cabi.Q_E, cabi.Q_B) for d in range(3)]
        pageable = {nm: prtls.read(q, d, 0, n) for nm, (q, d) in zip(names, qd)}  # plain NumPy arrays
        pb = rg.Logbins(*PHOTON_BINS, nbins, rg.EnergyUnits.mec2)
        gb = rg.Logbins(*GAMMA_BINS)

        def module_step():
            """what a drop-in user runs: rg.Particles_3D.fromArrays(pageable ndarrays) ->
            energyDistribution -> SynchrotronSpectrum_3D -> as_array()"""
            p = rg.Particles_3D("e-")
            p.fromArrays(pageable)
            h = p.energyDistribution(gb).F().as_array()
            s = rg.SynchrotronSpectrum_3D(p, pb, *CONSTS).as_array()
            return h, s

        e2e_steps = max(1, min(args.steps, 5))
        module_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            h_mod, s_mod = module_step()
        barrier()
        dt_mod = max_over_ranks(time.perf_counter() - t0)
        assert np.array_equal(s_mod, spec[0]), "module e2e result differs from the resident run"

        # the same step through the C-ABI with pinned host columns (rgc_host_alloc)
        pinned = [cabi.PinnedArray(n) for _ in range(9)]
        for k, nm in enumerate(names):
            pinned[k].array[:] = pageable[nm]
        target = cabi.Particles(3).allocate(n)
        target.n = n

        def h2d_only():
            for k, (q, d) in enumerate(qd):
                target.write(q, d, 0, pinned[k].array)
            cabi.synchronize()

        def cabi_step():
            for k, (q, d) in enumerate(qd):
                target.write(q, d, 0, pinned[k].array)
            h32, h64, s32, s64 = cabi.hist_and_spectrum(target, gbins, True, True, bins, *CONSTS, table=table)
            return (h32, None, h64), (s32, s64)

        cabi_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            h_c, s_c = cabi_step()
        barrier()
        dt_cabi = max_over_ranks(time.perf_counter() - t0)
        assert np.array_equal(s_c[1], spec[1]), "C-ABI e2e result differs from the resident run"
        # the box's H2D ceiling for this transfer pattern: the nine pinned column copies of
        # every rank, concurrently, and nothing else
        h2d_only()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            h2d_only()
        barrier()
        dt_h2d = max_over_ranks(time.perf_counter() - t0)
        h2d_bytes = 9 * 4 * n + 4 * (nbins + len(gbins))
        evals = world * n * nbins * e2e_steps
        e2e = {
            "value": evals / dt_mod, "unit": UNIT,
            "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4 * (nbins + len(gbins)),
            "steps": e2e_steps, "ms_per_step": 1e3 * dt_mod / e2e_steps,
            "path": "pybind11 module `ragnar`: Particles_3D.fromArrays(dict of pageable float32 ndarrays) "
                    "-> energyDistribution -> SynchrotronSpectrum_3D -> as_array(); a new container "
                    "(allocation, zero fill, H2D of nine columns through the pinned staging ring) every step",
            "timer": "host wall clock around the calls, max over ranks",
            "pinned_cabi": {
                "value": evals / dt_cabi, "ms_per_step": 1e3 * dt_cabi / e2e_steps,
                "path": "C-ABI rgc_particles_write from rgc_host_alloc pinned columns into an existing "
                        "container + rgc_hist_and_spectrum"},
            "h2d_ceiling": {
                "ms_per_step": 1e3 * dt_h2d / e2e_steps,
                "GBps_per_rank": 9 * 4 * n * e2e_steps / dt_h2d / 1e9,
                "GBps_all_ranks": world * 9 * 4 * n * e2e_steps / dt_h2d / 1e9,
                "what": "the nine pinned column copies alone, all ranks concurrently (bare "
                        "cudaMemcpyAsync loop through the same entry point)"},
            "pinned_cabi_over_h2d_ceiling": dt_h2d / dt_cabi,
            "module_over_h2d_ceiling": dt_h2d / dt_mod,
            "numa": numa,
            "note": "the copy is ~20x the compute (65 ms vs 3 ms at 1e8 particles): overlapping them by "
                    "passes could gain at most 4 %, the link is the bound",
        }
        for p in pinned:
            p.close()
        target.release()
        del pageable

    # ---- FFMA peak of this device, this run (MEASURED_PEAKS.json has no FP32 entry)
    ffma_peak_tflops = max(cabi.measure_peak(cabi.PEAK_FFMA) for _ in range(2)) / 1e3
    pair_loop_peak = max(cabi.measure_peak(cabi.PEAK_PAIR) for _ in range(2)) * 1e9
    default_workload = n == N_PER_GPU and nbins == NBINS and args.population == "config3"
    traffic = _ncu_traffic() if default_workload else {}
    roofline_pair, roofline_pro, roofline_sort, roofline_hist = rooflines(n, nbins, km, ffma_peak_tflops, traffic)
    roofline_pair["bare_pair_loop_evals_per_s"] = pair_loop_peak
    roofline_pair["frac_of_bare_pair_loop"] = km["issued"] / (km["spec"] * 1e-3) / pair_loop_peak
    roofline = dominant([roofline_pair, roofline_pro, roofline_sort, roofline_hist])
    step_roof = step_roofline(n, km, ms_per_step, ffma_peak_tflops)

    # ================================================= BASELINE configs[4] share: every N
    other = {}
    if not args.no_config5:
        prtls.release()
        n5, nb5 = args.config5_particles, 1000
        bins5 = cabi.logspace(1e-3, 1e6, nb5)
        p5 = cabi.Particles(3).allocate(n5)
        p5.generate(1, SEED, rank * n5, 0, n5, 1.0, 100.0)
        cabi.synchronize()
        ms5, launches5, km5, _, _ = timed_workload(p5, n5, bins5, max(2, min(args.steps, 5)), 2)
        r5 = rooflines(n5, nb5, km5, ffma_peak_tflops, {})
        par5 = None
        if not args.no_parity:
            par5 = parity_check(p5, bins5, min(600_000, n5), cpu_share)
        other["config5"] = {
            "workload": f"BASELINE configs[4]: SynchrotronSpectrum_3D + energyDistribution, full-3D "
                        f"population (isotropic U, |B| in [0.5,2], E = 0.1 B x random), {n5} particles per "
                        f"GPU x {world} GPU(s) = {world * n5} particles, {nb5} photon bins "
                        f"Logbins(1e-3, 1e6), 4-5 lane groups per sub-bucket in the pair kernel",
            "value": world * n5 * nb5 / (ms5 * 1e-3), "unit": UNIT, "ms_per_step": ms5,
            "steps": max(2, min(args.steps, 5)), "gpu_launches": int(launches5),
            "roofline": dominant(list(r5)), "roofline_pair_kernel": r5[0],
            "roofline_prologue_kernel": r5[1], "roofline_sort_kernel": r5[2],
            "roofline_histogram_kernel": r5[3], "step_roofline": step_roofline(n5, km5, ms5, ffma_peak_tflops),
            "parity_in_bench": par5,
        }
        p5.release()
        prtls = None
    clocks.__exit__(None, None, None)

    # ---- the other single-GPU BASELINE configs, N = 1 only (extra keys, not the metric):
    # configs[1] histogram on the contended population (u in [5e-3, 2e3]: ~53 % of the
    # particles fall into clamp bin 0) and configs[0] FromDist latency
    if world == 1 and not args.no_cpu_baseline:
        p2 = cabi.Particles(3).allocate(n)
        p2.generate(2, SEED, 0, 0, n, 5e-3, 2e3)
        cabi.synchronize()
        for _ in range(3):
            cabi.energy_histogram(p2, gbins, True, True, want_counts=False)
        ms2 = []
        for _ in range(10):
            cabi.energy_histogram(p2, gbins, True, True, want_counts=False)
            ms2.append(cabi.last_kernel_ms()[1])
        h2 = statistics.mean(ms2)
        gb1 = cabi.logspace(1, 100, 200)
        fdist = cabi.generator_eval(0, [-2.0, 1.0, 100.0], gb1)
        b1 = cabi.logspace(0.01, 1e7, 200)
        cabi.sync_spectrum_dist(gb1, fdist, True, b1, 1.0, 1.0, table=table)
        t0 = time.perf_counter()
        for _ in range(20):
            cabi.sync_spectrum_dist(gb1, fdist, True, b1, 1.0, 1.0, table=table)
        fd_us = 1e6 * (time.perf_counter() - t0) / 20
        other["config1_histogram"] = {
            "workload": "Particles.energyDistribution, 1e8 power-law electrons u in [5e-3, 2e3] "
                        "into Logbins(1e-2, 1e3, 200)" if n == N_PER_GPU else f"{n} particles",
            "ms_per_launch": h2, "particles_per_s": n / (h2 * 1e-3),
            "hbm_GBps": n * 12 / (h2 * 1e-3) / 1e9, "frac_of_hbm_peak": n * 12 / (h2 * 1e-3) / 1e9 / hbm_peak}
        # the gather kernel that serves tables / bin sets the hinge pipeline cannot take
        # (DESIGN.md 3.3), forced onto the headline workload for a number to hold against it
        p2.generate(0 if args.population == "config3" else 1, SEED, 0, 0, n, 1.0, 100.0)
        cabi.synchronize()
        os.environ["RGC_SPECTRUM_PATH"] = "gather"
        try:
            cabi.sync_spectrum_particles(p2, bins, *CONSTS, table=table)
            gms = []
            for _ in range(3):
                cabi.sync_spectrum_particles(p2, bins, *CONSTS, table=table)
                gms.append(cabi.last_kernel_times()[1])
        finally:
            os.environ.pop("RGC_SPECTRUM_PATH", None)
        g_ms = statistics.mean(gms)
        other["gather_fallback"] = {
            "workload": f"the headline workload ({n} particles x {nbins} bins) forced onto sync_spectrum_kernel "
                        "(RGC_SPECTRUM_PATH=gather)",
            "ms_per_launch": g_ms, "evals_per_s": n * nbins / (g_ms * 1e-3),
            "lds_gather_GBps": n * nbins * 8 / (g_ms * 1e-3) / 1e9,
            "note": "8 B shared-memory gather + 7 issue slots per 32 evaluations; every pair is evaluated"}
        other["config0_fromdist"] = {
            "workload": "SynchrotronSpectrumFromDist: PlawGenerator(-2,1,100) on Logbins(1,100,200) -> "
                        "200 photon Logbins(0.01,1e7)", "evals": 40000,
            "us_per_call_host_wall": fd_us,
            "note": "latency-bound; literal kernel (the reference's float term per pair); F table cached"}

    # ---- CPU baseline beside it (rank 0, N = 1 only): bounded sample of the same workload
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        import oracle

        os.sched_setaffinity(0, all_cpus)
        cores = oracle.port.set_num_threads(len(all_cpus))
        sample = min(args.cpu_sample, n)
        src = p2 if prtls is None else prtls
        # (the contended-histogram population p2 shares nothing with the workload: regenerate)
        src.generate(0 if args.population == "config3" else 1, SEED, 0, 0, sample, 1.0, 100.0)
        cabi.synchronize()
        cols = [[src.read(q, d, 0, sample) for d in range(3)]
                for q in (cabi.Q_U, cabi.Q_E, cabi.Q_B)]
        obins = oracle.port.logspace(*PHOTON_BINS, nbins)
        ogbins = oracle.port.logspace(*GAMMA_BINS)
        dt, kind = cpu_step(cols, obins, ogbins)
        cpu_baseline = {
            "value": sample * nbins / dt, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"first {sample} particles of the workload x {nbins} bins, one pass "
                      f"({dt:.1f} s)",
        }
        p2.release()

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 pair terms, fp64 prologue and accumulation",
        "data": "synthetic", "config": workload_config(n, nbins, world, args.population),
        "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "roofline_pair_kernel": roofline_pair,
        "roofline_prologue_kernel": roofline_pro,
        "roofline_sort_kernel": roofline_sort,
        "roofline_histogram_kernel": roofline_hist,
        "step_roofline": step_roof,
        "parity_in_bench": parity,
        "cpu_baseline": cpu_baseline,
        "other_configs": other or None,
        "notes": "evals = particles x photon bins, every pair the reference's functor is launched for "
                 "(SURVEY.md 8d). The hinge of a (particle, bin) pair is exactly 0 or exactly linear "
                 "in the particle's table fraction unless the bin's threshold lies in the particle's own "
                 "eighth of the table cell, so the pair kernel evaluates roofline_pair_kernel."
                 "issued_over_total of the pairs one by one and takes the rest from sub-bucket moments "
                 "(rgc_sync_pair.cu); `roofline` is the kernel with the longest launch.",
    }
    emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def _claim_stdout() -> None:
    """Everything any library prints to fd 1 (NCCL's version banner, the reference's
    py::print) goes to stderr; only emit() writes to the real stdout: ONE JSON line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)


def main() -> None:
    global PHOTON_BINS
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--particles", type=int, default=N_PER_GPU, help="particles per GPU")
    ap.add_argument("--bins", type=int, default=NBINS)
    ap.add_argument("--bins-lo", type=float, default=PHOTON_BINS[0])
    ap.add_argument("--bins-hi", type=float, default=PHOTON_BINS[1])
    ap.add_argument("--population", choices=["config3", "full3d"], default="config3",
                    help="config3: U1~u^-2, E=0, B unit isotropic; full3d: BASELINE configs[4]")
    ap.add_argument("--cpu-sample", type=int, default=CPU_SAMPLE)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-bench oracle check")
    ap.add_argument("--parity-sample", type=int, default=1_000_000)
    ap.add_argument("--no-config5", action="store_true", help="skip the BASELINE configs[4] share")
    ap.add_argument("--config5-particles", type=int, default=500_000_000, help="per GPU")
    args = ap.parse_args()
    _claim_stdout()
    PHOTON_BINS = (args.bins_lo, args.bins_hi)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
