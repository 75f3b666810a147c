/*
 * ragnar_cuda.h — C-ABI of libragnar_cuda.so, the B200 (sm_100a) implementation
 * of ragnar's radiation hot path.
 *
 * This is the drop-in boundary: every entry point is what the reference's own
 * C++ layer (haykh/ragnar @ fceb6b08) would bind in place of the Kokkos code it
 * replaces; the reference interface each one stands in for is cited as
 * path:line relative to the reference root.  Plain C: opaque handles, raw
 * pointers and sizes; no CUDA, torch or C++ types.
 *
 * Conventions
 *   - every function returns 0 on success, a non-zero RGC_ERR_* code otherwise;
 *     rgc_last_error() then holds a thread-local message (CUDA / NCCL / HDF5
 *     error text included).  The reference throws C++ exceptions at the same
 *     points; the host layer (ragnar_b200/csrc/host) re-throws them.
 *   - "host" pointers may be pageable or pinned (rgc_host_alloc); pinned ones are
 *     copied with one async DMA, pageable ones are staged through a pinned ring.
 *   - calls are synchronous at return (the reference calls Kokkos::fence()
 *     before returning to Python: src/physics/synchrotron.cpp:102,142,
 *     src/containers/particles.cpp:256) unless documented otherwise.
 *   - there is NO CPU fallback: without a CUDA device rgc_init() fails and every
 *     compute entry point returns RGC_ERR_NOT_INITIALIZED.
 */
#ifndef RAGNAR_CUDA_H
#define RAGNAR_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* every entry point below is exported; the rest of the library is hidden */
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define RGC_OK                  0
#define RGC_ERR_INVALID         1 /* bad argument (-> std::runtime_error / range_error upstream) */
#define RGC_ERR_NOT_INITIALIZED 2 /* rgc_init() not called or no CUDA device */
#define RGC_ERR_CUDA            3 /* a CUDA runtime call failed */
#define RGC_ERR_NCCL            4 /* an NCCL call failed / libnccl not loadable */
#define RGC_ERR_IO              5 /* file / HDF5 format error */
#define RGC_ERR_OOM             6 /* device or pinned-host allocation failed */

/* element types of rgc_buf_t — Array1D<int|float|double>, src/containers/array.hpp:15-42 */
#define RGC_I32 0
#define RGC_F32 1
#define RGC_F64 2

/* particle quantities — Particles<D>::X,U,E,B, src/containers/particles.hpp:19-20 */
#define RGC_Q_X 0
#define RGC_Q_U 1
#define RGC_Q_E 2
#define RGC_Q_B 3

typedef struct rgc_buf       rgc_buf_t;       /* 1-D device array            */
typedef struct rgc_particles rgc_particles_t; /* SoA particle container      */

/* ------------------------------------------------------------------ runtime */

/* Kokkos::initialize() — src/pyinterface.cpp:30-39.  device < 0 selects
 * $LOCAL_RANK (one process per GPU under torchrun) or 0.  Idempotent. */
int rgc_init(int device);
/* Kokkos::finalize() — src/pyinterface.cpp:40-49 */
int rgc_finalize(void);
int rgc_is_initialized(void);
int rgc_device_count(int* count);
/* device ordinal, SM count and total HBM bytes of the active device */
int rgc_device_info(int* device, int* sm_count, size_t* hbm_bytes);
/* the library's compute stream (cudaStream_t) for event timing by harnesses */
int         rgc_stream(void** stream);
int         rgc_synchronize(void);
const char* rgc_last_error(void);
/* number of kernels this library has launched since rgc_init (bench evidence) */
uint64_t rgc_launch_count(void);

/* Particle columns are allocated from the library's own stream-ordered memory pool and stay cached
 * there when a container is released (re-creating a container per species / step then costs
 * no cudaMalloc / cudaFree); this returns the cached blocks to the device. */
int rgc_trim_memory(void);

/* pinned host memory for zero-staging H2D/D2H (cudaHostAlloc / cudaFreeHost) */
int rgc_host_alloc(size_t bytes, void** ptr);
int rgc_host_free(void* ptr);

/* ------------------------------------------------- multi-GPU (one rank/GPU) */

/* How result vectors are combined across ranks: 0 = single rank, 1 = ncclAllReduce,
 * 2 = peer-store exchange — every rank's 1 MiB exchange buffer is mapped into every other
 * rank (CUDA IPC, set up inside rgc_comm_init; handles travel through one ncclAllGather)
 * and one kernel stores the rank's vector into all peers over NVLink, raises a flag and
 * sums the slots in rank order.  Used for vectors of <= 8192 elements when all ranks share
 * a node with peer access; RGC_XCHG=0 forces NCCL.  The wait for the peers' flags is bounded in
 * time (RGC_XCHG_TIMEOUT_MS, default 600000): when a peer never delivers, the call that issued
 * the exchange returns RGC_ERR_NCCL (its result is poisoned, never a partial sum) and the
 * exchange stays unusable until the communicator is destroyed and re-created. */
int rgc_comm_exchange_kind(int* kind);

/* One process per GPU.  Rank 0 obtains an id, the launcher broadcasts its
 * RGC_COMM_ID_BYTES to all ranks out of band (torch.distributed / MPI / file),
 * every rank calls rgc_comm_init.  While a communicator is installed,
 * rgc_energy_histogram and rgc_sync_spectrum_particles all-reduce (sum) their
 * fp64 / u64 partials with ONE ncclAllReduce each, enqueued on the compute
 * stream right behind the reduction kernel.  The reference has no multi-GPU
 * path at all. */
#define RGC_COMM_ID_BYTES 128
int rgc_comm_get_unique_id(unsigned char id[RGC_COMM_ID_BYTES]);
int rgc_comm_init(const unsigned char id[RGC_COMM_ID_BYTES], int rank, int nranks);
int rgc_comm_destroy(void);
int rgc_comm_info(int* rank, int* nranks); /* 0,1 when no communicator */

/* -------------------------------------------- Array1D<T> device buffers */

/* Kokkos::View<T*>{label, n} (zero-initialised) */
int rgc_buf_create(int dtype, size_t n, rgc_buf_t** out);
/* Array1D(const py::array_t<T>&) — src/containers/array.hpp:20-30 */
int rgc_buf_from_host(int dtype, const void* host, size_t n, rgc_buf_t** out);
/* Array1D::as_array / as_vector / head — src/containers/array.cpp:18-60 */
int    rgc_buf_to_host(const rgc_buf_t* buf, size_t start, size_t n, void* host);
size_t rgc_buf_size(const rgc_buf_t* buf);
int    rgc_buf_dtype(const rgc_buf_t* buf);
/* raw device pointer (for harnesses that time kernels / share memory with torch) */
void* rgc_buf_device_ptr(const rgc_buf_t* buf);
/* Kokkos::View handles are ref-counted; so are these */
int rgc_buf_retain(rgc_buf_t* buf);
int rgc_buf_release(rgc_buf_t* buf);

/* ------------------------------------------------------------- Particles<D> */

/* Particles<D>(label) — src/containers/particles.hpp:22 (label stays host-side) */
int rgc_particles_create(int dim, rgc_particles_t** out);
int rgc_particles_release(rgc_particles_t* p);
/* Particles<D>::allocate — src/containers/particles.cpp:115-130: zero-filled SoA
 * columns U1..3,E1..3,B1..3 (+X1..XD when with_coords) of nalloc floats each */
int rgc_particles_allocate(rgc_particles_t* p, size_t nalloc, int with_coords);
/* Particles<D>::reallocate — src/containers/particles.cpp:132-187: grow, keep data */
int rgc_particles_reallocate(rgc_particles_t* p, size_t nalloc);
/* add (zeroed) coordinate columns to a container allocated without them */
int    rgc_particles_enable_coords(rgc_particles_t* p);
int    rgc_particles_has_coords(const rgc_particles_t* p);
size_t rgc_particles_nalloc(const rgc_particles_t* p);
int    rgc_particles_dim(const rgc_particles_t* p);
/* one column of fromArrays / readPrtlQuantity: host[0..n) -> column (quantity,
 * comp) at [start, start+n).  src/containers/particles.cpp:56-98,
 * src/plugins/tristan-v2.cpp:76-93.  Asynchronous: returns once the source
 * buffer may be reused; rgc_synchronize() (or any compute call) orders it. */
int rgc_particles_write(rgc_particles_t* p, int quantity, int comp, size_t start,
                        const float* host, size_t n);
/* getSubview — src/containers/particles.cpp:346-383: column -> host / new buffer */
int rgc_particles_read(const rgc_particles_t* p, int quantity, int comp, size_t start,
                       size_t n, float* host);
int rgc_particles_column(const rgc_particles_t* p, int quantity, int comp, size_t n,
                         rgc_buf_t** out);
/* raw device pointer of a column (harness use) */
void* rgc_particles_device_ptr(const rgc_particles_t* p, int quantity, int comp);

/* Synthetic populations generated ON DEVICE (counter-based Philox4x32-10 keyed by
 * (seed, global particle index), so any shard of any size reproduces the same
 * particles).  Used by bench.py / tests for N too large to stage through the
 * host.  Fills U,E,B of [start, start+n) for global indices
 * [global_offset, global_offset+n).
 *   kind 0  "config 3": |U| power law p=-2 on [1,100] along x, E=0, B isotropic unit
 *   kind 1  "full 3-D": isotropic U, |U| power law p=-2 on [umin,umax],
 *                        |B| in [0.5,2] isotropic, E = 0.1 * (B x random unit)
 *   kind 2  "config 2": isotropic U, |U| power law p=-2 on [umin,umax], E=B=0 */
int rgc_particles_generate(rgc_particles_t* p, int kind, uint64_t seed,
                           uint64_t global_offset, size_t start, size_t n, float umin,
                           float umax);

/* -------------------------------------------------- host-exact small pieces */

/* Linspace / Logspace — src/utils/snippets.cpp:21-62 (bit-identical bin edges
 * are a precondition of bit-exact histograms; evaluated on the host with the
 * reference's exact float/double promotions) */
int rgc_linspace(float start, float stop, size_t num, float* out);
int rgc_logspace(float start, float stop, size_t num, float* out);
/* The same grids built ON THE DEVICE into a new buffer (grids of >= 1e6 points),
 * bit-identical to rgc_linspace / rgc_logspace: Linspace is IEEE float arithmetic; Logspace
 * evaluates pow(10, exponent) on the device and re-evaluates on the host, with the reference's
 * libm, the few elements (~1e-8 of them) whose double result lies within 16 ulp of a float
 * rounding boundary.  Same errors as the host versions. */
int rgc_linspace_device(float start, float stop, size_t num, rgc_buf_t** out);
int rgc_logspace_device(float start, float stop, size_t num, rgc_buf_t** out);
/* TabulatedFunction<LG>::findMinMax — src/containers/tabulation.cpp:84-102 (Kokkos MinMax
 * reduce: `<` / `>` comparisons, NaNs never win) on a device-resident float buffer */
int rgc_buf_minmax(const rgc_buf_t* buf, float* min_out, float* max_out);
/* InterpolateTabulatedFunction<LG> — src/containers/tabulation.hpp:19-53 — of a device-resident
 * table (tab_x, tab_y) at every element of x0, the reference's float sequence bit for bit
 * (glibc's log10f restated on the device); verify() errors of tabulation.cpp:104-117 are
 * returned as RGC_ERR_INVALID.  *out is a new float buffer of x0's length. */
int rgc_tabulated_eval(int loggrid, const rgc_buf_t* tab_x, const rgc_buf_t* tab_y, float yfill,
                       const rgc_buf_t* x0, rgc_buf_t** out);
/* sync::Ffunc_integrand — src/physics/synchrotron.cpp:28-46 */
int rgc_sync_ffunc_integrand(float x, float* out);
/* sync::TabulateFfunc — src/physics/synchrotron.cpp:48-65 (cached per (n,xmin,xmax)) */
int rgc_sync_tabulate_ffunc(size_t npoints, float xmin, float xmax, float* xs, float* ys);
/* InterpolateTabulatedFunction<LG> — src/containers/tabulation.hpp:19-53 (host scalar) */
int rgc_interpolate(int loggrid, float x0, const float* x, const float* y, size_t n,
                    float yfill, float* out);
/* generators — src/containers/distributions.cpp:37-130.  kind 0 Plaw(p,emin,emax),
 * 1 BrokenPlaw(e_break,p1,p2,emin,emax), 2 Delta(energy0,denergy); params in that order */
int rgc_generator_eval(int kind, const float* params, const float* energy, size_t n,
                       float* out);

/* ------------------------------------------------------------- the hot path */

/* Particles<D>::energyDistribution — src/containers/particles.cpp:189-260.
 * bins: n host floats (left edges; index formula is logarithmic whatever
 * log_spaced says, exactly as the reference).  Over particles [0, nactive):
 *   out_hist[n]    float  sum of 1/energy (log_spaced) or particle count (!log_spaced),
 *                         accumulated wide and rounded once
 *   out_counts[n]  u64    particle counts per bin (bit-exact vs the reference's index)
 *   out_sum64[n]   double the wide sum before rounding           (each may be NULL)
 * With a communicator installed the sums are all-reduced over ranks. */
int rgc_energy_histogram(const rgc_particles_t* p, size_t nactive, const float* bins,
                         size_t n, int log_spaced, int fourvel, float* out_hist,
                         uint64_t* out_counts, double* out_sum64);

/* SynchrotronSpectrum<D> + sync::Kernel<D> — src/physics/synchrotron.cpp:107-145,
 * src/physics/synchrotron.hpp:103-233.  bins_e_syn: nbins host floats (units are
 * checked by the caller, synchrotron.hpp:139-142); (tab_x, tab_y): the F(x) table
 * of sync::TabulateFfunc (log grid).  out_spec[nbins] float, out_spec64 (optional)
 * the fp64 sums it was rounded from.  All-reduced over ranks when a communicator
 * is installed; every rank must make the call (p may be NULL / unallocated with
 * nactive = 0: an empty shard that still joins the all-reduce).  Populations of up to
 * RGC_LITERAL_MAX_N particles (default 2^19) are evaluated with the reference's own
 * float arithmetic per pair (rgc_sync_literal.cu), larger ones by the hinge pipeline. */
int rgc_sync_spectrum_particles(const rgc_particles_t* p, size_t nactive,
                                const float* bins_e_syn, size_t nbins, const float* tab_x,
                                const float* tab_y, size_t tab_n, float B0, float g_syn,
                                float e_syn_at_g_syn, float* out_spec, double* out_spec64);
/* Particles.energyDistribution + SynchrotronSpectrum_<D>D of the same particles in one call — what
 * the batched driver (legacy/simulation.cpp.bak:67-219, ragnar_b200/pipeline.py) does per species.
 * The histogram's kernels and its all-reduce are enqueued without a wait, the spectrum pipeline
 * follows on the same stream and both results are collected after one synchronisation: the
 * same kernels as the two separate calls, bit-identical results, one host round trip less.
 * out_hist / out_hist64: as rgc_energy_histogram (no counts output). */
int rgc_hist_and_spectrum(const rgc_particles_t* p, size_t nactive, const float* gbins, size_t ng,
                          int log_spaced, int fourvel, float* out_hist, double* out_hist64,
                          const float* bins_e_syn, size_t nbins, const float* tab_x,
                          const float* tab_y, size_t tab_n, float B0, float g_syn,
                          float e_syn_at_g_syn, float* out_spec, double* out_spec64);

/* SynchrotronSpectrumFromDist + sync::KernelFromDist —
 * src/physics/synchrotron.cpp:69-105, src/physics/synchrotron.hpp:41-96.
 * (gbeta, f): the TabulatedDistribution (ndist host floats each).  Replicated,
 * never all-reduced. */
int rgc_sync_spectrum_dist(const float* gbeta, const float* f, size_t ndist,
                           int islog_bins_prtls, const float* bins_e_syn, size_t nbins,
                           const float* tab_x, const float* tab_y, size_t tab_n, float g_syn,
                           float e_syn_at_g_syn, float* out_spec, double* out_spec64);
/* A batch of distributions on the SAME bins (steps x species of a run: the caller of the path,
 * legacy/simulation.cpp.bak:67-219): f holds nbatch rows of ndist values, out_spec /
 * out_spec64 nbatch rows of nbins.  mode 0: every (distribution, bin, photon bin) term with
 * the reference's float arithmetic, as rgc_sync_spectrum_dist; mode 1: the kernel matrix
 * K[g][j] = e_syn[j] * gbeta[g] * F(e_syn[j] / e_peak[g]) is built once and the batch is
 * contracted in fp64, out = f . K (each term within three float roundings, ~1e-7, of the
 * reference's); mode -1: the contraction from 64 distributions on (where it is the faster
 * one, profiles/r2_fromdist_batch.txt). */
int rgc_sync_spectrum_dist_batch(const float* gbeta, const float* f, size_t nbatch, size_t ndist,
                                 int islog_bins_prtls, const float* bins_e_syn, size_t nbins,
                                 const float* tab_x, const float* tab_y, size_t tab_n,
                                 float g_syn, float e_syn_at_g_syn, int mode, float* out_spec,
                                 double* out_spec64);

/* ICSpectrum + ic::Kernel + ic::KNfunc — src/physics/ic.cpp:15-46,
 * src/physics/ic.hpp:20-85.  (g_prtls, f_prtls): particle TabulatedDistribution;
 * (e_soft, f_soft): soft-photon TabulatedDistribution (mec^2); bins_e_ic: nic IC
 * energies.  out[j] = sum over (g, s) of the reference's float term, summed in fp64
 * (out_spec64) and rounded once to float (out_spec); either may be NULL.  The unit
 * checks of ic::Kernel's constructor (ic.hpp:48-55) live in the host wrapper, which
 * owns the Bins' units.  Index intent, not the reference's swapped MDRange extents
 * (see rgc_ic.cu); identical when nsoft == nic.  Replicated, never all-reduced. */
int rgc_ic_spectrum(const float* g_prtls, const float* f_prtls, size_t nprtls,
                    int islog_bins_prtls, const float* e_soft, const float* f_soft,
                    size_t nsoft, const float* bins_e_ic, size_t nic, float* out_spec,
                    double* out_spec64);

/* Device time (ms, CUDA events on the compute stream) of the kernels of the last
 * hot-path call on this thread: [0] total, [1] dominant kernel only. */
int rgc_last_kernel_ms(float ms[2]);
/* same, up to n = 4 entries: [2] the per-particle prologue kernel of the spectrum
 * (sync_prologue_kernel), [3] its bucket-sort kernels (column scan + sync_sort_kernel) */
int rgc_last_kernel_times(float* ms, int n);

/* Hinge evaluations (32 per lane group and sorted entry) the pair kernel actually issued
 * in the last rgc_sync_spectrum_particles call on this rank.  A particle is streamed only
 * through the lane groups of the bins whose hinge threshold lies in the particle's own
 * sub-bucket of the table cell (about one pair in eight; the rest enters through sub-bucket
 * moments), and lane groups whose bins are all beyond the table's zero tail for a bucket are
 * skipped (roofline accounting). */
int rgc_last_pair_lane_evals(double* lane_evals);
/* (particle, bin) pairs of the last rgc_sync_spectrum_particles call on this rank whose table
 * cell pair is not identically zero — what an ideal kernel would have to evaluate (the
 * whole-step roofline of bench.py); counted on the device by pair_moments_kernel. */
int rgc_last_pair_ontable_evals(double* evals);
/* Host-only description of the hinge path's plan for photon bins (e_syn, any order) on an F
 * table — no device needed; what the CPU tests check.  info[0] 1 if the hinge path takes these
 * bins, [1] lane groups, [2] slots (32 per group), [3] buckets, [4] chunks of <= 8 groups,
 * [5] buckets in which some sub-bucket's groups also see a neighbouring run, [6] most groups of
 * one sub-bucket, [7] sub-buckets per table cell; *phase = the phase phi of the sub-bucket
 * boundaries, s = floor(8 fc + phi); slot_bin[i] = index of the bin that slot i carries, -1 for
 * moment lanes and spare lanes (at most cap entries are written).  Replaces nothing in the
 * reference (its MDRange visits every pair, src/physics/synchrotron.cpp:124-139). */
int rgc_pair_plan_describe(const float* bins_e_syn, size_t nbins, const float* tab_x,
                           const float* tab_y, size_t tab_n, int info[8], float* phase,
                           int* slot_bin, size_t cap);
/* How the bucket sort of the hinge pipeline ranks particles: 1 = one shared-memory atomic per
 * particle, after a probe kernel verified ON THIS DEVICE that the lanes of one instruction
 * hitting one address are served in lane order (what makes results bitwise reproducible);
 * 0 = the probe failed, ranking by ballots (order fixed by construction); -1 = no hinge call
 * yet.  RGC_SORT_RANK=ballot forces 0, RGC_SORT_RANK=atomic skips the probe. */
int rgc_sort_rank_mode(int* mode);

/* Roofline denominators measured on the device, on the compute stream (bench
 * harness only; MEASURED_PEAKS.json has no FP32 / shared-memory entry).
 * kind 0: FFMA GFLOP/s; 1: G evaluations/s of the pair loop's FADD.SAT+FFMA mix;
 * 2: conflict-free LDS.64 gather GB/s; 3: HBM streaming-read GB/s;
 * 4: integer IMAD/SHF/LOP3 mix, G lane-instructions/s.  *sm_clock_mhz receives the
 * device's nominal SM clock. */
int rgc_measure_peak(int kind, double* value, double* sm_clock_mhz);

/* ------------------------------------------------------------ HDF5 arrays */

/* A self-contained HDF5 layer stands in for HighFive v2.10.1 + libhdf5 (the
 * reference's un-vendored dependency, cmake/dependencies.cmake:15-18).  Reads files
 * written by libhdf5 (superblock v0-v3, old- and new-style compact groups, IEEE
 * float / integer types of either byte order, contiguous / compact / chunked
 * layouts with deflate, shuffle and fletcher32); writes superblock-v0 files with
 * contiguous datasets, as libhdf5 does with default property lists.  The
 * rgc_h5_open..rgc_h5_write group is pure host code and works without a GPU. */
typedef struct rgc_h5 rgc_h5_t;
#define RGC_H5_READONLY  0 /* HighFive::File::ReadOnly */
#define RGC_H5_READWRITE 1 /* ReadWrite | Create: append datasets, create if missing */
#define RGC_H5_TRUNCATE  2 /* start a new file */
#define RGC_H5_CLASS_INTEGER 0
#define RGC_H5_CLASS_FLOAT   1
#define RGC_H5_LAYOUT_COMPACT    0
#define RGC_H5_LAYOUT_CONTIGUOUS 1
#define RGC_H5_LAYOUT_CHUNKED    2
/* HighFive::File{filename, mode} — src/plugins/tristan-v2.cpp:117, src/io/h5.cpp:21,54 */
int rgc_h5_open(const char* filename, int mode, rgc_h5_t** out);
int rgc_h5_close(rgc_h5_t* f);
/* newline-separated link names of a group ("/" or NULL = root), NUL-terminated;
 * *needed receives the buffer size the full list takes */
int rgc_h5_list(rgc_h5_t* f, const char* group, char* names, size_t cap, size_t* needed);
/* file.getDataSet(name).getDimensions() — tristan-v2.cpp:58-59,125; h5.cpp:22-23 */
int rgc_h5_dataset_info(rgc_h5_t* f, const char* name, int* rank, uint64_t* dims,
                        int max_rank, int* type_class, int* elem_size, int* layout);
/* dataset.select({start},{count},{stride}).read<T>(host) — tristan-v2.cpp:72,
 * h5.cpp:40; dtype = RGC_I32/F32/F64 of the destination (source types convert) */
int rgc_h5_read(rgc_h5_t* f, const char* name, size_t start, size_t count, size_t stride,
                int dtype, void* host);
/* file.createDataSet<T>(name, DataSpace({n})) — h5.cpp:59-60 (contiguous, zero-filled) */
int rgc_h5_create_dataset(rgc_h5_t* f, const char* name, int dtype, size_t n);
/* write_raw of elements [start, start+count) — h5.cpp:61 */
int rgc_h5_write(rgc_h5_t* f, const char* name, size_t start, size_t count, int dtype,
                 const void* host);
/* io::h5::Read1DArray<T> — src/io/h5.cpp:16-48: selection {0},{size or extent},{stride}
 * streamed disk -> pinned lanes -> device.  Same checks and messages as the reference. */
int rgc_h5_read_array(const char* filename, const char* dsetname, int dtype, size_t size,
                      size_t stride, rgc_buf_t** out);
/* io::h5::Write1DArray<T> — src/io/h5.cpp:50-68: opens ReadWrite|Create, creates the
 * dataset (fails if it exists) and writes the array */
int rgc_h5_write_array(const char* filename, const char* dsetname, const rgc_buf_t* array);

/* ------------------------------------------------------ Tristan-v2 plugin */

/* TristanV2<D>::readParticles — src/plugins/tristan-v2.cpp:95-188.  Opens
 * <path>/output/prtl/prtl.tot.<step:05d> (HDF5), reads x_/y_/z_ (first dim, unless
 * ignore_coords), u_/v_/w_, ex_/ey_/ez_, bx_/by_/bz_<sp> as float32 hyperslabs
 * [start, start+count*stride) and streams them disk -> pinned lanes -> device
 * (all columns concurrently, several I/O threads; $RGC_IO_THREADS overrides).
 * *ntotal receives the dataset length, *nread the particles stored.  Validation
 * (stride == 0, stride != 1 with size != 0, start + size >= ntotal) follows
 * tristan-v2.cpp:102-107,126-128 and returns RGC_ERR_INVALID with the reference's
 * message. */
int rgc_tristan_read_particles(const char* path, size_t step, unsigned sp, size_t start,
                               size_t size, size_t stride, int ignore_coords, int dim,
                               rgc_particles_t** out, size_t* ntotal, size_t* nread);
/* Sharded read of the multi-GPU path (not in the reference): exactly particles
 * [start, start + count) of species sp, count > 0, start + count <= ntotal — the
 * reference's selection rejects a range that ends at the last particle
 * (tristan-v2.cpp:126-128), which a rank's contiguous share must be able to do. */
int rgc_tristan_read_range(const char* path, size_t step, unsigned sp, size_t start,
                           size_t count, int ignore_coords, int dim, rgc_particles_t** out,
                           size_t* ntotal);
/* writes a synthetic Tristan-v2 particle file (test / bench fixture generator):
 * datasets named as above for species sp, n floats each, contiguous float32.
 * columns = x,y,z (only when with_coords), u,v,w, ex,ey,ez, bx,by,bz; columns[k]
 * may be NULL (sparse dataset of zeros; x_/y_/z_ are always created because the
 * reader sizes the species from x_<sp>).  append != 0 adds a species to an
 * existing file.  Host-only (no GPU needed). */
int rgc_tristan_write_species(const char* path, size_t step, unsigned sp, size_t n,
                              int with_coords, const float* const* columns, int append);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif

#ifdef __cplusplus
}
#endif

#endif /* RAGNAR_CUDA_H */
