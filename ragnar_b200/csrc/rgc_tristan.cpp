// HDF5 I/O entry points of the C-ABI: the generic 1-D array reader / writer
// (reference src/io/h5.cpp:16-68) and the Tristan-v2 particle plugin
// (src/plugins/tristan-v2.cpp:51-188).
//
// Device-bound reads are streamed: a small pool of I/O threads preads slabs of the
// datasets straight into pinned staging buffers (converting f64 / integer / strided
// sources on the way) and issues one async DMA per slab on the thread's own stream,
// so disk reads, host conversion and H2D copies of all columns overlap.
#include "rgc_h5.hpp"
#include "rgc_internal.hpp"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>

using namespace rgc;

struct rgc_h5 {
  std::unique_ptr<h5::File> file;
};

namespace {

  constexpr std::size_t kSlabBytes   = std::size_t(16) << 20; // output bytes per slab
  constexpr int         kMaxThreads  = 16;
  constexpr int         kSlotsPerThr = 2;

  // ------------------------------------------------------- pinned I/O lanes
  struct Lane {
    cudaStream_t stream { nullptr };
    void*        stage[kSlotsPerThr] { nullptr, nullptr };
    cudaEvent_t  done[kSlotsPerThr] { nullptr, nullptr };
  };

  struct LanePool {
    std::mutex mtx;
    std::mutex in_use; // held for a whole stream_to_device call: the lanes serve one reader at a time
    Lane       lanes[kMaxThreads];
    int        ready { 0 };
  };

  LanePool& pool() {
    static LanePool p;
    return p;
  }

  int io_threads() {
    int n = 0;
    if (const char* e = std::getenv("RGC_IO_THREADS")) {
      n = std::atoi(e);
    }
    if (n <= 0) {
      const unsigned hw = std::thread::hardware_concurrency();
      n                 = hw >= 16 ? 8 : (hw >= 4 ? (int)hw / 2 : 2);
    }
    return std::min(std::max(n, 1), kMaxThreads);
  }

  int ensure_lanes(int n) {
    auto&                       p = pool();
    std::lock_guard<std::mutex> lock(p.mtx);
    RGC_CUDA(cudaSetDevice(ctx().device));
    for (; p.ready < n; ++p.ready) {
      Lane& l = p.lanes[p.ready];
      RGC_CUDA(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
      for (int s = 0; s < kSlotsPerThr; ++s) {
        RGC_CUDA(cudaHostAlloc(&l.stage[s], kSlabBytes, cudaHostAllocDefault));
        RGC_CUDA(cudaEventCreateWithFlags(&l.done[s], cudaEventDisableTiming));
      }
    }
    return RGC_OK;
  }

  // one dataset selection bound for device memory
  struct Transfer {
    const h5::File*    file;
    const h5::Dataset* ds;
    std::uint64_t      start, count, stride;
    int                dtype;
    char*              dev;
  };

  struct Slab {
    const Transfer* t;
    std::uint64_t   o0, n; // output element range
  };

  // Streams every transfer to the device; returns when all DMAs have completed.
  int stream_to_device(const std::vector<Transfer>& transfers) {
    std::vector<Slab> slabs;
    for (const auto& t : transfers) {
      const std::uint64_t osz  = dtype_size(t.dtype);
      const std::uint64_t per  = kSlabBytes / std::max<std::uint64_t>(osz, t.ds->elem_size);
      for (std::uint64_t o = 0; o < t.count; o += per) {
        slabs.push_back({ &t, o, std::min(per, t.count - o) });
      }
    }
    if (slabs.empty()) {
      return RGC_OK;
    }
    const int nthr = (int)std::min<std::size_t>((std::size_t)io_threads(), slabs.size());
    // a prefetching reader thread and a main-thread read may arrive together: the second
    // one waits here instead of sharing staging slots with the first
    std::lock_guard<std::mutex> one_reader(pool().in_use);
    RGC_TRY(ensure_lanes(nthr));
    std::atomic<std::size_t> next { 0 };
    std::atomic<bool>        failed { false };
    std::mutex               err_mtx;
    std::string              err_msg;
    int                      err_code = RGC_OK;
    auto report = [&](int code, const std::string& msg) {
      std::lock_guard<std::mutex> lock(err_mtx);
      if (!failed.exchange(true)) {
        err_code = code;
        err_msg  = msg;
      }
    };
    auto worker = [&](int lane_id) {
      Lane& lane = pool().lanes[lane_id];
      if (cudaSetDevice(ctx().device) != cudaSuccess) {
        report(RGC_ERR_CUDA, "cudaSetDevice failed in an I/O thread");
        return;
      }
      std::vector<std::uint8_t> scratch;
      int                       slot = 0;
      try {
        for (;;) {
          const std::size_t k = next.fetch_add(1);
          if (k >= slabs.size() || failed.load()) {
            break;
          }
          const Slab&     s   = slabs[k];
          const Transfer& t   = *s.t;
          const auto&     ds  = *t.ds;
          const std::uint64_t osz = dtype_size(t.dtype);
          cudaError_t     ce  = cudaEventSynchronize(lane.done[slot]);
          if (ce != cudaSuccess) {
            report(RGC_ERR_CUDA, std::string("cudaEventSynchronize: ") + cudaGetErrorString(ce));
            break;
          }
          const bool direct = t.stride == 1 && !ds.big_endian && ds.elem_size == osz &&
                              ((ds.type_class == h5::kFloat && t.dtype != RGC_I32) ||
                               (ds.type_class == h5::kFixed && t.dtype == RGC_I32 &&
                                ds.is_signed));
          const std::uint64_t first = t.start + s.o0 * t.stride;
          if (direct) {
            t.file->read_raw(ds, first, s.n, lane.stage[slot]);
          } else {
            const std::uint64_t nsrc = (s.n - 1) * t.stride + 1;
            // strided sources are fetched in bounded pieces
            const std::uint64_t piece_out =
              std::max<std::uint64_t>(1, (std::uint64_t(8) << 20) / (ds.elem_size * t.stride));
            (void)nsrc;
            for (std::uint64_t o = 0; o < s.n; o += piece_out) {
              const std::uint64_t no = std::min(piece_out, s.n - o);
              const std::uint64_t ns = (no - 1) * t.stride + 1;
              scratch.resize(ns * ds.elem_size);
              t.file->read_raw(ds, first + o * t.stride, ns, scratch.data());
              h5::File::convert(ds, scratch.data(), no, t.stride, t.dtype,
                                static_cast<char*>(lane.stage[slot]) + o * osz);
            }
          }
          ce = cudaMemcpyAsync(t.dev + s.o0 * osz, lane.stage[slot], s.n * osz,
                               cudaMemcpyHostToDevice, lane.stream);
          if (ce == cudaSuccess) {
            ce = cudaEventRecord(lane.done[slot], lane.stream);
          }
          if (ce != cudaSuccess) {
            report(RGC_ERR_CUDA, std::string("H2D copy failed: ") + cudaGetErrorString(ce));
            break;
          }
          slot = (slot + 1) % kSlotsPerThr;
        }
      } catch (const std::exception& e) {
        report(RGC_ERR_IO, e.what());
      }
      cudaStreamSynchronize(lane.stream);
    };
    std::vector<std::thread> threads;
    for (int i = 1; i < nthr; ++i) {
      threads.emplace_back(worker, i);
    }
    worker(0);
    for (auto& th : threads) {
      th.join();
    }
    if (failed.load()) {
      return fail(err_code, "%s", err_msg.c_str());
    }
    return RGC_OK;
  }

  std::string tristan_filename(const char* path, std::size_t step) {
    const std::string s = std::to_string(step);
    return std::string(path) + "/output/prtl/prtl.tot." +
           std::string(s.length() < 5 ? 5 - s.length() : 0, '0') + s;
  }

  const char* const kCoord[3] = { "x", "y", "z" };
  const char* const kVel[3]   = { "u", "v", "w" };

  // read1DArray's checks — reference tristan-v2.cpp:58-72
  std::uint64_t checked_count(const h5::Dataset& ds, std::uint64_t start, std::uint64_t size,
                              std::uint64_t stride) {
    if (ds.dims.size() != 1) {
      throw h5::Error("Dataset is not 1D");
    }
    if (start >= ds.dims[0]) {
      throw h5::Error("Start index out of bounds");
    }
    if (start + size > ds.dims[0]) {
      throw h5::Error("Size exceeds dataset dimensions");
    }
    if (size == 0) {
      size = ds.dims[0] / stride;
    }
    // HighFive's select({start},{size},{stride}) rejects a hyperslab that leaves the extent
    if (size > 0 && start + (size - 1) * stride >= ds.dims[0]) {
      throw h5::Error("Unable to select the hyperslab of \"" + ds.name + "\": start " +
                      std::to_string(start) + " + count " + std::to_string(size) + " x stride " +
                      std::to_string(stride) + " exceeds its extent " +
                      std::to_string(ds.dims[0]));
    }
    return size;
  }

} // namespace

namespace rgc {
  // called by rgc_finalize
  void io_release_lanes() {
    auto&                       p = pool();
    std::lock_guard<std::mutex> lock(p.mtx);
    for (int i = 0; i < p.ready; ++i) {
      Lane& l = p.lanes[i];
      for (int s = 0; s < kSlotsPerThr; ++s) {
        cudaFreeHost(l.stage[s]);
        cudaEventDestroy(l.done[s]);
        l.stage[s] = nullptr;
        l.done[s]  = nullptr;
      }
      cudaStreamDestroy(l.stream);
      l.stream = nullptr;
    }
    p.ready = 0;
  }
} // namespace rgc

#define RGC_H5_GUARD(...)                                   \
  try {                                                     \
    __VA_ARGS__                                             \
  } catch (const h5::Error& e) {                            \
    return fail(RGC_ERR_IO, "%s", e.what());                \
  } catch (const std::bad_alloc&) {                         \
    return fail(RGC_ERR_OOM, "out of host memory");         \
  } catch (const std::exception& e) {                       \
    return fail(RGC_ERR_IO, "%s", e.what());                \
  }

extern "C" {

// ------------------------------------------------------------ host-side HDF5
int rgc_h5_open(const char* filename, int mode, rgc_h5_t** out) {
  if (!filename || !out || mode < 0 || mode > 2) {
    return fail(RGC_ERR_INVALID, "rgc_h5_open: bad argument");
  }
  RGC_H5_GUARD({
    auto h  = std::make_unique<rgc_h5>();
    h->file = std::make_unique<h5::File>(filename, mode);
    *out    = h.release();
  })
  return RGC_OK;
}

int rgc_h5_close(rgc_h5_t* f) {
  if (!f) {
    return RGC_OK;
  }
  int rc = RGC_OK;
  try {
    f->file->flush();
  } catch (const std::exception& e) {
    rc = fail(RGC_ERR_IO, "%s", e.what());
  }
  delete f;
  return rc;
}

int rgc_h5_list(rgc_h5_t* f, const char* group, char* names, size_t cap, size_t* needed) {
  if (!f) {
    return fail(RGC_ERR_INVALID, "rgc_h5_list: null file");
  }
  RGC_H5_GUARD({
    std::string joined;
    for (const auto& n : f->file->list(group ? group : "/")) {
      joined += n;
      joined += '\n';
    }
    if (needed) {
      *needed = joined.size() + 1;
    }
    if (names && cap) {
      const std::size_t k = std::min(cap - 1, joined.size());
      std::memcpy(names, joined.data(), k);
      names[k] = 0;
    }
  })
  return RGC_OK;
}

int rgc_h5_dataset_info(rgc_h5_t* f, const char* name, int* rank, uint64_t* dims, int max_rank,
                        int* type_class, int* elem_size, int* layout) {
  if (!f || !name) {
    return fail(RGC_ERR_INVALID, "rgc_h5_dataset_info: bad argument");
  }
  RGC_H5_GUARD({
    const auto ds = f->file->dataset(name);
    if (rank) {
      *rank = (int)ds.dims.size();
    }
    for (int k = 0; dims && k < max_rank && k < (int)ds.dims.size(); ++k) {
      dims[k] = ds.dims[(std::size_t)k];
    }
    if (type_class) {
      *type_class = ds.type_class;
    }
    if (elem_size) {
      *elem_size = (int)ds.elem_size;
    }
    if (layout) {
      *layout = ds.layout;
    }
  })
  return RGC_OK;
}

int rgc_h5_read(rgc_h5_t* f, const char* name, size_t start, size_t count, size_t stride,
                int dtype, void* host) {
  if (!f || !name || (count && !host) || dtype < RGC_I32 || dtype > RGC_F64) {
    return fail(RGC_ERR_INVALID, "rgc_h5_read: bad argument");
  }
  RGC_H5_GUARD({
    const auto ds = f->file->dataset(name);
    f->file->read(ds, start, count, stride, dtype, host);
  })
  return RGC_OK;
}

int rgc_h5_create_dataset(rgc_h5_t* f, const char* name, int dtype, size_t n) {
  if (!f || !name || dtype < RGC_I32 || dtype > RGC_F64) {
    return fail(RGC_ERR_INVALID, "rgc_h5_create_dataset: bad argument");
  }
  RGC_H5_GUARD({ f->file->create_dataset(name, dtype, n); })
  return RGC_OK;
}

int rgc_h5_write(rgc_h5_t* f, const char* name, size_t start, size_t count, int dtype,
                 const void* host) {
  if (!f || !name || (count && !host)) {
    return fail(RGC_ERR_INVALID, "rgc_h5_write: bad argument");
  }
  RGC_H5_GUARD({
    const auto ds = f->file->dataset(name);
    const bool ok = (dtype == RGC_I32 && ds.type_class == h5::kFixed && ds.elem_size == 4) ||
                    (dtype == RGC_F32 && ds.type_class == h5::kFloat && ds.elem_size == 4) ||
                    (dtype == RGC_F64 && ds.type_class == h5::kFloat && ds.elem_size == 8);
    if (!ok) {
      return fail(RGC_ERR_INVALID, "rgc_h5_write: dtype does not match dataset \"%s\"", name);
    }
    f->file->write(ds, start, count, host);
  })
  return RGC_OK;
}

// ------------------------------------------- io::h5::Read1DArray / Write1DArray
int rgc_h5_read_array(const char* filename, const char* dsetname, int dtype, size_t size,
                      size_t stride, rgc_buf_t** out) {
  RGC_REQUIRE_INIT();
  if (!filename || !dsetname || !out || dtype < RGC_I32 || dtype > RGC_F64) {
    return fail(RGC_ERR_INVALID, "rgc_h5_read_array: bad argument");
  }
  rgc_buf_t* buf = nullptr;
  RGC_H5_GUARD({
    h5::File   file(filename, h5::kReadOnly);
    const auto ds = file.dataset(dsetname);
    if (ds.dims.size() != 1) { // reference h5.cpp:24-26
      return fail(RGC_ERR_INVALID, "Dataset is not 1D");
    }
    if (stride == 0) { // :27-28
      return fail(RGC_ERR_INVALID, "Stride must be greater than 0");
    } else if (size == 0) { // :29-30
      size = ds.dims[0];
    } else if (ds.dims[0] / stride > size) { // :31-34
      return fail(RGC_ERR_INVALID, "Number of read quantity exceeds allocated space");
    }
    if (size > 0 && (size - 1) * stride >= ds.dims[0]) { // HighFive select() bound check
      return fail(RGC_ERR_IO,
                  "Unable to select the hyperslab of \"%s\": count %zu x stride %zu exceeds its "
                  "extent %llu", dsetname, size, stride, (unsigned long long)ds.dims[0]);
    }
    RGC_TRY(rgc_buf_create(dtype, size, &buf));
    std::vector<Transfer> tr { { &file, &ds, 0, size, stride, dtype,
                                 static_cast<char*>(rgc_buf_device_ptr(buf)) } };
    // rgc_buf_create zero-fills on the compute stream; order the DMAs behind it
    cudaStreamSynchronize(ctx().stream);
    const int rc = stream_to_device(tr);
    if (rc != RGC_OK) {
      rgc_buf_release(buf);
      return rc;
    }
  })
  *out = buf;
  return RGC_OK;
}

int rgc_h5_write_array(const char* filename, const char* dsetname, const rgc_buf_t* array) {
  RGC_REQUIRE_INIT();
  if (!filename || !dsetname || !array) {
    return fail(RGC_ERR_INVALID, "rgc_h5_write_array: bad argument");
  }
  const std::size_t n     = rgc_buf_size(array);
  const int         dtype = rgc_buf_dtype(array);
  std::vector<char> host(n * dtype_size(dtype));
  RGC_TRY(rgc_buf_to_host(array, 0, n, host.data()));
  RGC_H5_GUARD({
    h5::File file(filename, h5::kReadWrite); // HighFive ReadWrite | Create, h5.cpp:54-55
    file.create_dataset(dsetname, dtype, n);
    file.write(file.dataset(dsetname), 0, n, host.data());
    file.flush();
  })
  return RGC_OK;
}

// ------------------------------------------------------------ Tristan-v2 plugin
// range_mode: exactly [start, start + size) with start + size <= ntotal (the sharded
// reader of the multi-GPU path); otherwise the reference's selection rules
static int read_particles_impl(const char* path, size_t step, unsigned sp, size_t start,
                               size_t size, size_t stride, int ignore_coords, int dim,
                               rgc_particles_t** out, size_t* ntotal_out, size_t* nread_out,
                               bool range_mode) {
  RGC_REQUIRE_INIT();
  if (!path || !out) {
    return fail(RGC_ERR_INVALID, "rgc_tristan_read_particles: bad argument");
  }
  if (range_mode && (stride != 1 || size == 0)) {
    return fail(RGC_ERR_INVALID, "rgc_tristan_read_range: count must be > 0");
  }
  if (dim < 1 || dim > 3) {
    return fail(RGC_ERR_INVALID, "dim must be 1, 2 or 3 (got %d)", dim);
  }
  // reference tristan-v2.cpp:102-107
  if (stride == 0) {
    return fail(RGC_ERR_INVALID, "Stride must be greater than 0");
  } else if (stride != 1 && size != 0) {
    return fail(RGC_ERR_INVALID, "Size must be determined automatically (0) when stride != 1");
  }
  rgc_particles_t* prtls = nullptr;
  RGC_H5_GUARD({
    h5::File          file(tristan_filename(path, step), h5::kReadOnly);
    const std::string sfx = "_" + std::to_string(sp);
    // :125-129 — the particle count comes from x_<sp> even when coordinates are ignored
    const auto xds = file.dataset("x" + sfx);
    if (xds.dims.empty()) {
      return fail(RGC_ERR_IO, "Dataset is not 1D");
    }
    const std::uint64_t ntotal = xds.dims[0];
    if (range_mode ? start + size > ntotal : start + size >= ntotal) {
      return fail(RGC_ERR_INVALID, "start + size >= total number of particles");
    }
    const std::uint64_t nparticles = size == 0 ? ntotal / stride : size;
    if (ntotal_out) {
      *ntotal_out = ntotal;
    }
    // gather and validate every dataset before any device work (:58-72,:142-144)
    struct Column {
      int         q, comp;
      h5::Dataset ds;
    };
    std::vector<Column> cols;
    auto add = [&](int q, int comp, const std::string& name) {
      Column c { q, comp, file.dataset(name + sfx) };
      if (checked_count(c.ds, start, size, stride) != nparticles) {
        throw h5::Error("Number of particles mismatch");
      }
      cols.push_back(std::move(c));
    };
    if (!ignore_coords) {
      for (int d = 0; d < dim; ++d) {
        add(RGC_Q_X, d, kCoord[d]);
      }
    }
    for (int d = 0; d < 3; ++d) {
      add(RGC_Q_U, d, kVel[d]);
      add(RGC_Q_E, d, std::string("e") + kCoord[d]);
      add(RGC_Q_B, d, std::string("b") + kCoord[d]);
    }
    RGC_TRY(rgc_particles_create(dim, &prtls));
    int rc = rgc_particles_allocate(prtls, nparticles, ignore_coords ? 0 : 1);
    if (rc == RGC_OK) {
      // the columns are zero-filled on the compute stream; the DMAs run on I/O streams
      cudaStreamSynchronize(ctx().stream);
      std::vector<Transfer> tr;
      for (const auto& c : cols) {
        tr.push_back({ &file, &c.ds, start, nparticles, stride, RGC_F32,
                       reinterpret_cast<char*>(prtls->col[c.q][c.comp]) });
      }
      rc = stream_to_device(tr);
    }
    if (rc != RGC_OK) {
      rgc_particles_release(prtls);
      return rc;
    }
    if (nread_out) {
      *nread_out = nparticles;
    }
  })
  *out = prtls;
  return RGC_OK;
}

int rgc_tristan_read_particles(const char* path, size_t step, unsigned sp, size_t start,
                               size_t size, size_t stride, int ignore_coords, int dim,
                               rgc_particles_t** out, size_t* ntotal_out, size_t* nread_out) {
  return read_particles_impl(path, step, sp, start, size, stride, ignore_coords, dim, out,
                             ntotal_out, nread_out, false);
}

int rgc_tristan_read_range(const char* path, size_t step, unsigned sp, size_t start,
                           size_t count, int ignore_coords, int dim, rgc_particles_t** out,
                           size_t* ntotal_out) {
  size_t nread = 0;
  return read_particles_impl(path, step, sp, start, count, 1, ignore_coords, dim, out,
                             ntotal_out, &nread, true);
}

int rgc_tristan_write_species(const char* path, size_t step, unsigned sp, size_t n,
                              int with_coords, const float* const* columns, int append) {
  if (!path || !columns) {
    return fail(RGC_ERR_INVALID, "rgc_tristan_write_species: bad argument");
  }
  RGC_H5_GUARD({
    h5::File          file(tristan_filename(path, step), append ? h5::kReadWrite : h5::kTruncate);
    const std::string sfx = "_" + std::to_string(sp);
    std::vector<std::string> names;
    for (int d = 0; d < 3; ++d) {
      names.push_back(std::string(kCoord[d]) + sfx);
    }
    for (int d = 0; d < 3; ++d) {
      names.push_back(std::string(kVel[d]) + sfx);
    }
    for (const char* q : { "e", "b" }) {
      for (int d = 0; d < 3; ++d) {
        names.push_back(std::string(q) + kCoord[d] + sfx);
      }
    }
    // columns: x,y,z (only when with_coords), u,v,w, ex,ey,ez, bx,by,bz.  x_/y_/z_ are
    // always created (the reader sizes the species from x_<sp>); unwritten datasets
    // are sparse zeros.
    for (std::size_t k = 0; k < names.size(); ++k) {
      file.create_dataset(names[k], RGC_F32, n);
      const float* src = nullptr;
      if (k < 3) {
        src = with_coords ? columns[k] : nullptr;
      } else {
        src = columns[with_coords ? k : k - 3];
      }
      if (src && n) {
        file.write(file.dataset(names[k]), 0, n, src);
      }
    }
    file.flush();
  })
  return RGC_OK;
}

} // extern "C"
