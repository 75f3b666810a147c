// Minimal self-contained HDF5 layer (no libhdf5 / HighFive in this image).
//
// Stands in for the HighFive calls of the reference: File / getDataSet /
// getDimensions / select({start},{count},{stride}).read<T>() / createDataSet<T> /
// write_raw (src/plugins/tristan-v2.cpp:51-74,117,125; src/io/h5.cpp:21-23,40,54-62).
//
// Read side (files written by libhdf5): superblock v0-v3, object headers v1/v2,
// old-style groups (B-tree v1 + local heap + SNOD) and compact new-style groups
// (link messages), dataspace v1/v2, IEEE float and fixed-point datatypes of either
// byte order, data layouts compact / contiguous / chunked (B-tree v1 index, and the
// v4 single-chunk / implicit / fixed-array indexes) with the deflate, shuffle and
// fletcher32 filters.  Write side: superblock v0 files with an old-style root group
// and contiguous 1-D datasets — the shape libhdf5 itself produces with default
// property lists (and what Tristan-v2's Fortran writer emits).
//
// Pure host code, no CUDA: usable and tested without a GPU.
#ifndef RGC_H5_HPP
#define RGC_H5_HPP

#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace rgc::h5 {

  struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
  };

  constexpr std::uint64_t kUndef = ~std::uint64_t(0);

  enum TypeClass : int { kFixed = 0, kFloat = 1 };
  enum Layout : int { kCompact = 0, kContiguous = 1, kChunked = 2 };
  enum OpenMode : int { kReadOnly = 0, kReadWrite = 1, kTruncate = 2 };

  struct ChunkRec {
    std::uint64_t elem_offset; // first element of the chunk along dim 0
    std::uint64_t addr;        // file address (relative to base), kUndef = never written
    std::uint64_t nbytes;      // stored (filtered) size
    std::uint32_t filter_mask; // bit i set = filter i skipped
  };

  struct Dataset {
    std::string                name;
    std::uint64_t              header_addr { kUndef };
    std::vector<std::uint64_t> dims;
    std::uint64_t              nelem { 0 };
    int                        type_class { -1 };
    std::uint32_t              elem_size { 0 };
    bool                       big_endian { false };
    bool                       is_signed { false };
    int                        layout { -1 };
    std::uint64_t              data_addr { kUndef }; // contiguous
    std::uint64_t              data_bytes { 0 };
    std::vector<std::uint8_t>  compact;
    std::vector<std::uint64_t> chunk_dims; // element counts per dim (without the size dim)
    std::vector<std::uint16_t> filters;    // pipeline order
    std::vector<ChunkRec>      chunks;     // rank-1 chunked datasets, sorted by elem_offset
  };

  class File {
  public:
    File(const std::string& path, int mode);
    ~File();
    File(const File&)            = delete;
    File& operator=(const File&) = delete;

    const std::string& path() const { return m_path; }
    int                fd() const { return m_fd; }
    std::uint64_t      base() const { return m_base; }
    int                superblock_version() const { return m_sb_version; }

    // names of the links of a group ("/" = root), sorted
    std::vector<std::string> list(const std::string& group = "/");
    bool                     exists(const std::string& name);
    Dataset                  dataset(const std::string& name);

    // raw source-typed elements [first, first+n) of the flattened dataset -> dst
    // (n * elem_size bytes).  Thread-safe for concurrent calls on one File.
    void read_raw(const Dataset& ds, std::uint64_t first, std::uint64_t n, void* dst) const;
    // dst[i] = T(src[i * stride]) for i < n_out; src holds source-typed elements
    static void convert(const Dataset& ds, const void* src, std::uint64_t n_out,
                        std::uint64_t stride, int out_dtype, void* dst);
    // selection {start},{count},{stride} converted to out_dtype (RGC_I32/F32/F64)
    void read(const Dataset& ds, std::uint64_t start, std::uint64_t count,
              std::uint64_t stride, int out_dtype, void* out) const;

    // ---- writer (mode != kReadOnly; root group must be old-style)
    // contiguous 1-D dataset of n elements of dtype, zero-filled (sparse) until written
    void create_dataset(const std::string& name, int dtype, std::uint64_t n);
    // elements [start, start+count) of a contiguous native-typed dataset <- data
    void write(const Dataset& ds, std::uint64_t start, std::uint64_t count, const void* data);
    // (re)writes the root group index and the superblock; called by the destructor
    void flush();

  private:
    struct Msg {
      std::uint16_t             type;
      std::uint8_t              flags;
      std::vector<std::uint8_t> body;
      std::uint64_t             body_addr; // absolute file position of the body
    };
    std::vector<std::uint8_t> rd(std::uint64_t addr, std::uint64_t n) const;
    void                      pread_abs(std::uint64_t pos, void* dst, std::uint64_t n) const;
    void                      pwrite_abs(std::uint64_t pos, const void* src, std::uint64_t n);
    std::vector<Msg>          object_header(std::uint64_t addr) const;
    std::map<std::string, std::uint64_t> group_links(std::uint64_t header_addr) const;
    void walk_group_btree(std::uint64_t btree, std::uint64_t heap_data,
                          std::map<std::string, std::uint64_t>& out, int depth) const;
    void walk_chunk_btree(std::uint64_t btree, unsigned rank, std::vector<ChunkRec>& out,
                          int depth) const;
    std::uint64_t resolve(const std::string& name, std::string* leaf = nullptr) const;
    Msg           deshare(const Msg& m) const;
    void          parse_layout(const Msg& m, Dataset& ds) const;
    void          read_chunk(const Dataset& ds, const ChunkRec& c, std::uint64_t chunk_elems,
                             std::vector<std::uint8_t>& out) const;
    std::uint64_t alloc(std::uint64_t nbytes, std::uint64_t align);
    void          load_superblock();
    void          init_new_file();

    std::string   m_path;
    int           m_fd { -1 };
    int           m_mode { kReadOnly };
    std::uint64_t m_base { 0 };
    int           m_sb_version { 0 };
    unsigned      m_size_offsets { 8 }, m_size_lengths { 8 };
    unsigned      m_leaf_k { 4 }, m_internal_k { 16 };
    std::uint64_t m_root_header { kUndef };
    std::uint64_t m_eof { 0 }; // end-of-allocation address (relative to base)
    // writer state
    bool                                 m_dirty { false };
    std::uint64_t                        m_root_symtab_body { kUndef }; // absolute position
    std::map<std::string, std::uint64_t> m_root_links;
  };

} // namespace rgc::h5

#endif // RGC_H5_HPP
