// Minimal HDF5 reader / writer — see rgc_h5.hpp.  Format per the published
// "HDF5 File Format Specification Version 3.0" (The HDF Group); HighFive v2.10.1 +
// libhdf5 are the reference's (absent) third-party dependency for this boundary
// (cmake/dependencies.cmake:15-18).
#include "rgc_h5.hpp"

#include "ragnar_cuda.h"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <cerrno>
#include <cstring>

namespace rgc::h5 {

  namespace {

    const unsigned char kSignature[8] = { 0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n' };

    [[noreturn]] void bad(const std::string& what) { throw Error(what); }

    // little-endian field reader over a byte vector, bounds-checked
    struct Cursor {
      const std::uint8_t* p;
      std::size_t         n;
      std::size_t         i { 0 };
      Cursor(const std::vector<std::uint8_t>& v, std::size_t start = 0)
        : p(v.data()), n(v.size()), i(start) {}
      std::uint64_t u(unsigned bytes) {
        if (i + bytes > n) {
          bad("HDF5: truncated structure");
        }
        std::uint64_t v = 0;
        for (unsigned b = 0; b < bytes; ++b) {
          v |= std::uint64_t(p[i + b]) << (8 * b);
        }
        i += bytes;
        return v;
      }
      // an address / length field: all-ones of its width means "undefined"
      std::uint64_t addr(unsigned bytes) {
        const std::uint64_t v = u(bytes);
        if (bytes < 8 && v == ((std::uint64_t(1) << (8 * bytes)) - 1)) {
          return kUndef;
        }
        return v;
      }
      void skip(std::size_t k) {
        if (i + k > n) {
          bad("HDF5: truncated structure");
        }
        i += k;
      }
      bool sig(const char* s) {
        if (i + 4 > n) {
          return false;
        }
        const bool ok = std::memcmp(p + i, s, 4) == 0;
        i += 4;
        return ok;
      }
    };

    void put(std::vector<std::uint8_t>& v, std::uint64_t x, unsigned bytes) {
      for (unsigned b = 0; b < bytes; ++b) {
        v.push_back(std::uint8_t(x >> (8 * b)));
      }
    }
    void put_bytes(std::vector<std::uint8_t>& v, const void* s, std::size_t n) {
      const auto* c = static_cast<const std::uint8_t*>(s);
      v.insert(v.end(), c, c + n);
    }
    void pad_to(std::vector<std::uint8_t>& v, std::size_t multiple) {
      while (v.size() % multiple) {
        v.push_back(0);
      }
    }

    template <class T>
    T bswap(T v) {
      unsigned char b[sizeof(T)];
      std::memcpy(b, &v, sizeof(T));
      std::reverse(b, b + sizeof(T));
      std::memcpy(&v, b, sizeof(T));
      return v;
    }

    template <class S, class D>
    void convert_loop(const void* src, std::uint64_t n, std::uint64_t stride, bool swap, D* dst) {
      const auto* s = static_cast<const unsigned char*>(src);
      for (std::uint64_t i = 0; i < n; ++i) {
        S v;
        std::memcpy(&v, s + i * stride * sizeof(S), sizeof(S));
        if (swap) {
          v = bswap(v);
        }
        dst[i] = static_cast<D>(v);
      }
    }

    template <class D>
    void convert_to(const Dataset& ds, const void* src, std::uint64_t n, std::uint64_t stride,
                    D* dst) {
      const bool swap = ds.big_endian;
      if (ds.type_class == kFloat) {
        if (ds.elem_size == 4) {
          if (std::is_same<D, float>::value && stride == 1 && !swap) {
            std::memcpy(dst, src, n * 4);
          } else {
            convert_loop<float, D>(src, n, stride, swap, dst);
          }
        } else if (ds.elem_size == 8) {
          if (std::is_same<D, double>::value && stride == 1 && !swap) {
            std::memcpy(dst, src, n * 8);
          } else {
            convert_loop<double, D>(src, n, stride, swap, dst);
          }
        } else {
          bad("HDF5: unsupported floating-point size");
        }
        return;
      }
      switch (ds.elem_size) {
        case 1:
          ds.is_signed ? convert_loop<std::int8_t, D>(src, n, stride, false, dst)
                       : convert_loop<std::uint8_t, D>(src, n, stride, false, dst);
          break;
        case 2:
          ds.is_signed ? convert_loop<std::int16_t, D>(src, n, stride, swap, dst)
                       : convert_loop<std::uint16_t, D>(src, n, stride, swap, dst);
          break;
        case 4:
          if (std::is_same<D, int>::value && stride == 1 && !swap) {
            std::memcpy(dst, src, n * 4);
          } else {
            ds.is_signed ? convert_loop<std::int32_t, D>(src, n, stride, swap, dst)
                         : convert_loop<std::uint32_t, D>(src, n, stride, swap, dst);
          }
          break;
        case 8:
          ds.is_signed ? convert_loop<std::int64_t, D>(src, n, stride, swap, dst)
                       : convert_loop<std::uint64_t, D>(src, n, stride, swap, dst);
          break;
        default:
          bad("HDF5: unsupported integer size");
      }
    }

    std::vector<std::string> split_path(const std::string& name) {
      std::vector<std::string> parts;
      std::size_t              i = 0;
      while (i < name.size()) {
        const std::size_t j = name.find('/', i);
        const std::string s = name.substr(i, j == std::string::npos ? std::string::npos : j - i);
        if (!s.empty() && s != ".") {
          parts.push_back(s);
        }
        if (j == std::string::npos) {
          break;
        }
        i = j + 1;
      }
      return parts;
    }

  } // namespace

  // ------------------------------------------------------------------ raw I/O
  void File::pread_abs(std::uint64_t pos, void* dst, std::uint64_t n) const {
    auto*         c    = static_cast<char*>(dst);
    std::uint64_t done = 0;
    while (done < n) {
      const ssize_t r = ::pread(m_fd, c + done, n - done, (off_t)(pos + done));
      if (r < 0) {
        if (errno == EINTR) {
          continue;
        }
        bad("HDF5: read error on " + m_path + ": " + std::strerror(errno));
      }
      if (r == 0) {
        bad("HDF5: unexpected end of file in " + m_path);
      }
      done += (std::uint64_t)r;
    }
  }

  void File::pwrite_abs(std::uint64_t pos, const void* src, std::uint64_t n) {
    const auto*   c    = static_cast<const char*>(src);
    std::uint64_t done = 0;
    while (done < n) {
      const ssize_t r = ::pwrite(m_fd, c + done, n - done, (off_t)(pos + done));
      if (r < 0) {
        if (errno == EINTR) {
          continue;
        }
        bad("HDF5: write error on " + m_path + ": " + std::strerror(errno));
      }
      done += (std::uint64_t)r;
    }
  }

  std::vector<std::uint8_t> File::rd(std::uint64_t addr, std::uint64_t n) const {
    if (addr == kUndef) {
      bad("HDF5: undefined address dereferenced");
    }
    // metadata reads may be sized generously (fixed-size nodes, name pieces); bytes
    // beyond the end of the file read as zero and fail the structure's own checks
    std::vector<std::uint8_t> v(n, 0);
    std::uint64_t             done = 0;
    while (done < n) {
      const ssize_t r = ::pread(m_fd, v.data() + done, n - done, (off_t)(m_base + addr + done));
      if (r < 0) {
        if (errno == EINTR) {
          continue;
        }
        bad("HDF5: read error on " + m_path + ": " + std::strerror(errno));
      }
      if (r == 0) {
        break;
      }
      done += (std::uint64_t)r;
    }
    return v;
  }

  // ----------------------------------------------------------------- open
  File::File(const std::string& path, int mode) : m_path(path), m_mode(mode) {
    if (mode == kReadOnly) {
      m_fd = ::open(path.c_str(), O_RDONLY);
      if (m_fd < 0) {
        bad("Unable to open file " + path + ": " + std::strerror(errno));
      }
    } else {
      int flags = O_RDWR | O_CREAT;
      if (mode == kTruncate) {
        flags |= O_TRUNC;
      }
      m_fd = ::open(path.c_str(), flags, 0644);
      if (m_fd < 0) {
        bad("Unable to open file " + path + " for writing: " + std::strerror(errno));
      }
    }
    try {
      struct stat st;
      if (::fstat(m_fd, &st) != 0) {
        bad("HDF5: cannot stat " + path);
      }
      if (st.st_size == 0 && mode != kReadOnly) {
        init_new_file();
      } else {
        load_superblock();
        if (mode != kReadOnly) {
          // the writer re-indexes the root group; it must be an old-style group
          const auto msgs = object_header(m_root_header);
          for (const auto& m : msgs) {
            if (m.type == 0x11) {
              m_root_symtab_body = m.body_addr;
            }
          }
          if (m_root_symtab_body == kUndef || m_sb_version > 1 || m_size_offsets != 8 ||
              m_size_lengths != 8) {
            bad("HDF5: appending is supported for superblock v0/v1 files with a symbol-table "
                "root group only (" + path + ")");
          }
          m_root_links = group_links(m_root_header);
        }
      }
    } catch (...) {
      ::close(m_fd);
      m_fd = -1;
      throw;
    }
  }

  File::~File() {
    if (m_fd >= 0) {
      try {
        flush();
      } catch (...) {
      }
      ::close(m_fd);
    }
  }

  void File::load_superblock() {
    struct stat st;
    ::fstat(m_fd, &st);
    const std::uint64_t fsize = (std::uint64_t)st.st_size;
    // the superblock sits at 0 or at 512, 1024, 2048, ... (after a user block)
    std::uint64_t pos   = 0;
    bool          found = false;
    while (pos + 8 <= fsize) {
      unsigned char sig[8];
      pread_abs(pos, sig, 8);
      if (std::memcmp(sig, kSignature, 8) == 0) {
        found = true;
        break;
      }
      pos = pos == 0 ? 512 : pos * 2;
    }
    if (!found) {
      bad("Unable to open file " + m_path + ": not an HDF5 file (signature not found)");
    }
    std::vector<std::uint8_t> sb(std::min<std::uint64_t>(fsize - pos, 128));
    pread_abs(pos, sb.data(), sb.size());
    Cursor c(sb, 8);
    m_sb_version = (int)c.u(1);
    if (m_sb_version == 0 || m_sb_version == 1) {
      c.skip(3); // free-space, root-entry versions, reserved
      c.skip(1); // shared header message format version
      m_size_offsets = (unsigned)c.u(1);
      m_size_lengths = (unsigned)c.u(1);
      c.skip(1);
      m_leaf_k     = (unsigned)c.u(2);
      m_internal_k = (unsigned)c.u(2);
      c.skip(4); // consistency flags
      if (m_sb_version == 1) {
        c.skip(4); // indexed-storage internal node K + reserved
      }
      const std::uint64_t base = c.addr(m_size_offsets);
      c.addr(m_size_offsets); // free-space info
      m_eof = c.addr(m_size_offsets);
      c.addr(m_size_offsets); // driver info
      // root group symbol table entry
      c.addr(m_size_offsets); // link name offset
      m_root_header = c.addr(m_size_offsets);
      // Files with a user block record base == 0 in some writers (MATLAB) and the
      // user-block size in others; addresses are relative to the superblock either way.
      m_base = (base == 0 || base == kUndef) ? pos : base;
    } else if (m_sb_version == 2 || m_sb_version == 3) {
      m_size_offsets = (unsigned)c.u(1);
      m_size_lengths = (unsigned)c.u(1);
      c.skip(1);
      const std::uint64_t base = c.addr(m_size_offsets);
      c.addr(m_size_offsets); // superblock extension
      m_eof         = c.addr(m_size_offsets);
      m_root_header = c.addr(m_size_offsets);
      m_base        = (base == 0 || base == kUndef) ? pos : base;
    } else {
      bad("HDF5: unsupported superblock version " + std::to_string(m_sb_version));
    }
    // the stored end-of-file address includes the base address (libhdf5 writes
    // rel_eoa + base_addr); allocation below never reuses bytes the file already holds
    m_eof = (m_eof != kUndef && m_eof >= m_base) ? m_eof - m_base : 0;
    if (fsize > m_base) {
      m_eof = std::max(m_eof, fsize - m_base);
    }
    if (m_size_offsets != 8 && m_size_offsets != 4 && m_size_offsets != 2) {
      bad("HDF5: unsupported size of offsets");
    }
    if (m_root_header == kUndef) {
      bad("HDF5: file has no root group");
    }
  }

  // ------------------------------------------------------- object headers
  std::vector<File::Msg> File::object_header(std::uint64_t addr) const {
    std::vector<Msg> msgs;
    auto             head = rd(addr, 16);
    if (std::memcmp(head.data(), "OHDR", 4) == 0) {
      // ---- version 2
      Cursor c(head, 4);
      if (c.u(1) != 2) {
        bad("HDF5: unsupported object header version");
      }
      const unsigned flags      = (unsigned)c.u(1);
      std::size_t    prefix     = 6;
      if (flags & 0x20) {
        prefix += 16; // access, modification, change, birth times
      }
      if (flags & 0x10) {
        prefix += 4; // max compact / min dense attributes
      }
      const unsigned szbytes = 1u << (flags & 3);
      auto           pre     = rd(addr, prefix + szbytes);
      Cursor         pc(pre, prefix);
      const std::uint64_t chunk0 = pc.u(szbytes);
      struct Block {
        std::uint64_t addr, size;
      };
      std::vector<Block> blocks { { addr + prefix + szbytes, chunk0 } };
      const unsigned     mhdr = 4 + ((flags & 0x04) ? 2 : 0);
      for (std::size_t b = 0; b < blocks.size(); ++b) {
        if (b > 4096) {
          bad("HDF5: object header continuation loop");
        }
        auto   data = rd(blocks[b].addr, blocks[b].size);
        Cursor mc(data);
        while (mc.i + mhdr <= data.size()) {
          Msg m;
          m.type  = (std::uint16_t)mc.u(1);
          const std::size_t sz = (std::size_t)mc.u(2);
          m.flags = (std::uint8_t)mc.u(1);
          if (flags & 0x04) {
            mc.skip(2);
          }
          if (mc.i + sz > data.size()) {
            break; // trailing gap
          }
          m.body_addr = m_base + blocks[b].addr + mc.i;
          m.body.assign(data.begin() + (long)mc.i, data.begin() + (long)(mc.i + sz));
          mc.skip(sz);
          if (m.type == 0x10) {
            Cursor              cc(m.body);
            const std::uint64_t a = cc.addr(m_size_offsets);
            const std::uint64_t l = cc.u(m_size_lengths);
            // continuation chunk: "OCHK" + messages + checksum
            if (l < 8) {
              bad("HDF5: bad continuation block");
            }
            blocks.push_back({ a + 4, l - 8 });
          } else if (m.type != 0) {
            msgs.push_back(std::move(m));
          }
        }
      }
      return msgs;
    }
    // ---- version 1
    Cursor c(head);
    if (c.u(1) != 1) {
      bad("HDF5: unsupported object header version at address " + std::to_string(addr));
    }
    c.skip(1);
    const std::size_t nmsg = (std::size_t)c.u(2);
    c.skip(4); // reference count
    const std::uint64_t hsize = c.u(4);
    struct Block {
      std::uint64_t addr, size;
    };
    std::vector<Block> blocks { { addr + 16, hsize } };
    std::size_t        seen = 0;
    for (std::size_t b = 0; b < blocks.size() && seen < nmsg; ++b) {
      if (b > 4096) {
        bad("HDF5: object header continuation loop");
      }
      auto   data = rd(blocks[b].addr, blocks[b].size);
      Cursor mc(data);
      while (mc.i + 8 <= data.size() && seen < nmsg) {
        Msg m;
        m.type               = (std::uint16_t)mc.u(2);
        const std::size_t sz = (std::size_t)mc.u(2);
        m.flags              = (std::uint8_t)mc.u(1);
        mc.skip(3);
        if (mc.i + sz > data.size()) {
          bad("HDF5: object header message overruns its block");
        }
        m.body_addr = m_base + blocks[b].addr + mc.i;
        m.body.assign(data.begin() + (long)mc.i, data.begin() + (long)(mc.i + sz));
        mc.skip(sz);
        ++seen;
        if (m.type == 0x10) {
          Cursor              cc(m.body);
          const std::uint64_t a = cc.addr(m_size_offsets);
          const std::uint64_t l = cc.u(m_size_lengths);
          blocks.push_back({ a, l });
        } else if (m.type != 0) {
          msgs.push_back(std::move(m));
        }
      }
    }
    return msgs;
  }

  // a message stored in another object header (committed datatype etc.)
  File::Msg File::deshare(const Msg& m) const {
    if (!(m.flags & 0x02)) {
      return m;
    }
    Cursor         c(m.body);
    const unsigned ver = (unsigned)c.u(1);
    std::uint64_t  a   = kUndef;
    if (ver == 1) {
      c.skip(7);
      a = c.addr(m_size_offsets);
    } else if (ver == 2) {
      c.skip(1);
      a = c.addr(m_size_offsets);
    } else if (ver == 3) {
      const unsigned t = (unsigned)c.u(1);
      if (t != 2) {
        bad("HDF5: shared message stored in a heap is not supported");
      }
      a = c.addr(m_size_offsets);
    } else {
      bad("HDF5: unknown shared message version");
    }
    for (const auto& o : object_header(a)) {
      if (o.type == m.type) {
        return o;
      }
    }
    bad("HDF5: shared message not found in its object header");
  }

  // ----------------------------------------------------------------- groups
  void File::walk_group_btree(std::uint64_t btree, std::uint64_t heap_data,
                              std::map<std::string, std::uint64_t>& out, int depth) const {
    if (depth > 16) {
      bad("HDF5: group B-tree too deep");
    }
    const std::size_t node_size = 24 + (2 * m_internal_k + 1) * m_size_lengths +
                                  2 * m_internal_k * m_size_offsets;
    auto   node = rd(btree, node_size);
    Cursor c(node);
    if (!c.sig("TREE")) {
      bad("HDF5: bad group B-tree signature");
    }
    if (c.u(1) != 0) {
      bad("HDF5: B-tree node is not a group node");
    }
    const unsigned level = (unsigned)c.u(1);
    const unsigned used  = (unsigned)c.u(2);
    c.skip(2 * m_size_offsets); // siblings
    for (unsigned e = 0; e < used; ++e) {
      c.skip(m_size_lengths); // key
      const std::uint64_t child = c.addr(m_size_offsets);
      if (level > 0) {
        walk_group_btree(child, heap_data, out, depth + 1);
        continue;
      }
      const std::size_t entry = 2 * m_size_offsets + 8 + 16;
      auto              snod  = rd(child, 8 + 2 * m_leaf_k * entry);
      Cursor            s(snod);
      if (!s.sig("SNOD")) {
        bad("HDF5: bad symbol table node signature");
      }
      s.skip(2);
      const unsigned nsym = (unsigned)s.u(2);
      for (unsigned k = 0; k < nsym; ++k) {
        Cursor              ec(snod, 8 + k * entry);
        const std::uint64_t name_off = ec.u(m_size_offsets);
        const std::uint64_t ohdr     = ec.addr(m_size_offsets);
        // names are NUL-terminated strings in the local heap's data segment
        std::string   name;
        std::uint64_t pos = heap_data + name_off;
        for (;;) {
          auto        piece = rd(pos, 64);
          const auto* z     = (const std::uint8_t*)std::memchr(piece.data(), 0, piece.size());
          if (z) {
            name.append((const char*)piece.data(), (std::size_t)(z - piece.data()));
            break;
          }
          name.append((const char*)piece.data(), piece.size());
          pos += 64;
          if (name.size() > 65536) {
            bad("HDF5: unterminated link name");
          }
        }
        out[name] = ohdr;
      }
    }
  }

  std::map<std::string, std::uint64_t> File::group_links(std::uint64_t header_addr) const {
    std::map<std::string, std::uint64_t> links;
    bool                                 is_group = false;
    for (const auto& m : object_header(header_addr)) {
      if (m.type == 0x11) { // symbol table: B-tree v1 + local heap
        is_group = true;
        Cursor              c(m.body);
        const std::uint64_t btree = c.addr(m_size_offsets);
        const std::uint64_t heap  = c.addr(m_size_offsets);
        auto                hh    = rd(heap, 8 + 2 * m_size_lengths + m_size_offsets);
        Cursor              hc(hh);
        if (!hc.sig("HEAP")) {
          bad("HDF5: bad local heap signature");
        }
        hc.skip(4 + 2 * m_size_lengths);
        const std::uint64_t heap_data = hc.addr(m_size_offsets);
        walk_group_btree(btree, heap_data, links, 0);
      } else if (m.type == 0x06) { // link message (compact new-style group)
        is_group = true;
        Cursor c(m.body);
        if (c.u(1) != 1) {
          bad("HDF5: unsupported link message version");
        }
        const unsigned flags = (unsigned)c.u(1);
        unsigned       ltype = 0;
        if (flags & 0x08) {
          ltype = (unsigned)c.u(1);
        }
        if (flags & 0x04) {
          c.skip(8);
        }
        if (flags & 0x10) {
          c.skip(1);
        }
        const std::uint64_t len = c.u(1u << (flags & 3));
        if (c.i + len > m.body.size()) {
          bad("HDF5: bad link message");
        }
        std::string name((const char*)m.body.data() + c.i, (std::size_t)len);
        c.skip((std::size_t)len);
        if (ltype == 0) {
          links[name] = c.addr(m_size_offsets);
        } // soft / external links are not followed
      } else if (m.type == 0x02) { // link info
        is_group = true;
        Cursor c(m.body);
        c.skip(1);
        const unsigned flags = (unsigned)c.u(1);
        if (flags & 1) {
          c.skip(8);
        }
        const std::uint64_t fheap = c.addr(m_size_offsets);
        if (fheap != kUndef) {
          bad("HDF5: groups with dense link storage (fractal heap) are not supported");
        }
      }
    }
    if (!is_group) {
      bad("HDF5: object is not a group");
    }
    return links;
  }

  std::uint64_t File::resolve(const std::string& name, std::string* leaf) const {
    const auto    parts = split_path(name);
    std::uint64_t cur   = m_root_header;
    for (std::size_t k = 0; k < parts.size(); ++k) {
      std::map<std::string, std::uint64_t> links;
      if (k == 0 && m_mode != kReadOnly) {
        links = m_root_links;
      } else {
        links = group_links(cur);
      }
      const auto it = links.find(parts[k]);
      if (it == links.end()) {
        bad("Unable to open the dataset \"" + name + "\": object '" + parts[k] +
            "' doesn't exist (" + m_path + ")");
      }
      cur = it->second;
    }
    if (leaf) {
      *leaf = parts.empty() ? std::string("/") : parts.back();
    }
    return cur;
  }

  std::vector<std::string> File::list(const std::string& group) {
    const auto               links = split_path(group).empty() && m_mode != kReadOnly
                                       ? m_root_links
                                       : group_links(resolve(group));
    std::vector<std::string> names;
    for (const auto& kv : links) {
      names.push_back(kv.first);
    }
    return names;
  }

  bool File::exists(const std::string& name) {
    try {
      resolve(name);
      return true;
    } catch (const Error&) {
      return false;
    }
  }

  // --------------------------------------------------------------- datasets
  void File::walk_chunk_btree(std::uint64_t btree, unsigned rank, std::vector<ChunkRec>& out,
                              int depth) const {
    if (btree == kUndef) {
      return; // no chunk was ever written
    }
    if (depth > 16) {
      bad("HDF5: chunk B-tree too deep");
    }
    auto   head = rd(btree, 8 + 2 * m_size_offsets);
    Cursor hc(head);
    if (!hc.sig("TREE")) {
      bad("HDF5: bad chunk B-tree signature");
    }
    if (hc.u(1) != 1) {
      bad("HDF5: B-tree node is not a chunk node");
    }
    const unsigned    level = (unsigned)hc.u(1);
    const unsigned    used  = (unsigned)hc.u(2);
    const std::size_t key   = 8 + 8 * (rank + 1);
    auto   body = rd(btree + 8 + 2 * m_size_offsets, used * (key + m_size_offsets) + key);
    Cursor c(body);
    for (unsigned e = 0; e < used; ++e) {
      ChunkRec r;
      r.nbytes      = c.u(4);
      r.filter_mask = (std::uint32_t)c.u(4);
      r.elem_offset = c.u(8);
      c.skip(8 * rank); // remaining dims + the element-size dim
      r.addr = c.addr(m_size_offsets);
      if (level > 0) {
        walk_chunk_btree(r.addr, rank, out, depth + 1);
      } else {
        out.push_back(r);
      }
    }
  }

  void File::parse_layout(const Msg& m, Dataset& ds) const {
    Cursor         c(m.body);
    const unsigned ver = (unsigned)c.u(1);
    if (ver == 1 || ver == 2) {
      const unsigned rank = (unsigned)c.u(1);
      ds.layout           = (int)c.u(1);
      c.skip(5);
      std::uint64_t a = kUndef;
      if (ds.layout != kCompact) {
        a = c.addr(m_size_offsets);
      }
      std::vector<std::uint64_t> d(rank);
      for (auto& x : d) {
        x = c.u(4);
      }
      if (ds.layout == kCompact) {
        const std::uint64_t sz = c.u(4);
        if (c.i + sz > m.body.size()) {
          bad("HDF5: bad compact layout");
        }
        ds.compact.assign(m.body.begin() + (long)c.i, m.body.begin() + (long)(c.i + sz));
      } else if (ds.layout == kContiguous) {
        ds.data_addr  = a;
        ds.data_bytes = ds.nelem * ds.elem_size;
      } else if (ds.layout == kChunked) {
        if (d.empty()) {
          bad("HDF5: bad chunked layout");
        }
        ds.chunk_dims.assign(d.begin(), d.end() - 1);
        if (ds.dims.size() == 1) {
          walk_chunk_btree(a, 1, ds.chunks, 0);
        }
      } else {
        bad("HDF5: unknown data layout class");
      }
      return;
    }
    if (ver != 3 && ver != 4) {
      bad("HDF5: unsupported data layout message version " + std::to_string(ver));
    }
    ds.layout = (int)c.u(1);
    if (ds.layout == kCompact) {
      const std::uint64_t sz = c.u(2);
      if (c.i + sz > m.body.size()) {
        bad("HDF5: bad compact layout");
      }
      ds.compact.assign(m.body.begin() + (long)c.i, m.body.begin() + (long)(c.i + sz));
    } else if (ds.layout == kContiguous) {
      ds.data_addr  = c.addr(m_size_offsets);
      ds.data_bytes = c.u(m_size_lengths);
    } else if (ds.layout == kChunked && ver == 3) {
      const unsigned      rank  = (unsigned)c.u(1); // dataset rank + 1
      const std::uint64_t btree = c.addr(m_size_offsets);
      if (rank < 2) {
        bad("HDF5: bad chunked layout");
      }
      for (unsigned k = 0; k + 1 < rank; ++k) {
        ds.chunk_dims.push_back(c.u(4));
      }
      if (ds.dims.size() == 1) {
        walk_chunk_btree(btree, 1, ds.chunks, 0);
      }
    } else if (ds.layout == kChunked && ver == 4) {
      const unsigned flags = (unsigned)c.u(1);
      const unsigned rank  = (unsigned)c.u(1); // dataset rank + 1
      const unsigned enc   = (unsigned)c.u(1);
      if (rank < 2 || enc == 0 || enc > 8) {
        bad("HDF5: bad chunked layout");
      }
      for (unsigned k = 0; k + 1 < rank; ++k) {
        ds.chunk_dims.push_back(c.u(enc));
      }
      c.skip(enc); // element-size dim
      const unsigned index_type = (unsigned)c.u(1);
      if (ds.dims.size() != 1) {
        return; // only rank-1 chunked datasets are readable; reported at read time
      }
      const std::uint64_t cd      = ds.chunk_dims[0];
      const std::uint64_t nchunks = cd ? (ds.dims[0] + cd - 1) / cd : 0;
      const std::uint64_t cbytes  = cd * ds.elem_size;
      if (index_type == 1) { // single chunk
        ChunkRec r { 0, kUndef, cbytes, 0 };
        if (flags & 0x02) {
          r.nbytes      = c.u(m_size_lengths);
          r.filter_mask = (std::uint32_t)c.u(4);
        }
        r.addr = c.addr(m_size_offsets);
        ds.chunks.push_back(r);
      } else if (index_type == 2) { // implicit: chunks laid out back to back
        const std::uint64_t a = c.addr(m_size_offsets);
        for (std::uint64_t k = 0; k < nchunks; ++k) {
          ds.chunks.push_back({ k * cd, a == kUndef ? kUndef : a + k * cbytes, cbytes, 0 });
        }
      } else if (index_type == 3) { // fixed array
        const unsigned      page_bits = (unsigned)c.u(1);
        const std::uint64_t hdr_addr  = c.addr(m_size_offsets);
        if (hdr_addr == kUndef) {
          return;
        }
        auto   hdr = rd(hdr_addr, 4 + 1 + 1 + 1 + 1 + m_size_lengths + m_size_offsets + 4);
        Cursor h(hdr);
        if (!h.sig("FAHD")) {
          bad("HDF5: bad fixed array header signature");
        }
        h.skip(1);
        const unsigned client     = (unsigned)h.u(1);
        const unsigned entry_size = (unsigned)h.u(1);
        h.skip(1); // page bits (again)
        const std::uint64_t nentries = h.u(m_size_lengths);
        const std::uint64_t dblk     = h.addr(m_size_offsets);
        if (dblk == kUndef) {
          return;
        }
        const std::uint64_t page_n   = std::uint64_t(1) << page_bits;
        const bool          paged    = nentries > page_n;
        const std::uint64_t npages   = paged ? (nentries + page_n - 1) / page_n : 0;
        const std::uint64_t prefix   = 4 + 1 + 1 + m_size_offsets + (paged ? (npages + 7) / 8 : 0);
        // paged: bitmap, the data block's own checksum, then pages of 2^page_bits
        // elements each followed by a page checksum
        const std::uint64_t body_len = paged ? 4 + nentries * entry_size + npages * 4
                                             : nentries * entry_size;
        auto   blk = rd(dblk, prefix + body_len);
        Cursor b(blk);
        if (!b.sig("FADB")) {
          bad("HDF5: bad fixed array data block signature");
        }
        b.skip(prefix - 4);
        if (paged) {
          b.skip(4);
        }
        for (std::uint64_t k = 0; k < nentries && k < nchunks; ++k) {
          if (paged && k > 0 && k % page_n == 0) {
            b.skip(4); // page checksum
          }
          ChunkRec r { k * cd, kUndef, cbytes, 0 };
          r.addr = b.addr(m_size_offsets);
          if (client == 1) {
            r.nbytes      = b.u(entry_size - m_size_offsets - 4);
            r.filter_mask = (std::uint32_t)b.u(4);
          }
          ds.chunks.push_back(r);
        }
      } else {
        bad("HDF5: chunk index type " + std::to_string(index_type) +
            " (extensible array / B-tree v2) is not supported");
      }
    } else {
      bad("HDF5: unsupported data layout class (virtual?)");
    }
  }

  Dataset File::dataset(const std::string& name) {
    Dataset ds;
    ds.header_addr  = resolve(name, &ds.name);
    const auto msgs = object_header(ds.header_addr);
    const Msg *space = nullptr, *type = nullptr, *layout = nullptr, *pipeline = nullptr;
    for (const auto& m : msgs) {
      switch (m.type) {
        case 0x01:
          space = &m;
          break;
        case 0x03:
          type = &m;
          break;
        case 0x08:
          layout = &m;
          break;
        case 0x0B:
          pipeline = &m;
          break;
        default:
          break;
      }
    }
    if (!space || !type || !layout) {
      bad("Unable to open the dataset \"" + name + "\": not a dataset (" + m_path + ")");
    }
    { // dataspace
      const Msg      sm = deshare(*space);
      Cursor         c(sm.body);
      const unsigned ver   = (unsigned)c.u(1);
      const unsigned rank  = (unsigned)c.u(1);
      const unsigned flags = (unsigned)c.u(1);
      (void)flags;
      if (ver == 1) {
        c.skip(5);
      } else if (ver == 2) {
        const unsigned kind = (unsigned)c.u(1);
        if (kind == 2) {
          bad("HDF5: null dataspace");
        }
      } else {
        bad("HDF5: unsupported dataspace version");
      }
      ds.nelem = 1;
      for (unsigned k = 0; k < rank; ++k) {
        ds.dims.push_back(c.u(m_size_lengths));
        ds.nelem *= ds.dims.back();
      }
    }
    { // datatype
      const Msg      tm  = deshare(*type);
      Cursor         c(tm.body);
      const unsigned cv  = (unsigned)c.u(1);
      const unsigned b0  = (unsigned)c.u(1);
      c.skip(2);
      ds.elem_size  = (std::uint32_t)c.u(4);
      ds.type_class = (int)(cv & 0x0f);
      ds.big_endian = (b0 & 1) != 0;
      if (ds.type_class == kFixed) {
        ds.is_signed = (b0 & 0x08) != 0;
        if (ds.elem_size != 1 && ds.elem_size != 2 && ds.elem_size != 4 && ds.elem_size != 8) {
          bad("HDF5: unsupported integer size");
        }
      } else if (ds.type_class == kFloat) {
        c.skip(4); // bit offset, precision
        const unsigned      eloc = (unsigned)c.u(1), esz = (unsigned)c.u(1);
        const unsigned      mloc = (unsigned)c.u(1), msz = (unsigned)c.u(1);
        const std::uint64_t bias = c.u(4);
        const bool f32 = ds.elem_size == 4 && eloc == 23 && esz == 8 && mloc == 0 && msz == 23 &&
                         bias == 127;
        const bool f64 = ds.elem_size == 8 && eloc == 52 && esz == 11 && mloc == 0 &&
                         msz == 52 && bias == 1023;
        if (!f32 && !f64) {
          bad("HDF5: dataset \"" + name + "\" is not IEEE binary32/binary64");
        }
      } else {
        bad("HDF5: dataset \"" + name + "\" has a non-numeric datatype (class " +
            std::to_string(ds.type_class) + ")");
      }
    }
    if (pipeline) {
      const Msg      pm  = deshare(*pipeline);
      Cursor         c(pm.body);
      const unsigned ver = (unsigned)c.u(1);
      const unsigned nf  = (unsigned)c.u(1);
      if (ver == 1) {
        c.skip(6);
      } else if (ver != 2) {
        bad("HDF5: unsupported filter pipeline version");
      }
      for (unsigned k = 0; k < nf; ++k) {
        const unsigned id      = (unsigned)c.u(2);
        unsigned       namelen = 0;
        if (ver == 1 || id >= 256) {
          namelen = (unsigned)c.u(2);
        }
        c.skip(2); // flags
        const unsigned ncd = (unsigned)c.u(2);
        c.skip(namelen);
        c.skip(4 * ncd);
        if (ver == 1 && (ncd & 1)) {
          c.skip(4);
        }
        ds.filters.push_back((std::uint16_t)id);
      }
    }
    parse_layout(deshare(*layout), ds);
    std::sort(ds.chunks.begin(), ds.chunks.end(),
              [](const ChunkRec& a, const ChunkRec& b) { return a.elem_offset < b.elem_offset; });
    return ds;
  }

  // one stored chunk -> chunk_elems * elem_size raw bytes
  void File::read_chunk(const Dataset& ds, const ChunkRec& c, std::uint64_t chunk_elems,
                        std::vector<std::uint8_t>& out) const {
    const std::uint64_t raw = chunk_elems * ds.elem_size;
    if (c.addr == kUndef) {
      out.assign(raw, 0);
      return;
    }
    std::vector<std::uint8_t> buf(c.nbytes);
    pread_abs(m_base + c.addr, buf.data(), c.nbytes);
    // filters are undone in reverse pipeline order
    for (std::size_t k = ds.filters.size(); k-- > 0;) {
      if (c.filter_mask & (1u << k)) {
        continue;
      }
      const unsigned id = ds.filters[k];
      if (id == 1) { // deflate
        std::vector<std::uint8_t> dec(raw + 4);
        uLongf                    dlen = (uLongf)dec.size();
        for (;;) {
          const int rc = ::uncompress(dec.data(), &dlen, buf.data(), (uLong)buf.size());
          if (rc == Z_OK) {
            break;
          }
          if (rc == Z_BUF_ERROR && dec.size() < (std::size_t(1) << 34)) {
            dec.resize(dec.size() * 2);
            dlen = (uLongf)dec.size();
            continue;
          }
          bad("HDF5: inflate failed on a chunk of \"" + ds.name + "\"");
        }
        dec.resize(dlen);
        buf.swap(dec);
      } else if (id == 2) { // shuffle
        const std::size_t es = ds.elem_size, ne = buf.size() / es;
        std::vector<std::uint8_t> un(buf.size());
        for (std::size_t b = 0; b < es; ++b) {
          const std::uint8_t* s = buf.data() + b * ne;
          for (std::size_t e = 0; e < ne; ++e) {
            un[e * es + b] = s[e];
          }
        }
        std::copy(buf.begin() + (long)(ne * es), buf.end(), un.begin() + (long)(ne * es));
        buf.swap(un);
      } else if (id == 3) { // fletcher32: 4-byte checksum trails the data
        if (buf.size() < 4) {
          bad("HDF5: bad fletcher32 chunk");
        }
        buf.resize(buf.size() - 4);
      } else {
        bad("HDF5: filter " + std::to_string(id) + " of \"" + ds.name + "\" is not supported");
      }
    }
    if (buf.size() < raw) {
      bad("HDF5: chunk of \"" + ds.name + "\" is shorter than its dimensions");
    }
    buf.resize(raw);
    out.swap(buf);
  }

  void File::read_raw(const Dataset& ds, std::uint64_t first, std::uint64_t n, void* dst) const {
    if (n == 0) {
      return;
    }
    if (first + n > ds.nelem) {
      bad("HDF5: selection exceeds the extent of \"" + ds.name + "\"");
    }
    const std::uint64_t es = ds.elem_size;
    auto*               o  = static_cast<std::uint8_t*>(dst);
    if (ds.layout == kContiguous) {
      if (ds.data_addr == kUndef) {
        std::memset(o, 0, n * es); // never allocated: fill value
        return;
      }
      pread_abs(m_base + ds.data_addr + first * es, o, n * es);
    } else if (ds.layout == kCompact) {
      if ((first + n) * es > ds.compact.size()) {
        bad("HDF5: compact dataset is shorter than its dataspace");
      }
      std::memcpy(o, ds.compact.data() + first * es, n * es);
    } else {
      if (ds.dims.size() != 1 || ds.chunk_dims.size() != 1 || ds.chunk_dims[0] == 0) {
        bad("HDF5: chunked datasets are supported for rank 1 only (\"" + ds.name + "\")");
      }
      const std::uint64_t       cd = ds.chunk_dims[0];
      std::vector<std::uint8_t> tmp;
      std::uint64_t             pos = first;
      while (pos < first + n) {
        const std::uint64_t c0   = (pos / cd) * cd;
        const std::uint64_t upto = std::min(first + n, c0 + cd);
        const auto          it   = std::lower_bound(
          ds.chunks.begin(), ds.chunks.end(), c0,
          [](const ChunkRec& r, std::uint64_t v) { return r.elem_offset < v; });
        if (it == ds.chunks.end() || it->elem_offset != c0) {
          std::memset(o + (pos - first) * es, 0, (upto - pos) * es); // unwritten chunk
        } else if (ds.filters.empty() && it->addr != kUndef) {
          pread_abs(m_base + it->addr + (pos - c0) * es, o + (pos - first) * es,
                    (upto - pos) * es);
        } else {
          read_chunk(ds, *it, cd, tmp);
          std::memcpy(o + (pos - first) * es, tmp.data() + (pos - c0) * es, (upto - pos) * es);
        }
        pos = upto;
      }
    }
  }

  void File::convert(const Dataset& ds, const void* src, std::uint64_t n_out,
                     std::uint64_t stride, int out_dtype, void* dst) {
    switch (out_dtype) {
      case RGC_I32:
        convert_to<int>(ds, src, n_out, stride, static_cast<int*>(dst));
        break;
      case RGC_F32:
        convert_to<float>(ds, src, n_out, stride, static_cast<float*>(dst));
        break;
      case RGC_F64:
        convert_to<double>(ds, src, n_out, stride, static_cast<double*>(dst));
        break;
      default:
        bad("HDF5: bad output dtype");
    }
  }

  void File::read(const Dataset& ds, std::uint64_t start, std::uint64_t count,
                  std::uint64_t stride, int out_dtype, void* out) const {
    if (count == 0) {
      return;
    }
    if (stride == 0) {
      bad("Stride must be greater than 0");
    }
    if (start + (count - 1) * stride >= ds.nelem) {
      bad("HDF5: selection exceeds the extent of \"" + ds.name + "\"");
    }
    const std::uint64_t osz   = out_dtype == RGC_F64 ? 8 : 4;
    const std::uint64_t block = std::max<std::uint64_t>(1, (std::uint64_t(8) << 20) / (ds.elem_size * stride));
    std::vector<std::uint8_t> tmp;
    for (std::uint64_t o0 = 0; o0 < count; o0 += block) {
      const std::uint64_t no   = std::min(block, count - o0);
      const std::uint64_t nsrc = (no - 1) * stride + 1;
      tmp.resize(nsrc * ds.elem_size);
      read_raw(ds, start + o0 * stride, nsrc, tmp.data());
      convert(ds, tmp.data(), no, stride, out_dtype, static_cast<char*>(out) + o0 * osz);
    }
  }

  // ------------------------------------------------------------------ writer
  std::uint64_t File::alloc(std::uint64_t nbytes, std::uint64_t align) {
    const std::uint64_t a = (m_eof + align - 1) / align * align;
    m_eof                 = a + nbytes;
    m_dirty               = true;
    return a;
  }

  void File::init_new_file() {
    m_base        = 0;
    m_sb_version  = 0;
    m_eof         = 96; // superblock v0 with 8-byte offsets / lengths
    // root group object header (v1): one symbol-table message
    m_root_header = alloc(16 + 24, 8);
    std::vector<std::uint8_t> oh;
    put(oh, 1, 1); // version
    put(oh, 0, 1);
    put(oh, 1, 2);  // number of messages
    put(oh, 1, 4);  // reference count
    put(oh, 24, 4); // header data size
    put(oh, 0, 4);  // pad to 8
    put(oh, 0x11, 2);
    put(oh, 16, 2);
    put(oh, 0, 1);
    put(oh, 0, 3);
    m_root_symtab_body = m_root_header + oh.size();
    put(oh, kUndef, 8);
    put(oh, kUndef, 8);
    pwrite_abs(m_root_header, oh.data(), oh.size());
    m_dirty = true;
    flush();
  }

  void File::create_dataset(const std::string& name, int dtype, std::uint64_t n) {
    if (m_mode == kReadOnly) {
      bad("HDF5: file " + m_path + " is open read-only");
    }
    const auto parts = split_path(name);
    if (parts.size() != 1) {
      bad("HDF5: only root-level dataset names can be created (\"" + name + "\")");
    }
    if (m_root_links.count(parts[0])) {
      bad("Unable to create the dataset \"" + name + "\": name already exists (" + m_path + ")");
    }
    const std::uint64_t es    = dtype == RGC_F64 ? 8 : 4;
    const std::uint64_t bytes = n * es;
    std::vector<std::uint8_t> msgs;
    auto begin_msg = [&](unsigned type, unsigned size, unsigned flags) {
      put(msgs, type, 2);
      put(msgs, size, 2);
      put(msgs, flags, 1);
      put(msgs, 0, 3);
    };
    // dataspace v1, rank 1, no max dims
    begin_msg(0x01, 16, 0);
    put(msgs, 1, 1);
    put(msgs, 1, 1);
    put(msgs, 0, 1);
    put(msgs, 0, 5);
    put(msgs, n, 8);
    // datatype v1
    if (dtype == RGC_I32) {
      begin_msg(0x03, 16, 1);
      put(msgs, 0x10, 1); // class 0 (fixed-point), version 1
      put(msgs, 0x08, 1); // little-endian, signed
      put(msgs, 0, 2);
      put(msgs, 4, 4);
      put(msgs, 0, 2);  // bit offset
      put(msgs, 32, 2); // precision
      put(msgs, 0, 4);
    } else {
      const bool d = dtype == RGC_F64;
      begin_msg(0x03, 24, 1);
      put(msgs, 0x11, 1);            // class 1 (floating-point), version 1
      put(msgs, 0x20, 1);            // little-endian, implied mantissa msb
      put(msgs, d ? 63 : 31, 1);     // sign bit position
      put(msgs, 0, 1);
      put(msgs, d ? 8 : 4, 4);       // size
      put(msgs, 0, 2);               // bit offset
      put(msgs, d ? 64 : 32, 2);     // precision
      put(msgs, d ? 52 : 23, 1);     // exponent location
      put(msgs, d ? 11 : 8, 1);      // exponent size
      put(msgs, 0, 1);               // mantissa location
      put(msgs, d ? 52 : 23, 1);     // mantissa size
      put(msgs, d ? 1023 : 127, 4);  // exponent bias
      put(msgs, 0, 4);
    }
    // fill value v2: allocate late, write if set, default (zero-size) value
    begin_msg(0x05, 8, 1);
    put(msgs, 2, 1);
    put(msgs, 2, 1);
    put(msgs, 2, 1);
    put(msgs, 1, 1);
    put(msgs, 0, 4);
    // layout v3, contiguous: address + size (patched below)
    begin_msg(0x08, 24, 1);
    const std::size_t layout_pos = msgs.size();
    put(msgs, 3, 1);
    put(msgs, 1, 1);
    put(msgs, 0, 8);
    put(msgs, bytes, 8);
    put(msgs, 0, 6);

    const std::uint64_t hdr  = alloc(16 + msgs.size(), 8);
    // large columns start on a 4 KiB boundary (page-aligned reads / O_DIRECT friendly)
    const std::uint64_t data = bytes ? alloc(bytes, bytes >= 65536 ? 4096 : 8) : kUndef;
    for (unsigned b = 0; b < 8; ++b) {
      msgs[layout_pos + 2 + b] = std::uint8_t(data >> (8 * b));
    }
    std::vector<std::uint8_t> oh;
    put(oh, 1, 1);
    put(oh, 0, 1);
    put(oh, 4, 2);
    put(oh, 1, 4);
    put(oh, msgs.size(), 4);
    put(oh, 0, 4);
    put_bytes(oh, msgs.data(), msgs.size());
    pwrite_abs(m_base + hdr, oh.data(), oh.size());
    if (::ftruncate(m_fd, (off_t)(m_base + m_eof)) != 0) {
      bad("HDF5: cannot extend " + m_path + ": " + std::strerror(errno));
    }
    m_root_links[parts[0]] = hdr;
    m_dirty                = true;
  }

  void File::write(const Dataset& ds, std::uint64_t start, std::uint64_t count, const void* data) {
    if (m_mode == kReadOnly) {
      bad("HDF5: file " + m_path + " is open read-only");
    }
    if (ds.layout != kContiguous || ds.big_endian || ds.data_addr == kUndef) {
      bad("HDF5: write needs a contiguous native-endian dataset");
    }
    if (start + count > ds.nelem) {
      bad("HDF5: write exceeds the extent of \"" + ds.name + "\"");
    }
    pwrite_abs(m_base + ds.data_addr + start * ds.elem_size, data, count * ds.elem_size);
  }

  // Re-index the root group: a fresh local heap, symbol-table nodes and B-tree v1
  // holding every link, appended at the end of the file; the root object header's
  // symbol-table message and the superblock are then patched in place.  (The
  // previous index, a few hundred bytes, becomes unreferenced space.)
  void File::flush() {
    if (m_mode == kReadOnly || !m_dirty || m_fd < 0) {
      return;
    }
    // ---- local heap data segment: "" at offset 0, then the names
    std::vector<std::uint8_t>  seg(8, 0);
    std::vector<std::uint64_t> name_off;
    std::vector<std::uint64_t> ohdr;
    for (const auto& kv : m_root_links) { // std::map iterates in strcmp order
      name_off.push_back(seg.size());
      ohdr.push_back(kv.second);
      put_bytes(seg, kv.first.c_str(), kv.first.size() + 1);
      pad_to(seg, 8);
    }
    const std::uint64_t heap_hdr  = alloc(32, 8);
    const std::uint64_t heap_data = alloc(seg.size(), 8);
    std::vector<std::uint8_t> hh;
    put_bytes(hh, "HEAP", 4);
    put(hh, 0, 4);
    put(hh, seg.size(), 8);
    put(hh, 1, 8); // H5HL_FREE_NULL: no free block
    put(hh, heap_data, 8);
    pwrite_abs(m_base + heap_hdr, hh.data(), hh.size());
    pwrite_abs(m_base + heap_data, seg.data(), seg.size());
    // ---- symbol table nodes, 2*leafK entries each
    const std::size_t per_node = 2 * m_leaf_k;
    struct Child {
      std::uint64_t addr, last_key;
    };
    std::vector<Child> level;
    const std::size_t  n = name_off.size();
    for (std::size_t i = 0; i < n || (n == 0 && level.empty()); i += per_node) {
      const std::size_t          cnt = std::min(per_node, n - i);
      std::vector<std::uint8_t> sn;
      put_bytes(sn, "SNOD", 4);
      put(sn, 1, 1);
      put(sn, 0, 1);
      put(sn, cnt, 2);
      for (std::size_t k = 0; k < per_node; ++k) {
        if (k < cnt) {
          put(sn, name_off[i + k], 8);
          put(sn, ohdr[i + k], 8);
        } else {
          put(sn, 0, 16);
        }
        put(sn, 0, 8);  // cache type 0 + reserved
        put(sn, 0, 16); // scratch pad
      }
      const std::uint64_t a = alloc(sn.size(), 8);
      pwrite_abs(m_base + a, sn.data(), sn.size());
      level.push_back({ a, cnt ? name_off[i + cnt - 1] : 0 });
      if (n == 0) {
        break;
      }
    }
    // ---- B-tree v1 levels, up to 2*internalK children per node
    const std::size_t per_tree  = 2 * m_internal_k;
    const std::size_t node_size = 24 + (per_tree + 1) * 8 + per_tree * 8;
    unsigned          lvl       = 0;
    for (;;) {
      std::vector<Child>         next;
      const std::size_t          nnodes = (level.size() + per_tree - 1) / per_tree;
      std::vector<std::uint64_t> addrs(nnodes);
      for (auto& a : addrs) {
        a = alloc(node_size, 8);
      }
      for (std::size_t j = 0; j < nnodes; ++j) {
        const std::size_t i0  = j * per_tree;
        const std::size_t cnt = std::min(per_tree, level.size() - i0);
        std::vector<std::uint8_t> t;
        put_bytes(t, "TREE", 4);
        put(t, 0, 1);
        put(t, lvl, 1);
        put(t, cnt, 2);
        put(t, j > 0 ? addrs[j - 1] : kUndef, 8);
        put(t, j + 1 < nnodes ? addrs[j + 1] : kUndef, 8);
        put(t, i0 > 0 ? level[i0 - 1].last_key : 0, 8); // key 0: below every name in the node
        for (std::size_t k = 0; k < cnt; ++k) {
          put(t, level[i0 + k].addr, 8);
          put(t, level[i0 + k].last_key, 8);
        }
        t.resize(node_size, 0);
        pwrite_abs(m_base + addrs[j], t.data(), t.size());
        next.push_back({ addrs[j], level[i0 + cnt - 1].last_key });
      }
      level.swap(next);
      ++lvl;
      if (level.size() == 1) {
        break;
      }
    }
    const std::uint64_t btree = level[0].addr;
    // ---- patch the root object header's symbol-table message
    std::vector<std::uint8_t> st;
    put(st, btree, 8);
    put(st, heap_hdr, 8);
    pwrite_abs(m_root_symtab_body, st.data(), st.size());
    // ---- superblock
    if (m_sb_version == 0 && m_base == 0) {
      std::vector<std::uint8_t> sb;
      put_bytes(sb, kSignature, 8);
      put(sb, 0, 1); // superblock version
      put(sb, 0, 1); // free-space storage version
      put(sb, 0, 1); // root group symbol table entry version
      put(sb, 0, 1);
      put(sb, 0, 1); // shared header message format version
      put(sb, 8, 1); // size of offsets
      put(sb, 8, 1); // size of lengths
      put(sb, 0, 1);
      put(sb, m_leaf_k, 2);
      put(sb, m_internal_k, 2);
      put(sb, 0, 4);      // file consistency flags
      put(sb, 0, 8);      // base address
      put(sb, kUndef, 8); // free-space info
      put(sb, m_eof, 8);  // end-of-file address
      put(sb, kUndef, 8); // driver information block
      put(sb, 0, 8);      // root entry: link name offset
      put(sb, m_root_header, 8);
      put(sb, 1, 4); // cache type 1: scratch holds B-tree + heap addresses
      put(sb, 0, 4);
      put(sb, btree, 8);
      put(sb, heap_hdr, 8);
      pwrite_abs(0, sb.data(), sb.size());
    } else {
      // foreign v0/v1 file: patch the end-of-file address and the cached root entry
      const std::uint64_t sb   = m_base; // superblock position == base for these files
      const std::uint64_t fix  = m_sb_version == 1 ? 4 : 0;
      std::vector<std::uint8_t> e;
      put(e, m_eof + m_base, 8);
      pwrite_abs(sb + 24 + fix + 16, e.data(), 8);
      std::vector<std::uint8_t> r;
      put(r, 1, 4);
      put(r, 0, 4);
      put(r, btree, 8);
      put(r, heap_hdr, 8);
      pwrite_abs(sb + 24 + fix + 32 + 16, r.data(), r.size());
    }
    if (::ftruncate(m_fd, (off_t)(m_base + m_eof)) != 0) {
      bad("HDF5: cannot extend " + m_path + ": " + std::strerror(errno));
    }
    m_dirty = false;
  }

} // namespace rgc::h5
