// Minimal HDF5 emitter for the results layout of the reference's HDF5Writer.
//
// Reference layout (src/results/hdf5_writer.cpp:136-217, base.cpp:123-142, :210-218): one file, one group per
// distinct group string ("Time", "Temperature", "Mass fractions"), one f64 dataset per registered scalar
// (/<group>/<set>), three string attributes Name / Notation / Unit per dataset.  scripts/results/results.py reads it
// as f[group][set][:] and .attrs.  libhdf5 is not available in this environment, so the file is produced directly
// from the HDF5 File Format Specification (version-0 superblock, version-1 object headers, symbol-table groups =
// v1 B-tree + local heap + SNOD nodes, little-endian IEEE f64 datasets), i.e. the oldest, most widely readable
// structures.  String attributes are variable-length strings like the reference's (H5::StrType(0, H5T_VARIABLE),
// hdf5_writer.cpp:194-208: a class-9 datatype whose value is a {length, global-heap collection, object index} reference
// into one "GCOL" collection), so that h5py returns `str` and results.py's `attrs['Notation'] + ' ('` works; fixed-length
// strings stay selectable.  The reference's datasets are 1-D, unlimited, chunked by 1000 and deflated at level 6 because it
// appends one sample at a time (hdf5_writer.cpp:156-181); the batch path knows the final length and by default writes
// contiguous datasets of rank 1 (one reactor) or rank 2 ([reactor][sample]); set_layout(1000, 6) reproduces the
// reference's storage for rank-1 datasets (layout class 2, v1 chunk B-tree, filter pipeline {deflate}, max dim unlimited).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include <zlib.h>

#include "h5_writer.hpp"

namespace ctb {
namespace {

const uint64_t UNDEF = ~0ull;
const int K_LEAF = 32;      // group leaf node K: a symbol table node holds up to 2K = 64 entries
const int K_INTERNAL = 16;  // group internal node K

struct Buf {
  std::vector<uint8_t> b;
  void u8(unsigned v) { b.push_back((uint8_t)v); }
  void u16(unsigned v) { u8(v & 0xff); u8((v >> 8) & 0xff); }
  void u32(uint32_t v) { for (int i = 0; i < 4; ++i) u8((v >> (8 * i)) & 0xff); }
  void u64(uint64_t v) { for (int i = 0; i < 8; ++i) u8((unsigned)((v >> (8 * i)) & 0xff)); }
  void bytes(const void* p, size_t n) { const uint8_t* q = (const uint8_t*)p; b.insert(b.end(), q, q + n); }
  void zeros(size_t n) { b.insert(b.end(), n, 0); }
  void pad8() { while (b.size() % 8) b.push_back(0); }
  size_t size() const { return b.size(); }
};

size_t pad8(size_t n) { return (n + 7) & ~size_t(7); }

// ---- header messages (data only, unpadded)
std::vector<uint8_t> msg_dataspace(const std::vector<uint64_t>& dims, bool unlimited = false) {
  Buf m;
  m.u8(1); m.u8((unsigned)dims.size()); m.u8(unlimited ? 1 : 0); m.u8(0); m.u32(0);
  for (uint64_t d : dims) m.u64(d);
  if (unlimited) for (size_t i = 0; i < dims.size(); ++i) m.u64(UNDEF);  // H5S_UNLIMITED
  return m.b;
}
std::vector<uint8_t> msg_datatype_f64() {
  Buf m;
  m.u8(0x11);              // version 1, class 1 (floating point)
  m.u8(0x20); m.u8(0x3f); m.u8(0x00);  // little-endian, implied mantissa msb, sign bit 63
  m.u32(8);                // size
  m.u16(0); m.u16(64);     // bit offset, precision
  m.u8(52); m.u8(11);      // exponent location, size
  m.u8(0); m.u8(52);       // mantissa location, size
  m.u32(1023);             // exponent bias
  return m.b;
}
std::vector<uint8_t> msg_datatype_string(size_t len) {
  Buf m;
  m.u8(0x13);              // version 1, class 3 (string)
  m.u8(0x00); m.u8(0); m.u8(0);  // null-terminated, ASCII
  m.u32((uint32_t)len);
  return m.b;
}
std::vector<uint8_t> msg_datatype_vlen_string() {
  Buf m;
  m.u8(0x19);              // version 1, class 9 (variable-length)
  m.u8(0x01); m.u8(0x00); m.u8(0);  // type = string, null-terminated, ASCII
  m.u32(16);               // on-disk element: u32 length + u64 collection address + u32 object index
  const std::vector<uint8_t> base = msg_datatype_string(1);  // base type: one character (H5T_C_S1)
  m.bytes(base.data(), base.size());
  return m.b;
}
std::vector<uint8_t> msg_fill_value(bool chunked) {
  Buf m;
  m.u8(2); m.u8(chunked ? 3 : 2); m.u8(0); m.u8(0);  // version 2, late / incremental allocation, write time on allocation, no fill value defined
  return m.b;
}
std::vector<uint8_t> msg_layout_chunked(uint64_t btree, const std::vector<uint64_t>& chunk_dims) {
  Buf m;
  m.u8(3); m.u8(2); m.u8((unsigned)chunk_dims.size() + 1); m.u64(btree);
  for (uint64_t d : chunk_dims) m.u32((uint32_t)d);
  m.u32(8);  // dataset element size
  return m.b;
}
std::vector<uint8_t> msg_filter_deflate(int level) {
  Buf m;
  m.u8(1); m.u8(1); m.u16(0); m.u32(0);                 // version 1, one filter
  m.u16(1); m.u16(8); m.u16(0); m.u16(1);               // H5Z_FILTER_DEFLATE, name length, flags, one client value
  m.bytes("deflate\0", 8);
  m.u32((uint32_t)level); m.u32(0);                     // client data + padding (odd count)
  return m.b;
}
std::vector<uint8_t> msg_layout_contiguous(uint64_t addr, uint64_t size) {
  Buf m;
  m.u8(3); m.u8(1); m.u64(addr); m.u64(size);
  return m.b;
}
// heap = {collection address, object index} of the value in the global heap, or {0, 0} for a fixed-length string
std::vector<uint8_t> msg_attribute(const std::string& name, const std::string& value, std::pair<uint64_t, uint32_t> heap) {
  const bool vlen = heap.second != 0;
  const std::vector<uint8_t> dt = vlen ? msg_datatype_vlen_string() : msg_datatype_string(value.size() + 1), ds = msg_dataspace({});
  Buf m;
  m.u8(1); m.u8(0); m.u16((unsigned)name.size() + 1); m.u16((unsigned)dt.size()); m.u16((unsigned)ds.size());
  m.bytes(name.c_str(), name.size() + 1); m.pad8();
  m.bytes(dt.data(), dt.size()); m.pad8();
  m.bytes(ds.data(), ds.size()); m.pad8();
  if (vlen) { m.u32((uint32_t)value.size()); m.u64(heap.first); m.u32(heap.second); }
  else m.bytes(value.c_str(), value.size() + 1);
  return m.b;
}
std::vector<uint8_t> msg_symbol_table(uint64_t btree, uint64_t heap) {
  Buf m;
  m.u64(btree); m.u64(heap);
  return m.b;
}

struct Msg { unsigned type; std::vector<uint8_t> data; };

std::vector<uint8_t> object_header(const std::vector<Msg>& msgs) {
  size_t body = 0;
  for (auto& m : msgs) body += 8 + pad8(m.data.size());
  Buf h;
  h.u8(1); h.u8(0); h.u16((unsigned)msgs.size()); h.u32(1); h.u32((uint32_t)body); h.u32(0);  // 16-byte prefix
  for (auto& m : msgs) {
    h.u16(m.type); h.u16((unsigned)pad8(m.data.size())); h.u8(0); h.u8(0); h.u8(0); h.u8(0);
    h.bytes(m.data.data(), m.data.size()); h.pad8();
  }
  return h.b;
}

struct Node {
  std::string name;
  bool is_group = true;
  std::map<std::string, std::unique_ptr<Node>> children;  // std::map: strcmp order, as the B-tree requires
  // dataset
  std::vector<uint64_t> dims;
  std::vector<double> data;
  std::vector<std::pair<std::string, std::string>> attrs;
  uint64_t chunk = 0;  // > 0: rank-1 dataset stored in chunks of that many elements, unlimited max dim
  int deflate = 0;
  std::vector<std::vector<uint8_t>> chunks;    // stored bytes of every chunk (after the filter)
  std::vector<uint64_t> chunk_addr, cleaf_addr;  // chunk data / chunk B-tree leaf nodes
  uint64_t croot_addr = 0;
  // layout
  uint64_t header_addr = 0, btree_addr = 0, heap_addr = 0, heap_data_addr = 0, data_addr = 0;
  std::vector<uint64_t> snod_addr;
  std::vector<uint64_t> name_off;  // heap offset per child (map order)
  size_t heap_data_size = 0, header_size = 0;
};

size_t btree_node_size() { return 24 + (2 * K_INTERNAL + 1) * 8 + 2 * K_INTERNAL * 8; }
size_t snod_size() { return 8 + 2 * K_LEAF * 40; }

// global heap of the attribute strings: unique values, one collection (object indices are 16 bits wide)
struct StringHeap {
  bool on = true;
  std::map<std::string, uint32_t> index;
  std::vector<const std::string*> order;
  uint64_t addr = 0;
  void add(const std::string& v) {
    if (!on || index.count(v)) return;
    const uint32_t i = (uint32_t)index.size() + 1;
    if (i > 65535) throw std::runtime_error("too many distinct attribute strings for one global heap collection");
    order.push_back(&index.emplace(v, i).first->first);
  }
  std::pair<uint64_t, uint32_t> ref(const std::string& v) const {
    if (!on) return {0, 0};
    return {addr, index.at(v)};
  }
  size_t size() const {  // collection size: header + objects + the free-space object, at least 4096 (H5HG_MINSIZE)
    size_t sz = 16;
    for (auto* v : order) sz += 16 + pad8(v->size());
    sz += 16;
    return std::max<size_t>(4096, (sz + 4095) & ~size_t(4095));
  }
  std::vector<uint8_t> bytes() const {
    const size_t total = size();
    Buf h;
    h.bytes("GCOL", 4); h.u8(1); h.u8(0); h.u8(0); h.u8(0); h.u64(total);
    for (auto* v : order) {
      h.u16(index.at(*v)); h.u16(1); h.u32(0); h.u64(v->size());
      h.bytes(v->data(), v->size()); h.pad8();
    }
    const size_t rest = total - h.size();  // object 0 = the free space, its size includes its own header
    h.u16(0); h.u16(0); h.u32(0); h.u64(rest);
    h.zeros(total - h.size());
    return h.b;
  }
};

const int K_CHUNK = 32;  // chunk B-tree K (the library default; a version-0 superblock cannot say otherwise)
size_t chunk_key_size() { return 8 + 8 * 2; }  // rank 1: size, filter mask, offset, 0
size_t chunk_node_size() { return 24 + (2 * K_CHUNK + 1) * chunk_key_size() + 2 * K_CHUNK * 8; }

std::vector<Msg> dataset_msgs(const Node& n, const StringHeap& sh) {
  std::vector<Msg> m;
  m.push_back({0x0001, msg_dataspace(n.dims, n.chunk > 0)});
  m.push_back({0x0003, msg_datatype_f64()});
  m.push_back({0x0005, msg_fill_value(n.chunk > 0)});
  if (n.chunk > 0) {
    if (n.deflate > 0) m.push_back({0x000B, msg_filter_deflate(n.deflate)});
    m.push_back({0x0008, msg_layout_chunked(n.chunks.empty() ? UNDEF : (n.croot_addr ? n.croot_addr : n.cleaf_addr[0]), {n.chunk})});
  } else {
    m.push_back({0x0008, msg_layout_contiguous(n.data_addr, n.data.size() * 8)});
  }
  for (auto& a : n.attrs) m.push_back({0x000C, msg_attribute(a.first, a.second, sh.ref(a.second))});
  return m;
}

// stored form of the chunks of a rank-1 dataset: full chunks (the tail is zero-filled), deflated if asked
void make_chunks(Node& n) {
  n.chunks.clear();
  if (n.chunk == 0) return;
  const size_t ne = n.data.size();
  for (size_t c0 = 0; c0 < ne; c0 += n.chunk) {
    std::vector<double> raw(n.chunk, 0.0);
    std::copy(n.data.begin() + c0, n.data.begin() + std::min(ne, c0 + (size_t)n.chunk), raw.begin());
    std::vector<uint8_t> out(raw.size() * 8);
    std::memcpy(out.data(), raw.data(), out.size());
    if (n.deflate > 0) {
      uLongf dl = compressBound((uLong)out.size());
      std::vector<uint8_t> z(dl);
      if (compress2(z.data(), &dl, out.data(), (uLong)out.size(), n.deflate) != Z_OK) throw std::runtime_error("deflate failed");
      z.resize(dl);
      out.swap(z);
    }
    n.chunks.push_back(std::move(out));
  }
  if (n.chunks.size() > (size_t)4 * K_CHUNK * K_CHUNK) throw std::runtime_error("too many chunks for a two-level chunk B-tree");
}

// assign addresses depth first; returns the next free address
uint64_t layout(Node& n, uint64_t at, const StringHeap& sh) {
  at = pad8(at);
  if (n.is_group) {
    n.header_addr = at;
    n.header_size = object_header({{0x0011, msg_symbol_table(0, 0)}}).size();
    at += n.header_size;
    // local heap data: "" at offset 0, then the child names
    size_t off = 8;
    n.name_off.clear();
    for (auto& kv : n.children) { n.name_off.push_back(off); off += pad8(kv.first.size() + 1); }
    n.heap_data_size = std::max<size_t>(off, 16);
    n.heap_addr = at; at += 32;
    n.heap_data_addr = at; at += n.heap_data_size;
    n.btree_addr = at; at += btree_node_size();
    const size_t nsn = std::max<size_t>(1, (n.children.size() + 2 * K_LEAF - 1) / (2 * K_LEAF));
    if (nsn > (size_t)2 * K_INTERNAL) throw std::runtime_error("too many entries in one HDF5 group for a single B-tree node");
    n.snod_addr.clear();
    for (size_t i = 0; i < nsn; ++i) { n.snod_addr.push_back(at); at += snod_size(); }
    for (auto& kv : n.children) at = layout(*kv.second, at, sh);
  } else {
    n.header_addr = at;
    make_chunks(n);
    if (n.chunk > 0 && !n.chunks.empty()) { n.cleaf_addr.assign((n.chunks.size() + 2 * K_CHUNK - 1) / (2 * K_CHUNK), 0); n.croot_addr = n.cleaf_addr.size() > 1 ? 1 : 0; }
    n.header_size = object_header(dataset_msgs(n, sh)).size();  // sizes only: the addresses are assigned below
    at += n.header_size;
    at = pad8(at);
    if (n.chunk > 0) {
      for (auto& a : n.cleaf_addr) { a = at; at += chunk_node_size(); }
      if (n.cleaf_addr.size() > 1) { n.croot_addr = at; at += chunk_node_size(); } else n.croot_addr = 0;
      n.chunk_addr.clear();
      for (auto& c : n.chunks) { n.chunk_addr.push_back(at); at = pad8(at + c.size()); }
    } else {
      n.data_addr = at;
      at += n.data.size() * 8;
    }
  }
  return at;
}

void put(std::vector<uint8_t>& file, uint64_t addr, const std::vector<uint8_t>& bytes) {
  if (file.size() < addr + bytes.size()) file.resize(addr + bytes.size(), 0);
  std::memcpy(file.data() + addr, bytes.data(), bytes.size());
}

// one node of a rank-1 chunk B-tree (node type 1): keys {stored size, filter mask, element offset, 0} around child pointers
std::vector<uint8_t> chunk_node(int level, const std::vector<uint64_t>& child, const std::vector<uint32_t>& nbytes, const std::vector<uint64_t>& offset,
                                uint64_t end_offset, uint64_t left, uint64_t right) {
  Buf t;
  t.bytes("TREE", 4); t.u8(1); t.u8((unsigned)level); t.u16((unsigned)child.size()); t.u64(left); t.u64(right);
  for (size_t i = 0; i < child.size(); ++i) {
    t.u32(nbytes[i]); t.u32(0); t.u64(offset[i]); t.u64(0);
    t.u64(child[i]);
  }
  t.u32(0); t.u32(0); t.u64(end_offset); t.u64(0);
  t.zeros(chunk_node_size() - t.size());
  return t.b;
}

void emit(const Node& n, std::vector<uint8_t>& file, const StringHeap& sh) {
  if (!n.is_group) {
    put(file, n.header_addr, object_header(dataset_msgs(n, sh)));
    if (n.chunk > 0) {
      const size_t per = 2 * K_CHUNK, nc = n.chunks.size();
      std::vector<uint64_t> root_child, root_off;
      std::vector<uint32_t> root_nb;
      for (size_t l = 0; l < n.cleaf_addr.size(); ++l) {
        const size_t lo = l * per, hi = std::min(nc, lo + per);
        std::vector<uint64_t> child, off;
        std::vector<uint32_t> nb;
        for (size_t c = lo; c < hi; ++c) { child.push_back(n.chunk_addr[c]); off.push_back(c * n.chunk); nb.push_back((uint32_t)n.chunks[c].size()); }
        put(file, n.cleaf_addr[l], chunk_node(0, child, nb, off, hi * n.chunk, l ? n.cleaf_addr[l - 1] : UNDEF,
                                              l + 1 < n.cleaf_addr.size() ? n.cleaf_addr[l + 1] : UNDEF));
        root_child.push_back(n.cleaf_addr[l]); root_off.push_back(lo * n.chunk); root_nb.push_back(nb[0]);
      }
      if (n.croot_addr) put(file, n.croot_addr, chunk_node(1, root_child, root_nb, root_off, nc * n.chunk, UNDEF, UNDEF));
      for (size_t c = 0; c < nc; ++c) put(file, n.chunk_addr[c], n.chunks[c]);
      return;
    }
    std::vector<uint8_t> raw(n.data.size() * 8);
    if (!raw.empty()) std::memcpy(raw.data(), n.data.data(), raw.size());  // host is little-endian (x86-64 / aarch64)
    put(file, n.data_addr, raw);
    return;
  }
  put(file, n.header_addr, object_header({{0x0011, msg_symbol_table(n.btree_addr, n.heap_addr)}}));
  {  // local heap
    Buf h;
    h.bytes("HEAP", 4); h.u8(0); h.u8(0); h.u8(0); h.u8(0);
    h.u64(n.heap_data_size);
    h.u64(1);  // H5HL_FREE_NULL: no free block
    h.u64(n.heap_data_addr);
    put(file, n.heap_addr, h.b);
    std::vector<uint8_t> d(n.heap_data_size, 0);
    size_t i = 0;
    for (auto& kv : n.children) { std::memcpy(d.data() + n.name_off[i], kv.first.c_str(), kv.first.size()); ++i; }
    put(file, n.heap_data_addr, d);
  }
  // symbol table nodes + B-tree leaf node
  std::vector<const Node*> kids;
  for (auto& kv : n.children) kids.push_back(kv.second.get());
  const size_t per = 2 * K_LEAF;
  Buf t;
  t.bytes("TREE", 4); t.u8(0); t.u8(0); t.u16((unsigned)n.snod_addr.size()); t.u64(UNDEF); t.u64(UNDEF);
  t.u64(0);  // key 0: the empty string at heap offset 0
  for (size_t s = 0; s < n.snod_addr.size(); ++s) {
    const size_t lo = s * per, hi = std::min(kids.size(), lo + per);
    Buf sn;
    sn.bytes("SNOD", 4); sn.u8(1); sn.u8(0); sn.u16((unsigned)(hi - lo));
    for (size_t k = lo; k < hi; ++k) {
      const Node* c = kids[k];
      sn.u64(n.name_off[k]); sn.u64(c->header_addr);
      if (c->is_group) { sn.u32(1); sn.u32(0); sn.u64(c->btree_addr); sn.u64(c->heap_addr); }
      else { sn.u32(0); sn.u32(0); sn.u64(0); sn.u64(0); }
    }
    sn.zeros(snod_size() - sn.size());
    put(file, n.snod_addr[s], sn.b);
    t.u64(n.snod_addr[s]);                        // child pointer
    t.u64(hi > lo ? n.name_off[hi - 1] : 0);      // key: largest name in that child
  }
  t.zeros(btree_node_size() - t.size());
  put(file, n.btree_addr, t.b);
  for (auto& kv : n.children) emit(*kv.second, file, sh);
}

void collect_strings(const Node& n, StringHeap& sh) {
  for (auto& a : n.attrs) sh.add(a.second);
  for (auto& kv : n.children) collect_strings(*kv.second, sh);
}

}  // namespace

struct H5File::Impl { Node root; bool vlen_strings = true; uint64_t chunk = 0; int deflate = 0; };

H5File::H5File() : impl_(new Impl) {}
H5File::~H5File() { delete impl_; }

static Node* descend(Node& root, const std::string& path, bool create_last_as_group, std::string* leaf) {
  Node* cur = &root;
  size_t i = 0;
  std::vector<std::string> parts;
  while (i < path.size()) {
    size_t j = path.find('/', i);
    if (j == std::string::npos) j = path.size();
    if (j > i) parts.push_back(path.substr(i, j - i));
    i = j + 1;
  }
  if (parts.empty()) throw std::runtime_error("empty HDF5 path");
  const size_t ngroups = create_last_as_group ? parts.size() : parts.size() - 1;
  for (size_t k = 0; k < ngroups; ++k) {
    auto& slot = cur->children[parts[k]];
    if (!slot) { slot.reset(new Node); slot->name = parts[k]; slot->is_group = true; }
    if (!slot->is_group) throw std::runtime_error("HDF5 path component is a dataset: " + parts[k]);
    cur = slot.get();
  }
  if (leaf) *leaf = parts.back();
  return cur;
}

void H5File::set_layout(uint64_t chunk_elems, int deflate_level, bool vlen_strings) {
  if (deflate_level < 0 || deflate_level > 9) throw std::runtime_error("deflate level must be 0..9");
  impl_->chunk = chunk_elems; impl_->deflate = chunk_elems ? deflate_level : 0; impl_->vlen_strings = vlen_strings;
}

void H5File::create_group(const std::string& path) { descend(impl_->root, path, true, nullptr); }

void H5File::create_dataset(const std::string& path, const std::vector<uint64_t>& dims, const double* data,
                            const std::vector<std::pair<std::string, std::string>>& attrs) {
  std::string leaf;
  Node* g = descend(impl_->root, path, false, &leaf);
  if (g->children.count(leaf)) throw std::runtime_error("HDF5 object exists: " + path);
  uint64_t n = 1;
  for (uint64_t d : dims) n *= d;
  std::unique_ptr<Node> d(new Node);
  d->name = leaf; d->is_group = false; d->dims = dims; d->attrs = attrs;
  d->data.assign(data, data + n);
  if (dims.size() == 1 && impl_->chunk > 0) { d->chunk = impl_->chunk; d->deflate = impl_->deflate; }  // the reference's storage, rank 1 only
  g->children[leaf] = std::move(d);
}

void H5File::write(const std::string& filename) {
  Node& root = impl_->root;
  StringHeap sh;
  sh.on = impl_->vlen_strings;
  if (sh.on) collect_strings(root, sh);
  if (sh.order.empty()) sh.on = false;
  uint64_t at = 96;
  if (sh.on) { sh.addr = at; at += sh.size(); }
  const uint64_t eof = layout(root, at, sh);
  std::vector<uint8_t> file((size_t)eof, 0);
  if (sh.on) put(file, sh.addr, sh.bytes());
  emit(root, file, sh);
  Buf sb;
  const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
  sb.bytes(sig, 8);
  sb.u8(0); sb.u8(0); sb.u8(0); sb.u8(0); sb.u8(0); sb.u8(8); sb.u8(8); sb.u8(0);
  sb.u16(K_LEAF); sb.u16(K_INTERNAL); sb.u32(0);
  sb.u64(0); sb.u64(UNDEF); sb.u64(eof); sb.u64(UNDEF);
  sb.u64(0); sb.u64(root.header_addr); sb.u32(1); sb.u32(0); sb.u64(root.btree_addr); sb.u64(root.heap_addr);
  put(file, 0, sb.b);
  std::ofstream f(filename, std::ios::binary);
  if (!f) throw std::runtime_error("cannot open " + filename + " for writing");
  f.write((const char*)file.data(), (std::streamsize)file.size());
  if (!f) throw std::runtime_error("write failed: " + filename);
}

}  // namespace ctb
