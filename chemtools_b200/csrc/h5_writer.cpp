// Minimal HDF5 emitter for the results layout of the reference's HDF5Writer.
//
// Reference layout (src/results/hdf5_writer.cpp:136-217, base.cpp:123-142, :210-218): one file, one group per
// distinct group string ("Time", "Temperature", "Mass fractions"), one f64 dataset per registered scalar
// (/<group>/<set>), three string attributes Name / Notation / Unit per dataset.  scripts/results/results.py reads it
// as f[group][set][:] and .attrs.  libhdf5 is not available in this environment, so the file is produced directly
// from the HDF5 File Format Specification (version-0 superblock, version-1 object headers, symbol-table groups =
// v1 B-tree + local heap + SNOD nodes, contiguous little-endian IEEE f64 datasets, fixed-length string attributes),
// i.e. the oldest, most widely readable structures.  The reference's datasets are 1-D extendible/chunked/deflated
// because it appends one sample at a time; the batch path knows the final length and writes contiguous datasets of
// rank 1 (one reactor) or rank 2 ([reactor][sample]).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

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
std::vector<uint8_t> msg_dataspace(const std::vector<uint64_t>& dims) {
  Buf m;
  m.u8(1); m.u8((unsigned)dims.size()); m.u8(0); m.u8(0); m.u32(0);
  for (uint64_t d : dims) m.u64(d);
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
std::vector<uint8_t> msg_fill_value() {
  Buf m;
  m.u8(2); m.u8(2); m.u8(0); m.u8(0);  // version 2, late allocation, write time on allocation, no fill value defined
  return m.b;
}
std::vector<uint8_t> msg_layout_contiguous(uint64_t addr, uint64_t size) {
  Buf m;
  m.u8(3); m.u8(1); m.u64(addr); m.u64(size);
  return m.b;
}
std::vector<uint8_t> msg_attribute(const std::string& name, const std::string& value) {
  const std::vector<uint8_t> dt = msg_datatype_string(value.size() + 1), ds = msg_dataspace({});
  Buf m;
  m.u8(1); m.u8(0); m.u16((unsigned)name.size() + 1); m.u16((unsigned)dt.size()); m.u16((unsigned)ds.size());
  m.bytes(name.c_str(), name.size() + 1); m.pad8();
  m.bytes(dt.data(), dt.size()); m.pad8();
  m.bytes(ds.data(), ds.size()); m.pad8();
  m.bytes(value.c_str(), value.size() + 1);
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
  // layout
  uint64_t header_addr = 0, btree_addr = 0, heap_addr = 0, heap_data_addr = 0, data_addr = 0;
  std::vector<uint64_t> snod_addr;
  std::vector<uint64_t> name_off;  // heap offset per child (map order)
  size_t heap_data_size = 0, header_size = 0;
};

size_t btree_node_size() { return 24 + (2 * K_INTERNAL + 1) * 8 + 2 * K_INTERNAL * 8; }
size_t snod_size() { return 8 + 2 * K_LEAF * 40; }

std::vector<Msg> dataset_msgs(const Node& n) {
  std::vector<Msg> m;
  m.push_back({0x0001, msg_dataspace(n.dims)});
  m.push_back({0x0003, msg_datatype_f64()});
  m.push_back({0x0005, msg_fill_value()});
  m.push_back({0x0008, msg_layout_contiguous(n.data_addr, n.data.size() * 8)});
  for (auto& a : n.attrs) m.push_back({0x000C, msg_attribute(a.first, a.second)});
  return m;
}

// assign addresses depth first; returns the next free address
uint64_t layout(Node& n, uint64_t at) {
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
    for (auto& kv : n.children) at = layout(*kv.second, at);
  } else {
    n.header_addr = at;
    n.header_size = object_header(dataset_msgs(n)).size();
    at += n.header_size;
    at = pad8(at);
    n.data_addr = at;
    at += n.data.size() * 8;
  }
  return at;
}

void put(std::vector<uint8_t>& file, uint64_t addr, const std::vector<uint8_t>& bytes) {
  if (file.size() < addr + bytes.size()) file.resize(addr + bytes.size(), 0);
  std::memcpy(file.data() + addr, bytes.data(), bytes.size());
}

void emit(const Node& n, std::vector<uint8_t>& file) {
  if (!n.is_group) {
    put(file, n.header_addr, object_header(dataset_msgs(n)));
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
  for (auto& kv : n.children) emit(*kv.second, file);
}

}  // namespace

struct H5File::Impl { Node root; };

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
  g->children[leaf] = std::move(d);
}

void H5File::write(const std::string& filename) {
  Node& root = impl_->root;
  const uint64_t eof = layout(root, 96);
  std::vector<uint8_t> file((size_t)eof, 0);
  emit(root, file);
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
