// Results writer: the HDF5 layout of the reference's Results::HDF5Writer (include/results/hdf5_writer.hpp,
// src/results/hdf5_writer.cpp:136-217) produced without libhdf5.
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

namespace ctb {

class H5File {
 public:
  H5File();
  ~H5File();
  H5File(const H5File&) = delete;
  H5File& operator=(const H5File&) = delete;
  // storage of the datasets created afterwards: chunk_elems > 0 = rank-1 datasets chunked (unlimited max dim), deflated
  // at deflate_level (the reference: 1000, 6; hdf5_writer.cpp:156-181); 0 = contiguous.  vlen_strings (file-wide, default
  // on): attributes as variable-length strings like the reference (hdf5_writer.cpp:194-208), else fixed-length
  void set_layout(uint64_t chunk_elems, int deflate_level, bool vlen_strings);
  void create_group(const std::string& path);
  // little-endian f64 dataset with string attributes (the reference attaches Name / Notation / Unit)
  void create_dataset(const std::string& path, const std::vector<uint64_t>& dims, const double* data,
                      const std::vector<std::pair<std::string, std::string>>& attrs);
  void write(const std::string& filename);

 private:
  struct Impl;
  Impl* impl_;
};

}  // namespace ctb
