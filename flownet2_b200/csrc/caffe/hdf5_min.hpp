// Minimal HDF5 reader for Caffe weight files (.caffemodel.h5: groups /data/<layer>/<blob index>, contiguous float datasets written
// by H5LTmake_dataset_float, util/hdf5.cpp:130-160; read back by hdf5_load_nd_dataset, :20-76).  There is no HDF5 library in this
// image, so the subset of the file format those files use is parsed here: superblock version 0, version-1 object headers (with
// continuation blocks), symbol-table groups (version-1 B-tree + local heap + SNOD nodes), simple dataspaces, fixed-point-free
// IEEE float / double little-endian datatypes, contiguous and compact layouts.  Chunked / filtered (gzip) datasets, new-style
// (link-message / fractal-heap) groups and superblock versions 2+ are reported as unsupported, loudly.
// Pinned on files written by the real library: the reference's own test data (src/caffe/test/test_data/*.h5, generated with h5py
// by generate_sample_data.py), copied as fixtures into tests/golden/ref_hdf5/.
#pragma once
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace caffe {

struct H5Error : std::runtime_error {
    explicit H5Error(const std::string& m) : std::runtime_error("hdf5: " + m) {}
};

struct H5Dataset {
    std::vector<int> dims;
    int elem_size = 0;          // 4 (float) or 8 (double)
    uint64_t address = 0;       // file offset of the raw data (contiguous) or of the compact data inside the header
    uint64_t bytes = 0;
};

class H5File {
 public:
    H5File(const void* data, size_t n);
    // every dataset of the file by absolute path ("/data/conv1/0"), in symbol-table (= name) order per group
    const std::vector<std::pair<std::string, H5Dataset> >& datasets() const { return datasets_; }
    // the names linked in a group ("/" or "/data" ...), in symbol-table order; throws if the path is not a group
    std::vector<std::string> links(const std::string& group) const;
    bool has(const std::string& path) const { return objects_.count(path) > 0; }
    // dataset contents as float (double is narrowed)
    std::vector<float> read(const std::string& path, std::vector<int>* dims = nullptr) const;

 private:
    struct Obj { bool is_group = false; std::vector<std::string> children; H5Dataset ds; };
    uint64_t u(uint64_t off, int n) const;
    void check(uint64_t off, uint64_t n) const;
    void walk_object(const std::string& path, uint64_t header, int depth);
    void walk_group(const std::string& path, uint64_t btree, uint64_t heap, Obj* g, int depth);
    void walk_btree(uint64_t node, uint64_t heap_data, const std::string& path, Obj* g, int depth);
    const uint8_t* p_;
    size_t n_;
    std::map<std::string, Obj> objects_;
    std::vector<std::pair<std::string, H5Dataset> > datasets_;
};

// Writer for the same subset (Net::ToHDF5, net.cpp:905-960 / hdf5_save_nd_dataset, util/hdf5.cpp:130-160): /data/<layer>/<index>
// float32 datasets with the blobs' shapes.  layers: (name, blobs as (dims, values)).
struct H5Blob { std::vector<int> dims; std::vector<float> data; };
std::string WriteCaffemodelH5(const std::vector<std::pair<std::string, std::vector<H5Blob> > >& layers);

}  // namespace caffe
