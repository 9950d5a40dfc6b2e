#include "hdf5_min.hpp"

#include <cstring>

namespace caffe {

static const uint64_t kUndef = ~0ULL;

uint64_t H5File::u(uint64_t off, int n) const {
    check(off, (uint64_t)n);
    uint64_t v = 0;
    for (int i = n - 1; i >= 0; i--) v = (v << 8) | p_[off + i];
    return v;
}
void H5File::check(uint64_t off, uint64_t n) const {
    if (off > n_ || n > n_ - off) throw H5Error("structure points outside the file (offset " + std::to_string(off) + ")");
}

H5File::H5File(const void* data, size_t n) : p_(static_cast<const uint8_t*>(data)), n_(n) {
    static const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    if (n < 96 || memcmp(p_, sig, 8)) throw H5Error("not an HDF5 file (signature)");
    const int version = (int)u(8, 1);
    if (version != 0 && version != 1) throw H5Error("superblock version " + std::to_string(version) + " is not supported (only 0 / 1: what libhdf5 writes by default)");
    if (u(13, 1) != 8 || u(14, 1) != 8) throw H5Error("only 8-byte offsets and lengths are supported");
    // v0: sig 8, versions 8, K 4, flags 4 -> 24; v1 adds indexed-storage K (2) + reserved (2) -> 28
    uint64_t off = version == 0 ? 24 : 28;
    if (u(off, 8) != 0) throw H5Error("non-zero base address");
    off += 32;                                             // base, free-space, end-of-file, driver-info addresses
    // root group symbol table entry: name offset 8, object header 8, cache type 4, reserved 4, scratch 16
    const uint64_t header = u(off + 8, 8);
    walk_object("/", header, 0);
}

// version-1 object header: the messages that matter for weight files
void H5File::walk_object(const std::string& path, uint64_t header, int depth) {
    if (depth > 16) throw H5Error("groups nested too deeply");
    if (u(header, 1) != 1) throw H5Error("object header version " + std::to_string(u(header, 1)) + " of '" + path + "' is not supported (new-style files: write with libver='earliest')");
    const int nmsg = (int)u(header + 2, 2);
    uint64_t block = header + 16, block_end = header + 16 + u(header + 8, 4);
    std::vector<std::pair<uint64_t, uint64_t> > more;       // continuation blocks
    Obj obj;
    bool have_space = false, have_type = false, have_layout = false;
    uint64_t btree = kUndef, heap = kUndef;
    int seen = 0;
    size_t next_block = 0;
    while (seen < nmsg) {
        if (block + 8 > block_end) {
            if (next_block >= more.size()) break;
            block = more[next_block].first; block_end = block + more[next_block].second; next_block++;
            continue;
        }
        const int type = (int)u(block, 2);
        const uint64_t size = u(block + 2, 2), body = block + 8;
        check(body, size);
        seen++;
        if (type == 0x0010) {                              // continuation: offset, length
            more.push_back({u(body, 8), u(body + 8, 8)});
        } else if (type == 0x0011) {                       // symbol table: B-tree, local heap
            btree = u(body, 8); heap = u(body + 8, 8);
        } else if (type == 0x0001) {                       // dataspace
            const int ver = (int)u(body, 1), rank = (int)u(body + 1, 1);
            if (ver != 1 && ver != 2) throw H5Error("dataspace version " + std::to_string(ver));
            uint64_t d = body + (ver == 1 ? 8 : 4);
            if (ver == 2 && u(body + 3, 1) != 1 && rank) throw H5Error("only simple dataspaces are supported");
            obj.ds.dims.clear();
            for (int i = 0; i < rank; i++) {
                const uint64_t v = u(d + 8 * i, 8);
                if (v > 0x7fffffffULL) throw H5Error("dimension too large");
                obj.ds.dims.push_back((int)v);
            }
            have_space = true;
        } else if (type == 0x0003) {                       // datatype
            const int cls = (int)(u(body, 1) & 15), bits0 = (int)u(body + 1, 1);
            const int sz = (int)u(body + 4, 4);
            if (cls != 1) throw H5Error("dataset '" + path + "' is not floating point (datatype class " + std::to_string(cls) + ")");
            if (bits0 & 1) throw H5Error("big-endian floats are not supported");
            if (sz != 4 && sz != 8) throw H5Error("float size " + std::to_string(sz));
            obj.ds.elem_size = sz;
            have_type = true;
        } else if (type == 0x0008) {                       // data layout
            const int ver = (int)u(body, 1);
            if (ver == 3) {
                const int cls = (int)u(body + 1, 1);
                if (cls == 1) { obj.ds.address = u(body + 2, 8); obj.ds.bytes = u(body + 10, 8); }
                else if (cls == 0) { obj.ds.bytes = u(body + 2, 2); obj.ds.address = body + 4; }
                else throw H5Error("dataset '" + path + "' is chunked / filtered (e.g. gzip); only contiguous datasets are supported");
            } else if (ver == 1 || ver == 2) {
                const int rank = (int)u(body + 1, 1), cls = (int)u(body + 2, 1);
                if (cls != 1) throw H5Error("dataset '" + path + "' is not contiguous (old-style layout class " + std::to_string(cls) + ")");
                obj.ds.address = u(body + 8, 8);
                obj.ds.bytes = 0;                          // derived from the dataspace below
                (void)rank;
            } else {
                throw H5Error("data layout version " + std::to_string(ver));
            }
            have_layout = true;
        } else if (type == 0x000B) {
            throw H5Error("dataset '" + path + "' uses a filter pipeline (e.g. gzip); not supported");
        } else if (type == 0x0002 || type == 0x0006) {
            throw H5Error("'" + path + "' is a new-style group (link messages); write the file with libver='earliest'");
        }
        block = body + ((size + 7) & ~7ULL);
    }
    if (btree != kUndef) {
        obj.is_group = true;
        walk_group(path, btree, heap, &obj, depth);
    } else if (have_space && have_type && have_layout) {
        uint64_t count = 1;
        for (int d : obj.ds.dims) count *= (uint64_t)d;
        const uint64_t need = count * obj.ds.elem_size;
        if (!obj.ds.bytes) obj.ds.bytes = need;
        if (obj.ds.address == kUndef) throw H5Error("dataset '" + path + "' has no storage allocated");
        if (obj.ds.bytes < need) throw H5Error("dataset '" + path + "' is smaller than its dataspace");
        check(obj.ds.address, need);
        datasets_.push_back({path, obj.ds});
    } else {
        throw H5Error("'" + path + "' is neither a symbol-table group nor a plain dataset");
    }
    objects_[path] = obj;
}

void H5File::walk_group(const std::string& path, uint64_t btree, uint64_t heap, Obj* g, int depth) {
    check(heap, 32);
    if (memcmp(p_ + heap, "HEAP", 4)) throw H5Error("bad local heap signature");
    const uint64_t heap_data = u(heap + 24, 8);
    walk_btree(btree, heap_data, path, g, depth);
}

void H5File::walk_btree(uint64_t node, uint64_t heap_data, const std::string& path, Obj* g, int depth) {
    check(node, 24);
    if (memcmp(p_ + node, "TREE", 4)) throw H5Error("bad B-tree signature");
    if (u(node + 4, 1) != 0) throw H5Error("B-tree node is not a group node");
    const int level = (int)u(node + 5, 1), used = (int)u(node + 6, 2);
    uint64_t q = node + 24;                                // key0, child0, key1, child1, ..., key_used
    for (int i = 0; i < used; i++) {
        const uint64_t child = u(q + 8, 8);
        q += 16;
        if (level > 0) { walk_btree(child, heap_data, path, g, depth); continue; }
        check(child, 8);
        if (memcmp(p_ + child, "SNOD", 4)) throw H5Error("bad symbol node signature");
        const int nsym = (int)u(child + 6, 2);
        for (int s = 0; s < nsym; s++) {
            const uint64_t e = child + 8 + 40ULL * s;
            const uint64_t name_off = u(e, 8), header = u(e + 8, 8);
            check(heap_data + name_off, 1);
            const char* nm = reinterpret_cast<const char*>(p_ + heap_data + name_off);
            const size_t maxlen = n_ - (heap_data + name_off);
            const size_t len = strnlen(nm, maxlen);
            if (len == maxlen) throw H5Error("unterminated link name");
            const std::string name(nm, len);
            g->children.push_back(name);
            walk_object(path == "/" ? "/" + name : path + "/" + name, header, depth + 1);
        }
    }
}

std::vector<std::string> H5File::links(const std::string& group) const {
    auto it = objects_.find(group);
    if (it == objects_.end() || !it->second.is_group) throw H5Error("'" + group + "' is not a group of this file");
    return it->second.children;
}

std::vector<float> H5File::read(const std::string& path, std::vector<int>* dims) const {
    auto it = objects_.find(path);
    if (it == objects_.end() || it->second.is_group) throw H5Error("'" + path + "' is not a dataset of this file");
    const H5Dataset& d = it->second.ds;
    uint64_t count = 1;
    for (int x : d.dims) count *= (uint64_t)x;
    std::vector<float> out((size_t)count);
    if (d.elem_size == 4) {
        memcpy(out.data(), p_ + d.address, (size_t)count * 4);
    } else {
        for (uint64_t i = 0; i < count; i++) { double v; memcpy(&v, p_ + d.address + 8 * i, 8); out[(size_t)i] = (float)v; }
    }
    if (dims) *dims = d.dims;
    return out;
}

}  // namespace caffe
