#include "hdf5_min.hpp"

#include <algorithm>
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

// ---- writer ------------------------------------------------------------------------------------------------------------------
namespace {
struct H5Out {
    std::string buf;
    H5Out() : buf(96, '\0') {}                              // superblock (56) + root symbol table entry (40), filled in last
    void put(uint64_t v, int n) { for (int i = 0; i < n; i++) buf.push_back((char)((v >> (8 * i)) & 0xff)); }
    uint64_t begin() { while (buf.size() % 8) buf.push_back('\0'); return buf.size(); }
    void set(size_t off, uint64_t v, int n) { for (int i = 0; i < n; i++) buf[off + i] = (char)((v >> (8 * i)) & 0xff); }
    void message(int type, const std::string& body) {
        const size_t padded = (body.size() + 7) & ~(size_t)7;
        put((uint64_t)type, 2); put(padded, 2); put(0, 1); put(0, 3);
        buf += body;
        buf.append(padded - body.size(), '\0');
    }
    uint64_t dataset(const H5Blob& b) {
        const uint64_t raw = begin();
        buf.append(reinterpret_cast<const char*>(b.data.data()), b.data.size() * sizeof(float));
        std::string space, dtype, layout;
        auto app = [](std::string& s, uint64_t v, int n) { for (int i = 0; i < n; i++) s.push_back((char)((v >> (8 * i)) & 0xff)); };
        app(space, 1, 1); app(space, b.dims.size(), 1); app(space, 0, 2); app(space, 0, 4);
        for (int d : b.dims) app(space, (uint64_t)d, 8);
        // IEEE float32 little endian: class 1 / version 1; bit field 0x20 0x1f 0x00; size 4; offset 0, precision 32, exponent 23/8,
        // mantissa 0/23, bias 127
        app(dtype, 0x11, 1); app(dtype, 0x20, 1); app(dtype, 0x1f, 1); app(dtype, 0, 1); app(dtype, 4, 4);
        app(dtype, 0, 2); app(dtype, 32, 2); app(dtype, 23, 1); app(dtype, 8, 1); app(dtype, 0, 1); app(dtype, 23, 1); app(dtype, 127, 4);
        app(layout, 3, 1); app(layout, 1, 1); app(layout, raw, 8); app(layout, b.data.size() * sizeof(float), 8);
        const uint64_t hdr = begin();
        put(1, 1); put(0, 1); put(3, 2); put(1, 4);
        const size_t size_at = buf.size();
        put(0, 4); put(0, 4);
        const size_t start = buf.size();
        message(1, space); message(3, dtype); message(8, layout);
        set(size_at, buf.size() - start, 4);
        return hdr;
    }
    // one group: local heap, symbol nodes of <= 8 links under ONE level-0 B-tree node (<= 32 children)
    void group(const std::vector<std::pair<std::string, uint64_t> >& sorted_links, uint64_t* hdr, uint64_t* btree, uint64_t* heap) {
        if (sorted_links.size() > 256) throw H5Error("more than 256 links in one group are not supported by the writer");
        std::string heap_data(8, '\0');
        std::vector<uint64_t> offs;
        for (const auto& l : sorted_links) {
            offs.push_back(heap_data.size());
            heap_data += l.first;
            heap_data.push_back('\0');
            while (heap_data.size() % 8) heap_data.push_back('\0');
        }
        const uint64_t data_addr = begin();
        buf += heap_data;
        *heap = begin();
        buf += "HEAP"; put(0, 1); put(0, 3); put(heap_data.size(), 8); put(kUndef, 8); put(data_addr, 8);
        std::vector<std::pair<uint64_t, uint64_t> > children;       // (symbol node address, heap offset of its largest name)
        for (size_t i = 0; i < sorted_links.size() || i == 0; i += 8) {
            const size_t n = sorted_links.size() > i ? std::min<size_t>(8, sorted_links.size() - i) : 0;
            const uint64_t snod = begin();
            buf += "SNOD"; put(1, 1); put(0, 1); put(n, 2);
            for (size_t k = 0; k < n; k++) { put(offs[i + k], 8); put(sorted_links[i + k].second, 8); put(0, 4); put(0, 4); buf.append(16, '\0'); }
            buf.append(40 * (8 - n), '\0');
            children.push_back({snod, n ? offs[i + n - 1] : 0});
            if (!n) break;
        }
        *btree = begin();
        buf += "TREE"; put(0, 1); put(0, 1); put(children.size(), 2); put(kUndef, 8); put(kUndef, 8);
        put(0, 8);
        for (const auto& c : children) { put(c.first, 8); put(c.second, 8); }
        buf.append(16 * (32 - children.size()), '\0');
        *hdr = begin();
        put(1, 1); put(0, 1); put(1, 2); put(1, 4); put(24, 4); put(0, 4);
        std::string body;
        for (int i = 0; i < 8; i++) body.push_back((char)((*btree >> (8 * i)) & 0xff));
        for (int i = 0; i < 8; i++) body.push_back((char)((*heap >> (8 * i)) & 0xff));
        message(0x11, body);
    }
};
}  // namespace

std::string WriteCaffemodelH5(const std::vector<std::pair<std::string, std::vector<H5Blob> > >& layers) {
    H5Out w;
    std::map<std::string, uint64_t> layer_groups;            // std::map: links sorted by name, as the B-tree requires
    for (const auto& l : layers) {
        std::map<std::string, uint64_t> links;
        for (size_t i = 0; i < l.second.size(); i++) links[std::to_string(i)] = w.dataset(l.second[i]);
        uint64_t hdr, bt, hp;
        w.group(std::vector<std::pair<std::string, uint64_t> >(links.begin(), links.end()), &hdr, &bt, &hp);
        if (layer_groups.count(l.first)) throw H5Error("duplicate layer name '" + l.first + "'");
        layer_groups[l.first] = hdr;
    }
    uint64_t data_hdr, bt, hp;
    w.group(std::vector<std::pair<std::string, uint64_t> >(layer_groups.begin(), layer_groups.end()), &data_hdr, &bt, &hp);
    uint64_t root_hdr, root_bt, root_hp;
    w.group({{"data", data_hdr}}, &root_hdr, &root_bt, &root_hp);
    static const unsigned char sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    for (int i = 0; i < 8; i++) w.buf[i] = (char)sig[i];
    w.buf[13] = 8; w.buf[14] = 8;                            // sizes of offsets / lengths; all version bytes 0
    w.set(16, 4, 2); w.set(18, 16, 2);                       // group leaf / internal node K
    w.set(24, 0, 8); w.set(32, kUndef, 8); w.set(40, w.buf.size(), 8); w.set(48, kUndef, 8);
    w.set(56, 0, 8); w.set(64, root_hdr, 8); w.set(72, 1, 4); w.set(80, root_bt, 8); w.set(88, root_hp, 8);
    return w.buf;
}

}  // namespace caffe
