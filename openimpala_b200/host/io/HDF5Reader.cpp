// Minimal HDF5 (file format spec 1.x objects) reader; semantics of the
// reference reader: dims (Z,Y,X) -> box (X,Y,Z) (src/io/HDF5Reader.cpp:136-153),
// native-type dispatch u8/i8/u16/i16/u32/i32/u64/i64/f32/f64 (:359-383),
// threshold rule double(v) > t ? a : b (:321-326).
// Layouts: contiguous, and chunked through the version-1 B-tree chunk index with the
// deflate (1), shuffle (2) and fletcher32 (3) filters -- what libhdf5 resolves behind
// DataSet::read for the reference (:255-402).  Not handled: compact layout, new-style
// groups / object headers (libver "latest"), szip / nbit / scale-offset filters, non-zero
// fill values of unallocated chunks.
#include "HDF5Reader.H"

#include <algorithm>
#include <cstdlib>
#include <climits>
#include <cstring>
#include <fstream>
#include <functional>
#include <sstream>
#include <stdexcept>
#include <thread>
#include <vector>

#include <zlib.h>

#include <AMReX_Utility.H>

namespace OpenImpala {
namespace {

class H5File {
public:
    explicit H5File(const std::string& path) : in(path, std::ios::binary) {
        if (!in) throw std::runtime_error("cannot open HDF5 file: " + path);
        static const unsigned char sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
        uint64_t base = 0;
        bool found = false;
        for (int t = 0; t < 8 && !found; ++t) {          // superblock at 0, 512, 1024, ...
            unsigned char b[8];
            if (!readAt(base, b, 8)) break;
            if (std::memcmp(b, sig, 8) == 0) found = true; else base = base ? base * 2 : 512;
        }
        if (!found) throw std::runtime_error("not an HDF5 file: " + path);
        unsigned char h[24];
        readAt(base + 8, h, 24);
        const int version = h[0];
        if (version > 1)
            throw std::runtime_error("HDF5 superblock version " + std::to_string(version) +
                                     " (latest-format file) needs libhdf5, which this build does not have");
        so = h[5]; sl = h[6];
        uint64_t p = base + 8 + 8 + 2 + 2 + 4 + (version == 1 ? 4 : 0);
        base_addr = rd(p, so); p += 4 * (uint64_t)so;      // base, free-space, eof, driver
        // root group symbol table entry
        p += so;                                            // link name offset
        root_header = rd(p, so);
    }
    uint64_t rootHeader() const { return root_header; }
    int sizeOffsets() const { return so; }
    int sizeLengths() const { return sl; }
    bool readAt(uint64_t off, void* dst, size_t n) {
        in.clear();
        in.seekg((std::streamoff)off);
        in.read(reinterpret_cast<char*>(dst), (std::streamsize)n);
        return (size_t)in.gcount() == n;
    }
    uint64_t rd(uint64_t off, int n) {
        unsigned char b[8] = {0};
        if (!readAt(off, b, (size_t)n)) throw std::runtime_error("HDF5: read past end of file");
        uint64_t v = 0;
        for (int q = 0; q < n; ++q) v |= (uint64_t)b[q] << (8 * q);
        return v;
    }
    struct Message { int type; std::vector<unsigned char> data; };
    // all messages of a version-1 object header, following continuation blocks
    std::vector<Message> messages(uint64_t addr) {
        addr += base_addr;
        unsigned char h[16];
        if (!readAt(addr, h, 16)) throw std::runtime_error("HDF5: bad object header address");
        if (h[0] != 1) throw std::runtime_error("HDF5: object header version " + std::to_string((int)h[0]) + " not supported without libhdf5");
        const int nmsg = h[2] | (h[3] << 8);
        const uint32_t hsize = h[8] | (h[9] << 8) | (h[10] << 16) | ((uint32_t)h[11] << 24);
        std::vector<Message> out;
        std::vector<std::pair<uint64_t, uint64_t>> blocks{{addr + 16, hsize}};
        for (size_t bi = 0; bi < blocks.size() && (int)out.size() < nmsg; ++bi) {
            uint64_t p = blocks[bi].first;
            const uint64_t end = p + blocks[bi].second;
            while (p + 8 <= end && (int)out.size() < nmsg) {
                unsigned char mh[8];
                if (!readAt(p, mh, 8)) break;
                Message m;
                m.type = mh[0] | (mh[1] << 8);
                const size_t sz = mh[2] | (mh[3] << 8);
                m.data.resize(sz);
                if (sz && !readAt(p + 8, m.data.data(), sz)) break;
                p += 8 + sz;
                if (m.type == 0x0010 && sz >= (size_t)(so + sl)) {          // continuation
                    uint64_t o = 0, l = 0;
                    for (int q = 0; q < so; ++q) o |= (uint64_t)m.data[q] << (8 * q);
                    for (int q = 0; q < sl; ++q) l |= (uint64_t)m.data[so + q] << (8 * q);
                    blocks.emplace_back(o + base_addr, l);
                }
                out.push_back(std::move(m));
            }
        }
        return out;
    }
    // object header address of `name` inside the old-style group whose header is at `group`
    uint64_t lookup(uint64_t group, const std::string& name) {
        uint64_t btree = 0, heap = 0;
        bool have = false;
        for (const Message& m : messages(group))
            if (m.type == 0x0011 && m.data.size() >= (size_t)(2 * so)) {
                for (int q = 0; q < so; ++q) btree |= (uint64_t)m.data[q] << (8 * q);
                for (int q = 0; q < so; ++q) heap |= (uint64_t)m.data[so + q] << (8 * q);
                have = true;
            }
        if (!have) throw std::runtime_error("HDF5: group without a symbol table (new-style groups need libhdf5)");
        heap += base_addr;
        char sig[4];
        readAt(heap, sig, 4);
        if (std::memcmp(sig, "HEAP", 4) != 0) throw std::runtime_error("HDF5: bad local heap");
        const uint64_t heap_data = rd(heap + 8 + 2 * (uint64_t)sl, so) + base_addr;
        return walk(btree + base_addr, heap_data, name);
    }
    uint64_t baseAddr() const { return base_addr; }
    uint64_t fileSize() {
        in.clear();
        in.seekg(0, std::ios::end);
        const std::streamoff e = in.tellg();
        return e > 0 ? (uint64_t)e : 0;
    }
    // every chunk below a version-1 B-tree node of type 1 (raw data chunks); `ndims` is the
    // dataset rank + 1 (the element-size dimension has an offset too)
    template <class Sink>
    void walkChunks(uint64_t node, int ndims, Sink&& sink, int depth = 0) {
        char sig[4];
        if (depth > 64 || !readAt(node, sig, 4) || std::memcmp(sig, "TREE", 4) != 0)
            throw std::runtime_error("HDF5: bad chunk B-tree node");
        unsigned char h[4];
        readAt(node + 4, h, 4);
        if (h[0] != 1) throw std::runtime_error("HDF5: chunk index is not a raw-data B-tree");
        const int level = h[1], used = h[2] | (h[3] << 8);
        const uint64_t key_bytes = 8 + 8 * (uint64_t)ndims;
        uint64_t p = node + 8 + 2 * (uint64_t)so;
        for (int e = 0; e < used; ++e) {
            const uint32_t bytes = (uint32_t)rd(p, 4), mask = (uint32_t)rd(p + 4, 4);
            uint64_t off[8] = {0};
            for (int d = 0; d < ndims && d < 8; ++d) off[d] = rd(p + 8 + 8 * (uint64_t)d, 8);
            const uint64_t child = rd(p + key_bytes, so) + base_addr;
            if (level == 0) sink(child, bytes, mask, off);
            else walkChunks(child, ndims, sink, depth + 1);
            p += key_bytes + (uint64_t)so;
        }
    }
private:
    uint64_t walk(uint64_t node, uint64_t heap_data, const std::string& name) {
        char sig[4];
        readAt(node, sig, 4);
        if (std::memcmp(sig, "TREE", 4) == 0) {
            unsigned char h[4];
            readAt(node + 4, h, 4);
            const int level = h[1], used = h[2] | (h[3] << 8);
            uint64_t p = node + 8 + 2 * (uint64_t)so;
            for (int e = 0; e < used; ++e) {
                p += sl;                                    // key e
                const uint64_t child = rd(p, so) + base_addr;
                p += so;
                const uint64_t r = walk(child, heap_data, name);
                if (r != UINT64_MAX) return r;
                (void)level;
            }
            return UINT64_MAX;
        }
        if (std::memcmp(sig, "SNOD", 4) == 0) {
            unsigned char h[4];
            readAt(node + 4, h, 4);
            const int nsym = h[2] | (h[3] << 8);
            uint64_t p = node + 8;
            for (int s = 0; s < nsym; ++s) {
                const uint64_t name_off = rd(p, so), hdr = rd(p + so, so);
                std::string nm;
                for (uint64_t q = heap_data + name_off;; ++q) {
                    char ch = 0;
                    if (!readAt(q, &ch, 1) || ch == 0) break;
                    nm += ch;
                }
                if (nm == name) return hdr;
                p += 2 * (uint64_t)so + 4 + 4 + 16;
            }
            return UINT64_MAX;
        }
        throw std::runtime_error("HDF5: unexpected node signature in group B-tree");
    }
    std::ifstream in;
    int so = 8, sl = 8;
    uint64_t base_addr = 0, root_header = 0;
};

}  // namespace

HDF5Reader::HDF5Reader() = default;

HDF5Reader::HDF5Reader(const std::string& filename, const std::string& hdf5dataset) {
    if (!readFile(filename, hdf5dataset))
        throw std::runtime_error("HDF5Reader: failed to read metadata of " + filename + ":" + hdf5dataset);
}

bool HDF5Reader::readFile(const std::string& filename, const std::string& hdf5dataset) {
    m_filename = filename; m_hdf5dataset = hdf5dataset; m_is_read = false;
    return readMetadataInternal();
}

bool HDF5Reader::readMetadataInternal() {
    try {
        H5File f(m_filename);
        uint64_t obj = f.rootHeader();
        std::stringstream path(m_hdf5dataset);
        std::string part;
        while (std::getline(path, part, '/')) {
            if (part.empty()) continue;
            obj = f.lookup(obj, part);
            if (obj == UINT64_MAX) throw std::runtime_error("dataset '" + m_hdf5dataset + "' not found");
        }
        const int so = f.sizeOffsets(), sl = f.sizeLengths();
        bool have_space = false, have_type = false, have_layout = false;
        uint64_t chunk_btree = UINT64_MAX;
        m_chunked = false;
        m_filters.clear();
        m_chunks.clear();
        for (const auto& m : f.messages(obj)) {
            const unsigned char* d = m.data.data();
            if (m.type == 0x0001 && m.data.size() >= 8) {                 // dataspace
                const int version = d[0], rank = d[1];
                if (rank != 3) throw std::runtime_error("dataset rank " + std::to_string(rank) + " != 3");
                const size_t off = version == 1 ? 8 : 4;
                if (m.data.size() < off + 3 * (size_t)sl) throw std::runtime_error("truncated dataspace message");
                uint64_t dims[3] = {0, 0, 0};
                for (int r = 0; r < 3; ++r)
                    for (int q = 0; q < sl; ++q) dims[r] |= (uint64_t)d[off + r * sl + q] << (8 * q);
                // the reference rejects extents that do not fit an int (src/io/HDF5Reader.cpp:142)
                for (int r = 0; r < 3; ++r)
                    if (dims[r] == 0 || dims[r] > (uint64_t)INT_MAX) throw std::runtime_error("dataset extent outside (0, INT_MAX]");
                m_depth = (int)dims[0]; m_height = (int)dims[1]; m_width = (int)dims[2];
                have_space = true;
            } else if (m.type == 0x0003 && m.data.size() >= 8) {          // datatype
                m_type_class = d[0] & 0x0f;
                m_type_big_endian = (d[1] & 0x01) != 0;
                m_type_signed = (d[1] & 0x08) != 0;
                m_type_size = (int)(d[4] | (d[5] << 8) | (d[6] << 16) | ((uint32_t)d[7] << 24));
                if (m_type_class > 1) throw std::runtime_error("unsupported HDF5 datatype class " + std::to_string(m_type_class));
                have_type = true;
            } else if (m.type == 0x0008 && m.data.size() >= 2) {          // data layout
                const int version = d[0];
                if (version == 3 && d[1] == 2) {                          // chunked: rank+1, B-tree address, chunk dims
                    const int ndims = d[2];
                    if (ndims != 4 || m.data.size() < (size_t)(3 + so + 4 * ndims))
                        throw std::runtime_error("chunked layout of a dataset that is not rank 3");
                    m_chunked = true;
                    chunk_btree = 0;
                    for (int q = 0; q < so; ++q) chunk_btree |= (uint64_t)d[3 + q] << (8 * q);
                    for (int r = 0; r < 3; ++r) {
                        const unsigned char* c = d + 3 + so + 4 * r;
                        m_chunk_dims[r] = (uint64_t)c[0] | ((uint64_t)c[1] << 8) | ((uint64_t)c[2] << 16) | ((uint64_t)c[3] << 24);
                        if (m_chunk_dims[r] == 0) throw std::runtime_error("zero chunk extent");
                    }
                    // a chunk larger than 4 Gi elements is implausible (HDF5 itself caps a chunk at 4 GiB)
                    // and would let the extent product wrap
                    if (m_chunk_dims[0] * m_chunk_dims[1] > ((uint64_t)1 << 32) ||
                        m_chunk_dims[0] * m_chunk_dims[1] * m_chunk_dims[2] > ((uint64_t)1 << 32))
                        throw std::runtime_error("implausible chunk extents");
                    m_data_offset = 0;
                } else if (version == 3) {
                    if (d[1] != 1) throw std::runtime_error("compact dataset layout is not supported without libhdf5");
                    m_data_offset = 0;
                    for (int q = 0; q < so; ++q) m_data_offset |= (uint64_t)d[2 + q] << (8 * q);
                } else if (version == 1 || version == 2) {
                    if (d[2] != 1) throw std::runtime_error("only contiguous dataset layout is supported without libhdf5");
                    m_data_offset = 0;
                    for (int q = 0; q < so; ++q) m_data_offset |= (uint64_t)d[8 + q] << (8 * q);
                } else throw std::runtime_error("unsupported layout message version");
                m_data_offset += f.baseAddr();
                have_layout = true;
            } else if (m.type == 0x000B && m.data.size() >= 2) {         // filter pipeline, versions 1 and 2
                const int version = d[0], nf = d[1];
                size_t p = version == 1 ? 8 : 2;
                for (int q = 0; q < nf; ++q) {
                    if (p + (version == 1 ? 8 : 6) > m.data.size()) throw std::runtime_error("truncated filter pipeline message");
                    Filter flt;
                    flt.id = d[p] | (d[p + 1] << 8);
                    p += 2;
                    size_t name_len = 0;
                    if (version == 1 || flt.id >= 256) { name_len = d[p] | (d[p + 1] << 8); p += 2; }
                    p += 2;                                               // flags
                    const int ncl = d[p] | (d[p + 1] << 8);
                    p += 2;
                    p += version == 1 ? (name_len + 7) / 8 * 8 : name_len;
                    if (p + 4 * (size_t)ncl > m.data.size()) throw std::runtime_error("truncated filter pipeline message");
                    for (int c = 0; c < ncl; ++c, p += 4)
                        flt.client.push_back((uint32_t)d[p] | ((uint32_t)d[p + 1] << 8) | ((uint32_t)d[p + 2] << 16) | ((uint32_t)d[p + 3] << 24));
                    if (version == 1 && (ncl & 1)) p += 4;
                    if (flt.id < 1 || flt.id > 3)
                        throw std::runtime_error("HDF5 filter " + std::to_string(flt.id) + " (only deflate, shuffle and fletcher32 are built in)");
                    m_filters.push_back(flt);
                }
            }
        }
        if (!(have_space && have_type && have_layout)) throw std::runtime_error("incomplete dataset header");
        if (m_width <= 0 || m_height <= 0 || m_depth <= 0) throw std::runtime_error("bad dataset dimensions");
        if (!m_chunked && !m_filters.empty()) throw std::runtime_error("filters on a dataset that is not chunked");
        if (m_type_size <= 0 || m_type_size > 8) throw std::runtime_error("unsupported element size");
        {   // a contiguous dataset lies inside the file; reject headers whose extents cannot be true
            const uint64_t fsz = f.fileSize();
            const long double bytes = (long double)m_width * (long double)m_height * (long double)m_depth * (long double)m_type_size;
            if (!m_chunked && (m_data_offset > fsz || bytes > (long double)(fsz - m_data_offset)))
                throw std::runtime_error("dataset extents exceed the file size");
            if (bytes > 1.0e15L) throw std::runtime_error("implausible dataset extents");
        }
        const uint64_t undef_addr = so >= 8 ? UINT64_MAX : ((1ull << (8 * so)) - 1);
        if (m_chunked && chunk_btree != UINT64_MAX && chunk_btree != undef_addr) {   // undefined address: no chunk was ever written
            f.walkChunks(chunk_btree + f.baseAddr(), 4, [&](uint64_t addr, uint32_t bytes, uint32_t mask, const uint64_t* off) {
                if (off[0] >= (uint64_t)m_depth || off[1] >= (uint64_t)m_height || off[2] >= (uint64_t)m_width) return;
                m_chunks.push_back(Chunk{addr, bytes, mask, {off[0], off[1], off[2]}});
            });
        }
    } catch (const std::exception& e) {
        amrex::Warning(std::string("[HDF5Reader] ") + e.what());
        return false;
    }
    m_is_read = true;
    return true;
}

amrex::Box HDF5Reader::box() const {
    if (!m_is_read) return amrex::Box();
    return amrex::Box(amrex::IntVect::TheZeroVector(), amrex::IntVect(m_width - 1, m_height - 1, m_depth - 1));
}

std::string HDF5Reader::getAttribute(const std::string& attr_name) const {
    if (!m_is_read) return "<Attribute read error: Metadata not read>";
    return "<HDF5 Error reading attr '" + attr_name + "'>";   // attributes need libhdf5
}

// out[((k - z_begin) * H + j) * W + i] = (double(v) > t) ? vt : vf (reference rule,
// src/io/HDF5Reader.cpp:321-326) for planes [z_begin, z_begin + nz) of the contiguous dataset:
// whole planes are read at once, 8- and 16-bit integers go through a lookup table of the rule,
// planes are spread over up to 16 threads (OI_IO_THREADS), each with its own file handle.
template <class OutT>
void HDF5Reader::thresholdInto(double t, OutT vt, OutT vf, int z_begin, int nz, OutT* out) const {
    const size_t bps = (size_t)m_type_size;
    const size_t plane = (size_t)m_width * (size_t)m_height;
    std::vector<OutT> lut;
    if (m_type_class != 1 && (bps == 1 || bps == 2)) {
        lut.resize((size_t)1 << (8 * bps));
        for (size_t v = 0; v < lut.size(); ++v) {
            const double sv = !m_type_signed ? (double)v : (bps == 1 ? (double)(int8_t)(uint8_t)v : (double)(int16_t)(uint16_t)v);
            lut[v] = (sv > t) ? vt : vf;
        }
    }
    // n samples at src (file byte order) -> thresholded values
    auto convert = [&](const unsigned char* src, size_t n, OutT* dst) {
        if (!lut.empty() && bps == 1) {
            for (size_t q = 0; q < n; ++q) dst[q] = lut[src[q]];
        } else if (!lut.empty()) {
            if (!m_type_big_endian) for (size_t q = 0; q < n; ++q) dst[q] = lut[(size_t)src[2 * q] | ((size_t)src[2 * q + 1] << 8)];
            else for (size_t q = 0; q < n; ++q) dst[q] = lut[((size_t)src[2 * q] << 8) | (size_t)src[2 * q + 1]];
        } else {
            for (size_t i = 0; i < n; ++i) {
                unsigned char b[8] = {0};
                const unsigned char* p = src + i * bps;
                for (size_t q = 0; q < bps && q < 8; ++q) b[q] = m_type_big_endian ? p[bps - 1 - q] : p[q];
                double v = 0.0;
                if (m_type_class == 1) {
                    if (bps == 4) { float x; std::memcpy(&x, b, 4); v = (double)x; }
                    else if (bps == 8) { std::memcpy(&v, b, 8); }
                } else if (m_type_signed) {
                    if (bps == 4) { int32_t x; std::memcpy(&x, b, 4); v = (double)x; }
                    else if (bps == 8) { int64_t x; std::memcpy(&x, b, 8); v = (double)x; }
                } else {
                    uint64_t x = 0;
                    std::memcpy(&x, b, bps > 8 ? 8 : bps);
                    v = (double)x;
                }
                dst[i] = (v > t) ? vt : vf;
            }
        }
    };
    int T = (int)std::thread::hardware_concurrency();
    if (T <= 0) T = 1;
    T = std::min(T, 16);
    if (const char* e = std::getenv("OI_IO_THREADS")) T = std::atoi(e);
    auto run_threads = [&](int n_threads, const std::function<void(int)>& work) {
        if (n_threads <= 1) { work(0); return; }
        std::vector<std::thread> pool;
        for (int w = 0; w < n_threads; ++w) pool.emplace_back(work, w);
        for (auto& th : pool) th.join();
    };

    if (m_chunked) {
        // unallocated chunks read as the fill value (0); allocated ones are read, run backwards
        // through the filter pipeline and thresholded into their part of the output
        const OutT fill = (0.0 > t) ? vt : vf;
        std::fill(out, out + (size_t)nz * plane, fill);
        std::vector<const Chunk*> todo;
        for (const Chunk& c : m_chunks)
            if (c.offset[0] < (uint64_t)(z_begin + nz) && c.offset[0] + m_chunk_dims[0] > (uint64_t)z_begin) todo.push_back(&c);
        const size_t chunk_elems = (size_t)(m_chunk_dims[0] * m_chunk_dims[1] * m_chunk_dims[2]);
        const size_t chunk_bytes = chunk_elems * bps;
        const int TC = std::max(1, std::min<int>(T, (int)todo.size()));
        std::vector<std::string> errors((size_t)TC);
        run_threads(TC, [&](int w) {
            try {
                std::ifstream in(m_filename, std::ios::binary);
                if (!in) throw std::runtime_error("cannot reopen " + m_filename);
                std::vector<unsigned char> raw, data(chunk_bytes), tmp;
                for (size_t ci = (size_t)w; ci < todo.size(); ci += (size_t)TC) {
                    const Chunk& c = *todo[ci];
                    raw.resize(c.bytes);
                    in.clear();
                    in.seekg((std::streamoff)c.address);
                    in.read(reinterpret_cast<char*>(raw.data()), (std::streamsize)raw.size());
                    if ((size_t)in.gcount() != raw.size()) throw std::runtime_error("chunk past the end of the file");
                    // undo the pipeline, last filter first; bit q of the mask: filter q was skipped for this chunk
                    for (int q = (int)m_filters.size() - 1; q >= 0; --q) {
                        if (c.filter_mask & (1u << q)) continue;
                        const Filter& flt = m_filters[(size_t)q];
                        if (flt.id == 3) {                               // fletcher32: checksum appended
                            if (raw.size() < 4) throw std::runtime_error("fletcher32 chunk shorter than its checksum");
                            raw.resize(raw.size() - 4);
                        } else if (flt.id == 1) {                        // deflate
                            tmp.resize(chunk_bytes + 4);                 // (+4: a checksum filter may sit before deflate)
                            uLongf len = (uLongf)tmp.size();
                            if (uncompress(tmp.data(), &len, raw.data(), (uLong)raw.size()) != Z_OK)
                                throw std::runtime_error("corrupt deflate stream in a chunk");
                            tmp.resize((size_t)len);
                            raw.swap(tmp);
                        } else if (flt.id == 2) {                        // shuffle: byte planes of the elements
                            const size_t es = flt.client.empty() ? bps : (size_t)flt.client[0];
                            if (es > 1) {
                                const size_t ne = raw.size() / es;
                                tmp.resize(raw.size());
                                for (size_t bq = 0; bq < es; ++bq)
                                    for (size_t i = 0; i < ne; ++i) tmp[i * es + bq] = raw[bq * ne + i];
                                for (size_t r = ne * es; r < raw.size(); ++r) tmp[r] = raw[r];
                                raw.swap(tmp);
                            }
                        }
                    }
                    if (raw.size() < chunk_bytes) throw std::runtime_error("chunk holds fewer bytes than its extents");
                    // the part of the chunk inside the dataset and inside [z_begin, z_begin + nz)
                    const uint64_t z0 = std::max<uint64_t>(c.offset[0], (uint64_t)z_begin);
                    const uint64_t z1 = std::min<uint64_t>({c.offset[0] + m_chunk_dims[0], (uint64_t)m_depth, (uint64_t)(z_begin + nz)});
                    const uint64_t y1 = std::min<uint64_t>(c.offset[1] + m_chunk_dims[1], (uint64_t)m_height);
                    const uint64_t x1 = std::min<uint64_t>(c.offset[2] + m_chunk_dims[2], (uint64_t)m_width);
                    for (uint64_t z = z0; z < z1; ++z)
                        for (uint64_t y = c.offset[1]; y < y1; ++y) {
                            const size_t src_elem = (size_t)(((z - c.offset[0]) * m_chunk_dims[1] + (y - c.offset[1])) * m_chunk_dims[2]);
                            OutT* dst = out + ((size_t)(z - (uint64_t)z_begin) * (size_t)m_height + (size_t)y) * (size_t)m_width + (size_t)c.offset[2];
                            convert(raw.data() + src_elem * bps, (size_t)(x1 - c.offset[2]), dst);
                        }
                }
            } catch (const std::exception& e) {
                errors[(size_t)w] = e.what();
            }
        });
        for (const std::string& e : errors)
            if (!e.empty()) amrex::Abort("[HDF5Reader::threshold] " + e);
        return;
    }

    T = std::max(1, std::min(T, nz));
    std::vector<std::string> errors((size_t)T);
    run_threads(T, [&](int w) {
        const int lo = z_begin + (int)((long long)nz * w / T), hi = z_begin + (int)((long long)nz * (w + 1) / T);
        if (lo >= hi) return;
        std::ifstream in(m_filename, std::ios::binary);
        if (!in) { errors[(size_t)w] = "cannot reopen " + m_filename; return; }
        std::vector<unsigned char> buf(plane * bps);
        for (int k = lo; k < hi; ++k) {
            in.clear();
            in.seekg((std::streamoff)(m_data_offset + (uint64_t)k * plane * bps));
            in.read(reinterpret_cast<char*>(buf.data()), (std::streamsize)buf.size());
            if ((size_t)in.gcount() != buf.size()) std::memset(buf.data() + in.gcount(), 0, buf.size() - (size_t)in.gcount());
            convert(buf.data(), plane, out + (size_t)(k - z_begin) * plane);
        }
    });
    for (const std::string& e : errors)
        if (!e.empty()) amrex::Abort("[HDF5Reader::threshold] " + e);
}

void HDF5Reader::threshold(double t, int value_if_true, int value_if_false, amrex::iMultiFab& mf) const {
    if (!m_is_read) amrex::Abort("[HDF5Reader::threshold] metadata not read");
    AMREX_ALWAYS_ASSERT_WITH_MESSAGE(mf.boxArray().minimalBox() == this->box(), "HDF5Reader: iMultiFab domain mismatch");
    // one rank: the whole dataset; z-slabs: this rank's planes (full rows and columns)
    const amrex::Box vb = mf.validBox(), full = this->box();
    AMREX_ALWAYS_ASSERT_WITH_MESSAGE(mf.nGrow() == 0 && vb.smallEnd(0) == full.smallEnd(0) && vb.bigEnd(0) == full.bigEnd(0) &&
                                     vb.smallEnd(1) == full.smallEnd(1) && vb.bigEnd(1) == full.bigEnd(1) &&
                                     vb.smallEnd(2) >= full.smallEnd(2) && vb.bigEnd(2) <= full.bigEnd(2),
                                     "HDF5Reader: the destination must be a ghost-free field over the dataset box (or a z-slab of it)");
    thresholdInto<int>(t, value_if_true, value_if_false, vb.smallEnd(2), vb.length(2), &mf(vb.smallEnd(0), vb.smallEnd(1), vb.smallEnd(2)));
}

void HDF5Reader::thresholdPlanesU8(double t, unsigned char value_if_true, unsigned char value_if_false, int z_begin,
                                   int nz, unsigned char* out) const {
    if (!m_is_read) amrex::Abort("[HDF5Reader::thresholdPlanesU8] metadata not read");
    if (z_begin < 0 || nz < 0 || z_begin + nz > m_depth) amrex::Abort("[HDF5Reader::thresholdPlanesU8] plane range outside the dataset");
    thresholdInto<unsigned char>(t, value_if_true, value_if_false, z_begin, nz, out);
}

void HDF5Reader::threshold(double t, amrex::iMultiFab& mf) const { threshold(t, 1, 0, mf); }

}  // namespace OpenImpala
