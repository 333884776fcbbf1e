// Baseline-TIFF decode + threshold, behaviour of src/io/TiffReader.cpp:289-444:
//   * 1-bit rows are unpacked with the scanline pitch ceil(W/8) for strips
//     (:419-426) and linearly with the tile width for tiles (:376-381);
//   * libtiff hands back MSB-first bytes even for FillOrder=2 files and the
//     reference then reads those LSB-first (:380, :425), i.e. MSB-first bits of the
//     raw file byte -- reproduced here;
//   * wider samples are converted through double (:41-79), byte-swapped to host
//     order as libtiff would;
//   * PhotometricInterpretation is ignored (the reference never inverts);
//   * rule: (double(sample) > threshold) ? value_if_true : value_if_false (:434).
#include "TiffReader.H"

#include <cstring>
#include <fstream>
#include <iomanip>
#include <map>
#include <sstream>
#include <memory>
#include <set>
#include <stdexcept>
#include <thread>

#include <zlib.h>

#include <AMReX_Utility.H>

namespace OpenImpala {
namespace {

struct Ifd {
    std::map<int, std::vector<uint64_t>> tags;
    uint64_t get(int tag, uint64_t dflt) const {
        auto it = tags.find(tag);
        return (it == tags.end() || it->second.empty()) ? dflt : it->second[0];
    }
    const std::vector<uint64_t>* arr(int tag) const {
        auto it = tags.find(tag);
        return it == tags.end() ? nullptr : &it->second;
    }
};

class TiffFile {
public:
    explicit TiffFile(const std::string& path) : in(path, std::ios::binary) {
        if (!in) throw std::runtime_error("cannot open TIFF file: " + path);
        unsigned char h[16];
        in.read(reinterpret_cast<char*>(h), 16);
        if (in.gcount() < 8) throw std::runtime_error("not a TIFF file: " + path);
        if (h[0] == 'I' && h[1] == 'I') little = true;
        else if (h[0] == 'M' && h[1] == 'M') little = false;
        else throw std::runtime_error("not a TIFF file: " + path);
        in.seekg(0, std::ios::end);
        file_size = (uint64_t)std::max<std::streamoff>(0, in.tellg());
        const uint64_t magic = rd(h + 2, 2);
        if (magic == 42) { big = false; first = rd(h + 4, 4); }
        else if (magic == 43) { big = true; first = rd(h + 8, 8); }
        else throw std::runtime_error("bad TIFF magic: " + path);
    }
    std::vector<Ifd> directories() {
        std::vector<Ifd> out;
        uint64_t off = first;
        std::set<uint64_t> seen;                              // a corrupt next-IFD offset must not loop for ever
        while (off != 0) {
            if (off >= file_size || !seen.insert(off).second)
                throw std::runtime_error("corrupt TIFF directory chain");
            Ifd d;
            std::vector<unsigned char> cnt(big ? 8 : 2);
            readAt(off, cnt.data(), cnt.size());
            const uint64_t n = rd(cnt.data(), cnt.size());
            const size_t esz = big ? 20 : 12, vsz = big ? 8 : 4;
            if (n > 65535 || off + cnt.size() + n * esz > file_size) throw std::runtime_error("corrupt TIFF directory");
            std::vector<unsigned char> ent(n * esz + vsz);
            readAt(off + cnt.size(), ent.data(), ent.size());
            for (uint64_t e = 0; e < n; ++e) {
                const unsigned char* p = ent.data() + e * esz;
                const int tag = (int)rd(p, 2), type = (int)rd(p + 2, 2);
                const uint64_t count = rd(p + 4, vsz);
                const size_t tsz = typeSize(type);
                if (tsz == 0 || count == 0) continue;
                if (count > file_size) throw std::runtime_error("corrupt TIFF tag (count larger than the file)");
                std::vector<unsigned char> raw(tsz * count);
                if (raw.size() <= vsz) std::memcpy(raw.data(), p + 4 + vsz, raw.size());
                else readAt(rd(p + 4 + vsz, vsz), raw.data(), raw.size());
                const uint64_t keep = std::min<uint64_t>(count, (type == 2) ? 0 : count);
                std::vector<uint64_t> vals;
                vals.reserve(keep);
                for (uint64_t q = 0; q < keep; ++q) vals.push_back(rd(raw.data() + q * tsz, tsz > 8 ? 8 : tsz));
                d.tags[tag] = vals;
            }
            out.push_back(d);
            off = rd(ent.data() + n * esz, vsz);
        }
        return out;
    }
    void readAt(uint64_t off, unsigned char* dst, size_t n) {
        in.clear();
        in.seekg((std::streamoff)off);
        in.read(reinterpret_cast<char*>(dst), (std::streamsize)n);
        if ((size_t)in.gcount() != n) std::memset(dst + in.gcount(), 0, n - (size_t)in.gcount());
    }
    bool littleEndian() const { return little; }
private:
    uint64_t rd(const unsigned char* p, size_t n) const {
        uint64_t v = 0;
        if (little) for (size_t q = 0; q < n; ++q) v |= (uint64_t)p[q] << (8 * q);
        else for (size_t q = 0; q < n; ++q) v = (v << 8) | p[q];
        return v;
    }
    static size_t typeSize(int t) {
        switch (t) {
            case 1: case 2: case 6: case 7: return 1;
            case 3: case 8: return 2;
            case 4: case 9: case 11: case 13: return 4;
            case 5: case 10: case 12: case 16: case 17: case 18: return 8;
            default: return 0;
        }
    }
    std::ifstream in;
    bool little = true, big = false;
    uint64_t first = 0, file_size = 0;
};

enum { T_WIDTH = 256, T_HEIGHT = 257, T_BPS = 258, T_COMPRESSION = 259, T_FILLORDER = 266,
       T_STRIPOFFSETS = 273, T_SPP = 277, T_ROWSPERSTRIP = 278, T_STRIPBYTECOUNTS = 279,
       T_PLANAR = 284, T_PREDICTOR = 317, T_TILEWIDTH = 322, T_TILELENGTH = 323, T_TILEOFFSETS = 324,
       T_TILEBYTECOUNTS = 325, T_SAMPLEFORMAT = 339 };
enum { C_NONE = 1, C_LZW = 5, C_DEFLATE = 8, C_DEFLATE_OLD = 32946, C_PACKBITS = 32773 };

// ---- strip / tile decompression (the reference gets these from libtiff) ----------------
// TIFF 6.0 LZW: MSB-first codes of 9..12 bits, Clear = 256, EOI = 257, "early change".
std::vector<unsigned char> lzwDecode(const unsigned char* in, size_t n, size_t expect) {
    std::vector<unsigned char> out;
    out.reserve(expect);
    std::vector<int> prefix(4096, -1), length(4096, 1);
    std::vector<unsigned char> suffix(4096, 0), first(4096, 0);
    for (int i = 0; i < 256; ++i) { suffix[i] = first[i] = (unsigned char)i; }
    int next = 258, width = 9, old = -1;
    uint64_t bitbuf = 0;
    int bits = 0;
    size_t pos = 0;
    auto emit = [&](int c) {
        const int l = length[c];
        const size_t base = out.size();
        out.resize(base + (size_t)l);
        for (int q = l - 1, cur = c; q >= 0; --q) { out[base + (size_t)q] = suffix[cur]; cur = prefix[cur]; }
    };
    while (out.size() < expect) {
        while (bits < width && pos < n) { bitbuf = (bitbuf << 8) | in[pos++]; bits += 8; }
        if (bits < width) break;
        const int code = (int)((bitbuf >> (bits - width)) & ((1u << width) - 1u));
        bits -= width;
        if (code == 257) break;
        if (code == 256) { next = 258; width = 9; old = -1; continue; }
        if (old < 0) {
            if (code >= 256) throw std::runtime_error("corrupt LZW stream");
            emit(code);
        } else {
            if (code > next || next >= 4096) throw std::runtime_error("corrupt LZW stream");
            prefix[next] = old;
            length[next] = length[old] + 1;
            first[next] = first[old];
            suffix[next] = (code < next) ? first[code] : first[old];      // KwKwK when code == next
            ++next;
            emit(code);
        }
        old = code;
        if (next >= (1 << width) - 1 && width < 12) ++width;
    }
    return out;
}

std::vector<unsigned char> packBitsDecode(const unsigned char* in, size_t n, size_t expect) {
    std::vector<unsigned char> out;
    out.reserve(expect);
    size_t i = 0;
    while (i < n && out.size() < expect) {
        const int c = (int)(signed char)in[i++];
        if (c >= 0) {
            const size_t m = std::min<size_t>((size_t)c + 1, n - i);
            out.insert(out.end(), in + i, in + i + m);
            i += m;
        } else if (c != -128 && i < n) {
            out.insert(out.end(), (size_t)(1 - c), in[i++]);
        }
    }
    return out;
}

std::vector<unsigned char> inflateAll(const unsigned char* in, size_t n, size_t expect) {
    std::vector<unsigned char> out(expect);
    uLongf len = (uLongf)expect;
    const int rc = uncompress(out.data(), &len, in, (uLong)n);
    if (rc != Z_OK && rc != Z_BUF_ERROR) throw std::runtime_error("corrupt Deflate stream in TIFF");
    out.resize((size_t)len);
    return out;
}

// Decode one strip / tile in place: `buf` holds the file bytes on entry and the samples on exit.
void decodeSegment(std::vector<unsigned char>& buf, uint64_t compression, uint64_t predictor, size_t expect,
                   size_t row_samples, size_t bytes_per_sample, bool file_little) {
    if (compression == C_LZW) buf = lzwDecode(buf.data(), buf.size(), expect);
    else if (compression == C_PACKBITS) buf = packBitsDecode(buf.data(), buf.size(), expect);
    else if (compression == C_DEFLATE || compression == C_DEFLATE_OLD) buf = inflateAll(buf.data(), buf.size(), expect);
    if (predictor == 2 && row_samples > 0) {                        // horizontal differencing, integer samples
        const size_t row_bytes = row_samples * bytes_per_sample;
        for (size_t r0 = 0; r0 + row_bytes <= buf.size(); r0 += row_bytes) {
            unsigned char* row = buf.data() + r0;
            if (bytes_per_sample == 1) {
                for (size_t i = 1; i < row_samples; ++i) row[i] = (unsigned char)(row[i] + row[i - 1]);
            } else {
                uint64_t prev = 0;
                for (size_t i = 0; i < row_samples; ++i) {
                    uint64_t v = 0;
                    for (size_t q = 0; q < bytes_per_sample; ++q)
                        v |= (uint64_t)row[i * bytes_per_sample + (file_little ? q : bytes_per_sample - 1 - q)] << (8 * q);
                    v = (i == 0) ? v : v + prev;
                    prev = v;
                    for (size_t q = 0; q < bytes_per_sample; ++q)
                        row[i * bytes_per_sample + (file_little ? q : bytes_per_sample - 1 - q)] = (unsigned char)(v >> (8 * q));
                }
            }
        }
    }
}

double sampleAsDouble(const unsigned char* p, int bps, int fmt, bool file_little) {
    const int nb = bps / 8;
    unsigned char b[8] = {0};
    const bool host_little = true;   // x86-64 / aarch64
    for (int q = 0; q < nb; ++q) b[q] = (file_little == host_little) ? p[q] : p[nb - 1 - q];
    switch (fmt) {
        case 1:
            if (nb == 1) return (double)b[0];
            if (nb == 2) { uint16_t v; std::memcpy(&v, b, 2); return (double)v; }
            if (nb == 4) { uint32_t v; std::memcpy(&v, b, 4); return (double)v; }
            if (nb == 8) { uint64_t v; std::memcpy(&v, b, 8); return (double)v; }
            return 0.0;
        case 2:
            if (nb == 1) { int8_t v; std::memcpy(&v, b, 1); return (double)v; }
            if (nb == 2) { int16_t v; std::memcpy(&v, b, 2); return (double)v; }
            if (nb == 4) { int32_t v; std::memcpy(&v, b, 4); return (double)v; }
            if (nb == 8) { int64_t v; std::memcpy(&v, b, 8); return (double)v; }
            return 0.0;
        case 3:
            if (nb == 4) { float v; std::memcpy(&v, b, 4); return (double)v; }
            if (nb == 8) { double v; std::memcpy(&v, b, 8); return v; }
            return 0.0;
        default: return 0.0;
    }
}

std::string sequenceName(const std::string& base, int index, int digits, const std::string& suffix) {
    std::ostringstream ss;
    ss << base << std::setw(digits) << std::setfill('0') << index << suffix;
    return ss.str();
}

void checkSupported(const Ifd& d, const std::string& name, int& w, int& h, uint16_t& bps, uint16_t& fmt,
                    uint16_t& spp, uint16_t& fill) {
    w = (int)d.get(T_WIDTH, 0); h = (int)d.get(T_HEIGHT, 0);
    bps = (uint16_t)d.get(T_BPS, 1); fmt = (uint16_t)d.get(T_SAMPLEFORMAT, 1);
    spp = (uint16_t)d.get(T_SPP, 1); fill = (uint16_t)d.get(T_FILLORDER, 1);
    const uint64_t planar = d.get(T_PLANAR, 1);
    const bool valid_bps = (bps == 1 || bps == 8 || bps == 16 || bps == 32 || bps == 64);
    if (w <= 0 || h <= 0 || !valid_bps || planar != 1 || spp != 1) {
        std::stringstream ss;
        ss << "[TiffReader] Invalid/unsupported TIFF: " << name << " (W=" << w << ", H=" << h
           << ", BPS=" << bps << ", Planar=" << planar << ", SPP=" << spp << ").";
        amrex::Abort(ss.str());
    }
    const uint64_t comp = d.get(T_COMPRESSION, 1);
    if (comp != C_NONE && comp != C_LZW && comp != C_DEFLATE && comp != C_DEFLATE_OLD && comp != C_PACKBITS)
        amrex::Abort("[TiffReader] unsupported TIFF compression scheme " + std::to_string(comp) + " in: " + name +
                     " (supported: none, LZW, Deflate, PackBits)");
    const uint64_t pred = d.get(T_PREDICTOR, 1);
    if (pred != 1 && !(pred == 2 && bps >= 8 && fmt != 3))
        amrex::Abort("[TiffReader] unsupported TIFF predictor " + std::to_string(pred) + " in: " + name);
}

// Every directory of a stack (and every file of a sequence) is decoded with the metadata of the first
// one, so each is checked against it before its bytes are thresholded: the reference re-queries libtiff
// per directory (src/io/TiffReader.cpp:320-352) and a page of another size, bit depth or compression
// scheme must not be read as raw samples.  Throws (worker threads collect the message).
void verifyDirectory(const Ifd& d, const std::string& name, int W, int H, int bps, int fmt, int fill) {
    const uint64_t comp = d.get(T_COMPRESSION, 1), pred = d.get(T_PREDICTOR, 1);
    std::ostringstream why;
    if ((int)d.get(T_WIDTH, 0) != W || (int)d.get(T_HEIGHT, 0) != H)
        why << "size " << d.get(T_WIDTH, 0) << "x" << d.get(T_HEIGHT, 0) << " differs from the first page's " << W << "x" << H;
    else if ((int)d.get(T_BPS, 1) != bps || (int)d.get(T_SAMPLEFORMAT, 1) != fmt)
        why << "BitsPerSample/SampleFormat " << d.get(T_BPS, 1) << "/" << d.get(T_SAMPLEFORMAT, 1)
            << " differ from the first page's " << bps << "/" << fmt;
    else if ((int)d.get(T_FILLORDER, 1) != fill)
        why << "FillOrder " << d.get(T_FILLORDER, 1) << " differs from the first page's " << fill;
    else if (d.get(T_PLANAR, 1) != 1 || d.get(T_SPP, 1) != 1)
        why << "Planar=" << d.get(T_PLANAR, 1) << ", SPP=" << d.get(T_SPP, 1) << " unsupported";
    else if (comp != C_NONE && comp != C_LZW && comp != C_DEFLATE && comp != C_DEFLATE_OLD && comp != C_PACKBITS)
        why << "unsupported compression scheme " << comp;
    else if (pred != 1 && !(pred == 2 && bps >= 8 && fmt != 3))
        why << "unsupported predictor " << pred;
    if (!why.str().empty()) throw std::runtime_error("inconsistent TIFF directory in " + name + ": " + why.str());
}

}  // namespace

TiffReader::TiffReader() = default;

TiffReader::TiffReader(const std::string& filename) : TiffReader() {
    if (!readFile(filename)) throw std::runtime_error("TiffReader(filename): Failed to read metadata from file: " + filename);
}

TiffReader::TiffReader(const std::string& base_pattern, int num_files, int start_index, int digits,
                       const std::string& suffix) : TiffReader() {
    if (!readFileSequence(base_pattern, num_files, start_index, digits, suffix))
        throw std::runtime_error("TiffReader(sequence): Failed to read metadata for sequence: " + base_pattern);
}

amrex::Box TiffReader::box() const {
    if (!m_is_read) return amrex::Box();
    return amrex::Box(amrex::IntVect::TheZeroVector(), amrex::IntVect(m_width - 1, m_height - 1, m_depth - 1));
}

bool TiffReader::readFile(const std::string& filename) {
    m_is_sequence = false; m_filename = filename; m_base_pattern.clear();
    if (filename.empty()) amrex::Abort("[TiffReader::readFile] Filename cannot be empty.");
    try {
        TiffFile f(filename);
        const std::vector<Ifd> dirs = f.directories();
        if (dirs.empty()) amrex::Abort("[TiffReader::readFile] No directories (depth=0) in: " + filename);
        checkSupported(dirs[0], filename, m_width, m_height, m_bits_per_sample, m_sample_format,
                       m_samples_per_pixel, m_fill_order);
        m_depth = (int)dirs.size();
    } catch (const std::exception& e) {
        amrex::Abort(std::string("[TiffReader::readFile] ") + e.what());
    }
    m_is_read = true;
    return true;
}

bool TiffReader::readFileSequence(const std::string& base_pattern, int num_files, int start_index,
                                  int digits, const std::string& suffix) {
    m_is_sequence = true; m_base_pattern = base_pattern; m_start_index = start_index;
    m_digits = digits; m_suffix = suffix; m_filename.clear();
    if (num_files <= 0 || digits <= 0 || base_pattern.empty()) amrex::Abort("[TiffReader::readFileSequence] Invalid sequence params.");
    const std::string first = sequenceName(base_pattern, start_index, digits, suffix);
    try {
        TiffFile f(first);
        const std::vector<Ifd> dirs = f.directories();
        if (dirs.empty()) amrex::Abort("[TiffReader::readFileSequence] empty file: " + first);
        checkSupported(dirs[0], first, m_width, m_height, m_bits_per_sample, m_sample_format,
                       m_samples_per_pixel, m_fill_order);
        if (dirs.size() > 1) amrex::Warning("[TiffReader::readFileSequence] First sequence file has >1 directory. Using first only for metadata.");
    } catch (const std::exception& e) {
        amrex::Abort(std::string("[TiffReader::readFileSequence] ") + e.what());
    }
    m_depth = num_files;
    m_is_read = true;
    return true;
}

// ---- plane decode + threshold ------------------------------------------------------------
// One directory (= one z-plane) is handed out as row spans of decoded sample bytes: row j,
// columns [x0, x0 + n), `src` pointing at the first sample of the span (bit-packed data: at
// the byte holding bit `bit0` of the span's bit string), `avail` bytes readable from `src`.
// Samples past the decoded data read as 0.0, as in the reference's bounds-checked loops
// (src/io/TiffReader.cpp:376-381, :419-432).
namespace {

struct RowSpan {
    int j, x0, n;
    const unsigned char* src;
    size_t avail, bit0;
};

template <class RowFn>
void decodeDirectoryRows(TiffFile& f, const Ifd& d, int W, int H, int bps, std::vector<unsigned char>& buf, RowFn&& row) {
    const size_t bytes_per_sample = bps >= 8 ? (size_t)bps / 8 : 1;
    const bool file_little = f.littleEndian();
    const uint64_t comp = d.get(T_COMPRESSION, 1), pred = d.get(T_PREDICTOR, 1);
    if (d.arr(T_TILEOFFSETS)) {                                   // tiled, reference :354-393
        const int tw = (int)d.get(T_TILEWIDTH, 0), th = (int)d.get(T_TILELENGTH, 0);
        const auto* offs = d.arr(T_TILEOFFSETS);
        const auto* cnts = d.arr(T_TILEBYTECOUNTS);
        if (tw <= 0 || th <= 0 || !cnts) amrex::Abort("Invalid tile params.");
        if (cnts->size() < offs->size()) throw std::runtime_error("TIFF directory with fewer tile byte counts than tile offsets.");
        const int tiles_x = (W + tw - 1) / tw;
        for (size_t t = 0; t < offs->size(); ++t) {
            buf.resize((size_t)(*cnts)[t]);
            f.readAt((*offs)[t], buf.data(), buf.size());
            if (comp != C_NONE || pred != 1)
                decodeSegment(buf, comp, pred, bps == 1 ? ((size_t)tw * th + 7) / 8 : (size_t)tw * th * bytes_per_sample,
                              (size_t)tw, bytes_per_sample, file_little);
            const size_t nbytes = buf.size();
            const int ox = (int)(t % tiles_x) * tw, oy = (int)(t / tiles_x) * th;
            const int n = std::min(ox + tw, W) - ox;
            if (n <= 0) continue;
            for (int j = oy; j < std::min(oy + th, H); ++j) {
                if (bps == 1) {                                   // bits run linearly through the tile (:376-381)
                    row(RowSpan{j, ox, n, buf.data(), nbytes, (size_t)(j - oy) * (size_t)tw});
                } else {
                    const size_t off = (size_t)(j - oy) * (size_t)tw * bytes_per_sample;
                    row(RowSpan{j, ox, n, buf.data() + std::min(off, nbytes), off < nbytes ? nbytes - off : 0, 0});
                }
            }
        }
    } else {                                                       // strips, reference :394-437
        uint64_t rps = d.get(T_ROWSPERSTRIP, (uint64_t)H);
        if (rps == 0 || rps > (uint64_t)H) rps = (uint64_t)H;
        const auto* offs = d.arr(T_STRIPOFFSETS);
        const auto* cnts = d.arr(T_STRIPBYTECOUNTS);
        if (!offs) amrex::Abort("[TiffReader] TIFF directory without strip offsets.");
        const size_t pitch1 = ((size_t)W + 7) / 8;                 // TIFFScanlineSize for 1-bit data
        for (size_t s = 0; s < offs->size(); ++s) {
            const int oy = (int)(s * rps);
            if (oy >= H) break;
            const int rows = (int)std::min<uint64_t>(rps, (uint64_t)(H - oy));
            const size_t expect = bps == 1 ? pitch1 * rows : (size_t)W * rows * bytes_per_sample;
            if (comp == C_NONE) {
                buf.resize(cnts && s < cnts->size() ? std::min<size_t>((size_t)(*cnts)[s], expect) : expect);
                f.readAt((*offs)[s], buf.data(), buf.size());
            } else {
                if (!cnts || s >= cnts->size()) amrex::Abort("[TiffReader] compressed TIFF without strip byte counts.");
                buf.resize((size_t)(*cnts)[s]);
                f.readAt((*offs)[s], buf.data(), buf.size());
            }
            if (comp != C_NONE || pred != 1) decodeSegment(buf, comp, pred, expect, (size_t)W, bytes_per_sample, file_little);
            const size_t nbytes = buf.size();
            const size_t row_bytes = bps == 1 ? pitch1 : (size_t)W * bytes_per_sample;
            for (int j = oy; j < oy + rows; ++j) {
                const size_t off = (size_t)(j - oy) * row_bytes;
                row(RowSpan{j, 0, W, buf.data() + std::min(off, nbytes), off < nbytes ? nbytes - off : 0, 0});
            }
        }
    }
}

int ioThreads(int planes) {
    int t = (int)std::thread::hardware_concurrency();
    if (t <= 0) t = 1;
    t = std::min(t, 16);
    if (const char* e = std::getenv("OI_IO_THREADS")) t = std::atoi(e);
    return std::max(1, std::min(t, planes));
}

}  // namespace

// out[((k - z_begin) * H + j) * W + i] = (double(sample) > thr) ? vt : vf for planes
// [z_begin, z_begin + nz).  Planes are independent, so they are decoded by up to 16 threads
// (OI_IO_THREADS), each with its own file handle; 1-, 8- and 16-bit integer samples go
// through a lookup table of the rule instead of a per-sample conversion.
template <class OutT>
void TiffReader::thresholdInto(double thr, OutT vt, OutT vf, int z_begin, int nz, OutT* out) const {
    const int W = m_width, H = m_height, bps = m_bits_per_sample, fmt = m_sample_format;
    const size_t bytes_per_sample = bps >= 8 ? (size_t)bps / 8 : 1;
    std::vector<OutT> lut;
    if (bps == 1) {
        lut = {(0.0 > thr) ? vt : vf, (1.0 > thr) ? vt : vf};
    } else if ((bps == 8 || bps == 16) && (fmt == 1 || fmt == 2)) {
        lut.resize((size_t)1 << bps);
        for (size_t v = 0; v < lut.size(); ++v) {
            const double sv = fmt == 1 ? (double)v : (bps == 8 ? (double)(int8_t)(uint8_t)v : (double)(int16_t)(uint16_t)v);
            lut[v] = (sv > thr) ? vt : vf;
        }
    }
    const OutT zero_value = (0.0 > thr) ? vt : vf;                 // samples past the decoded data
    std::vector<Ifd> dirs;
    if (!m_is_sequence) {
        try {
            TiffFile f(m_filename);
            dirs = f.directories();                                // parsed once, shared read-only
        } catch (const std::exception& e) {
            amrex::Abort(std::string("[TiffReader] ") + e.what());
        }
    }
    const int T = ioThreads(nz);
    std::vector<std::string> errors((size_t)T);
    auto work = [&](int t) {
        try {
            const int lo = z_begin + (int)((long long)nz * t / T), hi = z_begin + (int)((long long)nz * (t + 1) / T);
            std::vector<unsigned char> buf;
            std::unique_ptr<TiffFile> stack;
            if (!m_is_sequence && lo < hi) stack.reset(new TiffFile(m_filename));
            for (int k = lo; k < hi; ++k) {
                std::unique_ptr<TiffFile> single;
                std::vector<Ifd> one;
                const Ifd* d = nullptr;
                TiffFile* f = stack.get();
                if (m_is_sequence) {
                    const std::string name = sequenceName(m_base_pattern, m_start_index + k, m_digits, m_suffix);
                    single.reset(new TiffFile(name));
                    one = single->directories();
                    if (one.empty()) throw std::runtime_error("Seq: empty file " + name);
                    d = &one[0];
                    f = single.get();
                    verifyDirectory(*d, name, W, H, bps, fmt, m_fill_order);
                } else {
                    if (k >= (int)dirs.size()) break;
                    d = &dirs[(size_t)k];
                    verifyDirectory(*d, m_filename + " (directory " + std::to_string(k) + ")", W, H, bps, fmt, m_fill_order);
                }
                const bool file_little = f->littleEndian();
                OutT* plane = out + (size_t)(k - z_begin) * (size_t)H * (size_t)W;
                decodeDirectoryRows(*f, *d, W, H, bps, buf, [&](const RowSpan& r) {
                    OutT* dst = plane + (size_t)r.j * (size_t)W + (size_t)r.x0;
                    if (bps == 1) {
                        for (int i = 0; i < r.n; ++i) {
                            const size_t lin = r.bit0 + (size_t)i, byte_i = lin >> 3;
                            dst[i] = byte_i < r.avail ? lut[(r.src[byte_i] >> (7 - (int)(lin & 7))) & 1] : zero_value;
                        }
                        return;
                    }
                    const int n_ok = (int)std::min<size_t>((size_t)r.n, r.avail / bytes_per_sample);
                    if (!lut.empty() && bps == 8) {
                        for (int i = 0; i < n_ok; ++i) dst[i] = lut[r.src[i]];
                    } else if (!lut.empty()) {                     // 16 bits, file byte order
                        if (file_little) for (int i = 0; i < n_ok; ++i) dst[i] = lut[(size_t)r.src[2 * i] | ((size_t)r.src[2 * i + 1] << 8)];
                        else for (int i = 0; i < n_ok; ++i) dst[i] = lut[((size_t)r.src[2 * i] << 8) | (size_t)r.src[2 * i + 1]];
                    } else {
                        for (int i = 0; i < n_ok; ++i)
                            dst[i] = (sampleAsDouble(r.src + (size_t)i * bytes_per_sample, bps, fmt, file_little) > thr) ? vt : vf;
                    }
                    for (int i = n_ok; i < r.n; ++i) dst[i] = zero_value;
                });
            }
        } catch (const std::exception& e) {
            errors[(size_t)t] = e.what();
        }
    };
    if (T == 1) {
        work(0);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < T; ++t) pool.emplace_back(work, t);
        for (auto& th : pool) th.join();
    }
    for (const std::string& e : errors)
        if (!e.empty()) amrex::Abort("[TiffReader] " + e);
}

void TiffReader::readDistributedIntoFab(amrex::iMultiFab& dest, int v_true, int v_false, double thr) const {
    if (!m_is_read) amrex::Abort("[TiffReader::readDistributedIntoFab] Metadata not processed.");
    AMREX_ALWAYS_ASSERT_WITH_MESSAGE(dest.nComp() == 1, "Dest MF must have 1 component.");
    AMREX_ALWAYS_ASSERT_WITH_MESSAGE(dest.nGrow() == 0, "Dest MF must have 0 ghost cells.");
    // a ghost-free field over the image box is one dense x-fastest array
    // one rank: the whole image; z-slabs: this rank's planes (full rows and columns)
    const amrex::Box vb = dest.validBox(), full = box();
    AMREX_ALWAYS_ASSERT_WITH_MESSAGE(vb.smallEnd(0) == full.smallEnd(0) && vb.bigEnd(0) == full.bigEnd(0) &&
                                     vb.smallEnd(1) == full.smallEnd(1) && vb.bigEnd(1) == full.bigEnd(1) &&
                                     vb.smallEnd(2) >= full.smallEnd(2) && vb.bigEnd(2) <= full.bigEnd(2),
                                     "Dest MF must cover the image box (or a z-slab of it).");
    thresholdInto<int>(thr, v_true, v_false, vb.smallEnd(2), vb.length(2), &dest(vb.smallEnd(0), vb.smallEnd(1), vb.smallEnd(2)));
}

void TiffReader::thresholdPlanesU8(double raw_threshold, unsigned char value_if_true, unsigned char value_if_false,
                                   int z_begin, int nz, unsigned char* out) const {
    if (!m_is_read) amrex::Abort("[TiffReader::thresholdPlanesU8] Metadata not processed.");
    if (z_begin < 0 || nz < 0 || z_begin + nz > m_depth) amrex::Abort("[TiffReader::thresholdPlanesU8] plane range outside the stack.");
    thresholdInto<unsigned char>(raw_threshold, value_if_true, value_if_false, z_begin, nz, out);
}

void TiffReader::threshold(double raw_threshold, int value_if_true, int value_if_false, amrex::iMultiFab& mf) const {
    readDistributedIntoFab(mf, value_if_true, value_if_false, raw_threshold);
}

void TiffReader::threshold(double raw_threshold, amrex::iMultiFab& mf) const { threshold(raw_threshold, 1, 0, mf); }

}  // namespace OpenImpala
