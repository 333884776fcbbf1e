// Behaviour of src/io/RawReader.cpp: whole file into memory (:76-200), voxel
// (i,j,k) at ((k*H + j)*W + i) * bytes (:310-313), byte swap when file and host
// endianness differ, threshold rule value > t ? a : b (:379-491).
#include "RawReader.H"

#include <algorithm>
#include <cstdlib>
#include <thread>

#include <cstdint>
#include <cstring>
#include <fstream>
#include <stdexcept>

#include <AMReX_Utility.H>

namespace OpenImpala {

RawReader::RawReader() = default;

RawReader::RawReader(const std::string& filename, int width, int height, int depth, RawDataType data_type) {
    if (!readFile(filename, width, height, depth, data_type))
        throw std::runtime_error("RawReader: failed to read file " + filename);
}

size_t RawReader::getBytesPerVoxel() const {
    switch (m_data_type) {
        case RawDataType::UINT8: case RawDataType::INT8: return 1;
        case RawDataType::INT16_LE: case RawDataType::INT16_BE:
        case RawDataType::UINT16_LE: case RawDataType::UINT16_BE: return 2;
        case RawDataType::INT32_LE: case RawDataType::INT32_BE: case RawDataType::UINT32_LE:
        case RawDataType::UINT32_BE: case RawDataType::FLOAT32_LE: case RawDataType::FLOAT32_BE: return 4;
        case RawDataType::FLOAT64_LE: case RawDataType::FLOAT64_BE: return 8;
        default: return 0;
    }
}

bool RawReader::isHostLittleEndian() const {
    const uint16_t probe = 1;
    unsigned char b;
    std::memcpy(&b, &probe, 1);
    return b == 1;
}

bool RawReader::readFile(const std::string& filename, int width, int height, int depth, RawDataType data_type) {
    m_is_read = false;
    m_filename = filename; m_width = width; m_height = height; m_depth = depth; m_data_type = data_type;
    if (width <= 0 || height <= 0 || depth <= 0) { amrex::Warning("RawReader: dimensions must be positive"); return false; }
    const size_t bpv = getBytesPerVoxel();
    if (bpv == 0) { amrex::Warning("RawReader: unknown data type"); return false; }
    const size_t expect = (size_t)width * height * depth * bpv;
    std::ifstream in(filename, std::ios::binary | std::ios::ate);
    if (!in) { amrex::Warning("RawReader: cannot open " + filename); return false; }
    const size_t fsize = (size_t)in.tellg();
    if (fsize != expect) {
        amrex::Warning("RawReader: file size " + std::to_string(fsize) + " != expected " + std::to_string(expect));
        return false;
    }
    m_raw_bytes.resize(expect);
    in.seekg(0);
    in.read(reinterpret_cast<char*>(m_raw_bytes.data()), (std::streamsize)expect);
    if ((size_t)in.gcount() != expect) { amrex::Warning("RawReader: short read"); return false; }
    m_is_read = true;
    return true;
}

amrex::Box RawReader::box() const {
    if (!m_is_read) return amrex::Box();
    return amrex::Box(amrex::IntVect::TheZeroVector(), amrex::IntVect(m_width - 1, m_height - 1, m_depth - 1));
}

double RawReader::getValue(int i, int j, int k) const {
    if (!m_is_read) throw std::runtime_error("RawReader::getValue: no data");
    if (i < 0 || i >= m_width || j < 0 || j >= m_height || k < 0 || k >= m_depth)
        throw std::out_of_range("RawReader::getValue: index outside the volume");
    const size_t bpv = getBytesPerVoxel();
    const unsigned char* p = m_raw_bytes.data() + (((size_t)k * m_height + (size_t)j) * m_width + (size_t)i) * bpv;
    bool file_le = true;
    switch (m_data_type) {
        case RawDataType::INT16_BE: case RawDataType::UINT16_BE: case RawDataType::INT32_BE:
        case RawDataType::UINT32_BE: case RawDataType::FLOAT32_BE: case RawDataType::FLOAT64_BE: file_le = false; break;
        default: break;
    }
    unsigned char b[8] = {0};
    const bool swap = (bpv > 1) && (file_le != isHostLittleEndian());
    for (size_t q = 0; q < bpv; ++q) b[q] = swap ? p[bpv - 1 - q] : p[q];
    switch (m_data_type) {
        case RawDataType::UINT8: return (double)b[0];
        case RawDataType::INT8: { int8_t v; std::memcpy(&v, b, 1); return (double)v; }
        case RawDataType::INT16_LE: case RawDataType::INT16_BE: { int16_t v; std::memcpy(&v, b, 2); return (double)v; }
        case RawDataType::UINT16_LE: case RawDataType::UINT16_BE: { uint16_t v; std::memcpy(&v, b, 2); return (double)v; }
        case RawDataType::INT32_LE: case RawDataType::INT32_BE: { int32_t v; std::memcpy(&v, b, 4); return (double)v; }
        case RawDataType::UINT32_LE: case RawDataType::UINT32_BE: { uint32_t v; std::memcpy(&v, b, 4); return (double)v; }
        case RawDataType::FLOAT32_LE: case RawDataType::FLOAT32_BE: { float v; std::memcpy(&v, b, 4); return (double)v; }
        case RawDataType::FLOAT64_LE: case RawDataType::FLOAT64_BE: { double v; std::memcpy(&v, b, 8); return v; }
        default: return 0.0;
    }
}

// out[((k - z_begin) * H + j) * W + i] = (getValue(i, j, k) > t) ? vt : vf, the reference's
// rule (src/io/RawReader.cpp:379-491), evaluated through a lookup table for the 8- and 16-bit
// types and plane-parallel on up to 16 threads (OI_IO_THREADS).
template <class OutT>
void RawReader::thresholdInto(double t, OutT vt, OutT vf, int z_begin, int nz, OutT* out) const {
    const size_t bpv = getBytesPerVoxel();
    const size_t plane = (size_t)m_width * (size_t)m_height;
    std::vector<OutT> lut;
    bool be = false, is_signed = false;
    switch (m_data_type) {
        case RawDataType::INT8: is_signed = true; break;
        case RawDataType::INT16_LE: is_signed = true; break;
        case RawDataType::INT16_BE: is_signed = true; be = true; break;
        case RawDataType::UINT16_BE: be = true; break;
        default: break;
    }
    if (bpv == 1 || (bpv == 2 && (m_data_type == RawDataType::INT16_LE || m_data_type == RawDataType::INT16_BE ||
                                  m_data_type == RawDataType::UINT16_LE || m_data_type == RawDataType::UINT16_BE))) {
        lut.resize((size_t)1 << (8 * bpv));
        for (size_t v = 0; v < lut.size(); ++v) {
            const double sv = !is_signed ? (double)v : (bpv == 1 ? (double)(int8_t)(uint8_t)v : (double)(int16_t)(uint16_t)v);
            lut[v] = (sv > t) ? vt : vf;
        }
    }
    int T = (int)std::thread::hardware_concurrency();
    if (T <= 0) T = 1;
    T = std::min(T, 16);
    if (const char* e = std::getenv("OI_IO_THREADS")) T = std::atoi(e);
    T = std::max(1, std::min(T, nz));
    auto work = [&](int w) {
        const int lo = z_begin + (int)((long long)nz * w / T), hi = z_begin + (int)((long long)nz * (w + 1) / T);
        for (int k = lo; k < hi; ++k) {
            const unsigned char* src = m_raw_bytes.data() + (size_t)k * plane * bpv;
            OutT* dst = out + (size_t)(k - z_begin) * plane;
            if (!lut.empty() && bpv == 1) {
                for (size_t q = 0; q < plane; ++q) dst[q] = lut[src[q]];
            } else if (!lut.empty()) {
                if (!be) for (size_t q = 0; q < plane; ++q) dst[q] = lut[(size_t)src[2 * q] | ((size_t)src[2 * q + 1] << 8)];
                else for (size_t q = 0; q < plane; ++q) dst[q] = lut[((size_t)src[2 * q] << 8) | (size_t)src[2 * q + 1]];
            } else {
                size_t q = 0;
                for (int j = 0; j < m_height; ++j)
                    for (int i = 0; i < m_width; ++i, ++q) dst[q] = (getValue(i, j, k) > t) ? vt : vf;
            }
        }
    };
    if (T == 1) {
        work(0);
    } else {
        std::vector<std::thread> pool;
        for (int w = 0; w < T; ++w) pool.emplace_back(work, w);
        for (auto& th : pool) th.join();
    }
}

void RawReader::threshold(double t, int value_if_true, int value_if_false, amrex::iMultiFab& mf) const {
    if (!m_is_read) amrex::Abort("RawReader::threshold: no data has been read");
    AMREX_ALWAYS_ASSERT_WITH_MESSAGE(mf.boxArray().minimalBox() == this->box(), "RawReader: iMultiFab domain mismatch");
    {   // ghost-free field over the whole volume or over a z-slab of it: one dense x-fastest array
        const amrex::Box vb = mf.validBox(), full = this->box();
        if (mf.nGrow() == 0 && vb.smallEnd(0) == full.smallEnd(0) && vb.bigEnd(0) == full.bigEnd(0) &&
            vb.smallEnd(1) == full.smallEnd(1) && vb.bigEnd(1) == full.bigEnd(1) && vb.smallEnd(2) >= full.smallEnd(2) &&
            vb.bigEnd(2) <= full.bigEnd(2)) {
            thresholdInto<int>(t, value_if_true, value_if_false, vb.smallEnd(2), vb.length(2),
                               &mf(vb.smallEnd(0), vb.smallEnd(1), vb.smallEnd(2)));
            return;
        }
    }
    const amrex::Box& b = mf.validBox();
    for (int k = b.smallEnd(2); k <= b.bigEnd(2); ++k)
        for (int j = b.smallEnd(1); j <= b.bigEnd(1); ++j)
            for (int i = b.smallEnd(0); i <= b.bigEnd(0); ++i)
                mf(i, j, k) = (getValue(i, j, k) > t) ? value_if_true : value_if_false;
}

void RawReader::thresholdPlanesU8(double t, unsigned char value_if_true, unsigned char value_if_false, int z_begin,
                                  int nz, unsigned char* out) const {
    if (!m_is_read) amrex::Abort("RawReader::thresholdPlanesU8: no data has been read");
    if (z_begin < 0 || nz < 0 || z_begin + nz > m_depth) amrex::Abort("RawReader::thresholdPlanesU8: plane range outside the volume");
    thresholdInto<unsigned char>(t, value_if_true, value_if_false, z_begin, nz, out);
}

void RawReader::threshold(double t, amrex::iMultiFab& mf) const { threshold(t, 1, 0, mf); }

}  // namespace OpenImpala
