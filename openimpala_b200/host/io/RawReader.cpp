// Behaviour of src/io/RawReader.cpp: whole file into memory (:76-200), voxel
// (i,j,k) at ((k*H + j)*W + i) * bytes (:310-313), byte swap when file and host
// endianness differ, threshold rule value > t ? a : b (:379-491).
#include "RawReader.H"

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

void RawReader::threshold(double t, int value_if_true, int value_if_false, amrex::iMultiFab& mf) const {
    if (!m_is_read) amrex::Abort("RawReader::threshold: no data has been read");
    AMREX_ALWAYS_ASSERT_WITH_MESSAGE(mf.boxArray().minimalBox() == this->box(), "RawReader: iMultiFab domain mismatch");
    const amrex::Box& b = mf.validBox();
    for (int k = b.smallEnd(2); k <= b.bigEnd(2); ++k)
        for (int j = b.smallEnd(1); j <= b.bigEnd(1); ++j)
            for (int i = b.smallEnd(0); i <= b.bigEnd(0); ++i)
                mf(i, j, k) = (getValue(i, j, k) > t) ? value_if_true : value_if_false;
}

void RawReader::threshold(double t, amrex::iMultiFab& mf) const { threshold(t, 1, 0, mf); }

}  // namespace OpenImpala
