// DatReader (reference src/io/DatReader.cpp:60-248): header + uint16 payload, both
// little-endian on disk; bytes are assembled explicitly so the host byte order does not matter.
#include "DatReader.H"

#include <fstream>
#include <stdexcept>

#include <AMReX_Print.H>

namespace OpenImpala {

namespace {
inline std::int32_t le32(const unsigned char* p) {
    return (std::int32_t)((std::uint32_t)p[0] | ((std::uint32_t)p[1] << 8) | ((std::uint32_t)p[2] << 16) |
                          ((std::uint32_t)p[3] << 24));
}
}  // namespace

DatReader::DatReader(const std::string& filename) {
    if (!readFile(filename)) throw std::runtime_error("DatReader: Failed to read file: " + filename);
}

bool DatReader::readFile(const std::string& filename) {
    m_filename = filename;
    m_raw.clear();
    m_width = m_height = m_depth = 0;
    m_is_read = false;
    std::ifstream ifs(m_filename, std::ios::binary | std::ios::ate);
    if (!ifs.is_open()) {
        amrex::Print() << "Error: [DatReader] Could not open file: " << m_filename << "\n";
        return false;
    }
    const std::streamsize file_size = ifs.tellg();
    ifs.seekg(0, std::ios::beg);
    constexpr std::streamsize header_size = 12;
    if (file_size < header_size) {
        amrex::Print() << "Error: [DatReader] File too small for header: " << m_filename << "\n";
        return false;
    }
    unsigned char hdr[12];
    ifs.read(reinterpret_cast<char*>(hdr), header_size);
    if (!ifs.good()) {
        amrex::Print() << "Error: [DatReader] Failed reading header: " << m_filename << "\n";
        return false;
    }
    const std::int32_t dims[3] = {le32(hdr), le32(hdr + 4), le32(hdr + 8)};
    if (dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0) {
        amrex::Print() << "Error: [DatReader] Invalid dimensions in header: " << m_filename << " (W=" << dims[0]
                       << ", H=" << dims[1] << ", D=" << dims[2] << ")\n";
        return false;
    }
    const long long n = (long long)dims[0] * dims[1] * dims[2];
    const std::streamsize expected = (std::streamsize)(n * (long long)sizeof(DataType));
    const std::streamsize actual = file_size - header_size;
    if (actual < expected) {
        amrex::Print() << "Error: [DatReader] File size mismatch: " << m_filename << ". Expected data: " << expected
                       << " bytes, Available: " << actual << " bytes.\n";
        return false;
    }
    if (actual > expected) amrex::Warning("Warning: [DatReader] File contains more data than expected. Ignoring extra data.");
    std::vector<unsigned char> bytes((size_t)expected);
    ifs.read(reinterpret_cast<char*>(bytes.data()), expected);
    if (!ifs.good() && ifs.gcount() != expected) {
        amrex::Print() << "Error: [DatReader] Failed reading voxel data: " << m_filename << "\n";
        return false;
    }
    m_raw.resize((size_t)n);
    for (size_t v = 0; v < (size_t)n; ++v) m_raw[v] = (DataType)(bytes[2 * v] | (bytes[2 * v + 1] << 8));
    m_width = dims[0]; m_height = dims[1]; m_depth = dims[2];
    m_is_read = true;
    amrex::Print() << "Successfully read DAT file: " << m_filename << " (Dimensions: " << m_width << "x" << m_height
                   << "x" << m_depth << ")\n";
    return true;
}

amrex::Box DatReader::box() const {
    if (!m_is_read) return amrex::Box();
    return amrex::Box(amrex::IntVect(0, 0, 0), amrex::IntVect(m_width - 1, m_height - 1, m_depth - 1));
}

DatReader::DataType DatReader::getRawValue(int i, int j, int k) const {
    if (!m_is_read) throw std::out_of_range("[DatReader::getRawValue] Data not read yet.");
    if (i < 0 || i >= m_width || j < 0 || j >= m_height || k < 0 || k >= m_depth)
        throw std::out_of_range("[DatReader::getRawValue] Index (" + std::to_string(i) + "," + std::to_string(j) + "," +
                                std::to_string(k) + ") out of bounds (W:" + std::to_string(m_width) + ", H:" +
                                std::to_string(m_height) + ", D:" + std::to_string(m_depth) + ").");
    return m_raw[((size_t)k * m_height + (size_t)j) * m_width + (size_t)i];
}

void DatReader::threshold(DataType raw_threshold, int value_if_true, int value_if_false, amrex::iMultiFab& mf) const {
    if (!m_is_read) amrex::Abort("[DatReader::threshold] Cannot threshold, data not read successfully.");
    const amrex::Box& b = mf.validBox();
    for (int k = b.smallEnd(2); k <= b.bigEnd(2); ++k)
        for (int j = b.smallEnd(1); j <= b.bigEnd(1); ++j)
            for (int i = b.smallEnd(0); i <= b.bigEnd(0); ++i) {
                const bool inside = i >= 0 && i < m_width && j >= 0 && j < m_height && k >= 0 && k < m_depth;
                mf(i, j, k, 0) = (inside && m_raw[((size_t)k * m_height + (size_t)j) * m_width + (size_t)i] > raw_threshold)
                                     ? value_if_true : value_if_false;
            }
}

void DatReader::threshold(DataType raw_threshold, amrex::iMultiFab& mf) const { threshold(raw_threshold, 1, 0, mf); }

}  // namespace OpenImpala
