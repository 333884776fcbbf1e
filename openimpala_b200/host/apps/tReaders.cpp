// Reader + VolumeFraction checks in the mould of src/io/tTiffReader.cpp,
// tRawReader.cpp, tHDF5Reader.cpp and src/props/tVolumeFraction.cpp:
//   mode = tiff | tiffseq | raw | hdf5 | dat ; prints dims, sample metadata, thresholded min/max
//   and the phase counts, compares the GPU count with a direct host loop.
#include <functional>
#include <iomanip>
#include <vector>
#include <string>

#include <AMReX.H>
#include <AMReX_MultiFab.H>
#include <AMReX_ParmParse.H>
#include <AMReX_PlotFileUtil.H>
#include <AMReX_Print.H>

#include "../io/DatReader.H"
#include "../io/HDF5Reader.H"
#include "../io/RawReader.H"
#include "../io/TiffReader.H"
#include "../props/VolumeFraction.H"

int main(int argc, char* argv[]) {
    amrex::Initialize(argc, argv);
    bool passed = true;
    auto fail = [&](const std::string& why) { passed = false; amrex::Print() << "TEST FAILED: " << why << "\n"; };
    {
        std::string mode = "tiff", file, dataset = "image", datatype = "UINT8";
        int width = 0, height = 0, depth = 0, box_size = 32, use_gpu_count = 1;
        amrex::Real threshold = 0.5;
        amrex::ParmParse pp;
        pp.query("mode", mode);
        if (!pp.query("tifffile", file) && !pp.query("rawfile", file) && !pp.query("hdf5file", file) &&
            !pp.query("datfile", file)) pp.get("filename", file);
        pp.query("hdf5dataset", dataset);
        pp.query("datatype", datatype);
        pp.query("width", width); pp.query("height", height); pp.query("depth", depth);
        pp.query("threshold", threshold); pp.query("threshold_val", threshold);
        pp.query("box_size", box_size);
        pp.query("gpu_count", use_gpu_count);

        amrex::Box domain;
        amrex::BoxArray ba;
        amrex::DistributionMapping dm;
        amrex::iMultiFab mf;
        auto prepare = [&](const amrex::Box& b) {
            domain = b; ba.define(b); ba.maxSize(box_size); dm.define(ba); mf.define(ba, dm, 1, 0);
        };
        // u8_chunk = N: the streamed-upload entry point (thresholdPlanesU8, N planes at a time)
        // must reproduce the iMultiFab path voxel for voxel
        int u8_chunk = 0;
        pp.query("u8_chunk", u8_chunk);
        auto check_u8 = [&](const std::function<void(int, int, unsigned char*)>& planes) {
            if (u8_chunk <= 0) return;
            const size_t plane = (size_t)domain.length(0) * (size_t)domain.length(1);
            std::vector<unsigned char> buf((size_t)u8_chunk * plane);
            long long bad = 0;
            for (int z0 = 0; z0 < domain.length(2); z0 += u8_chunk) {
                const int nz = std::min(u8_chunk, domain.length(2) - z0);
                std::fill(buf.begin(), buf.end(), (unsigned char)0xee);
                planes(z0, nz, buf.data());
                size_t q = 0;
                for (int k = z0; k < z0 + nz; ++k)
                    for (int j = 0; j < domain.length(1); ++j)
                        for (int i = 0; i < domain.length(0); ++i, ++q)
                            if ((int)buf[q] != mf(i, j, k, 0)) ++bad;
            }
            amrex::Print() << "U8ChunkMismatches: " << bad << "\n";
            if (bad) fail("thresholdPlanesU8 differs from threshold()");
        };
        const double t_read0 = amrex::second();
        try {
            if (mode == "tiff") {
                OpenImpala::TiffReader r(file);
                amrex::Print() << "BitsPerSample: " << r.bitsPerSample() << " SampleFormat: " << r.sampleFormat()
                               << " SamplesPerPixel: " << r.samplesPerPixel() << "\n";
                prepare(r.box());
                r.threshold(threshold, 1, 0, mf);
                check_u8([&](int z0, int nz, unsigned char* out) { r.thresholdPlanesU8(threshold, 1, 0, z0, nz, out); });
            } else if (mode == "tiffseq") {
                std::string suffix = ".tif";
                int num_files = 0, start_index = 0, digits = 1;
                pp.get("num_files", num_files);
                pp.query("start_index", start_index);
                pp.query("digits", digits);
                pp.query("suffix", suffix);
                OpenImpala::TiffReader r(file, num_files, start_index, digits, suffix);   // file = base pattern
                prepare(r.box());
                r.threshold(threshold, 1, 0, mf);
                check_u8([&](int z0, int nz, unsigned char* out) { r.thresholdPlanesU8(threshold, 1, 0, z0, nz, out); });
            } else if (mode == "raw") {
                OpenImpala::RawDataType t = OpenImpala::RawDataType::UINT8;
                if (datatype == "INT16_LE") t = OpenImpala::RawDataType::INT16_LE;
                else if (datatype == "UINT16_LE") t = OpenImpala::RawDataType::UINT16_LE;
                else if (datatype == "UINT16_BE") t = OpenImpala::RawDataType::UINT16_BE;
                else if (datatype == "FLOAT32_LE") t = OpenImpala::RawDataType::FLOAT32_LE;
                else if (datatype != "UINT8") fail("datatype not handled by this driver: " + datatype);
                OpenImpala::RawReader r(file, width, height, depth, t);
                prepare(r.box());
                r.threshold(threshold, 1, 0, mf);
                check_u8([&](int z0, int nz, unsigned char* out) { r.thresholdPlanesU8(threshold, 1, 0, z0, nz, out); });
            } else if (mode == "hdf5") {
                OpenImpala::HDF5Reader r(file, dataset);
                prepare(r.box());
                r.threshold(threshold, 1, 0, mf);
                check_u8([&](int z0, int nz, unsigned char* out) { r.thresholdPlanesU8(threshold, 1, 0, z0, nz, out); });
            } else if (mode == "dat") {
                OpenImpala::DatReader r(file);
                prepare(r.box());
                r.threshold(static_cast<OpenImpala::DatReader::DataType>(threshold), 1, 0, mf);
                amrex::Print() << "RawCorner: " << r.getRawValue(0, 0, 0) << " "
                               << r.getRawValue(r.width() - 1, r.height() - 1, r.depth() - 1) << "\n";
            } else {
                fail("unknown mode " + mode);
            }
        } catch (const std::exception& e) {
            fail(e.what());
        }
        if (passed) {
            const double t_read = amrex::second() - t_read0;
            amrex::Print() << "ReadThresholdSeconds: " << t_read << " (" << domain.numPts() / std::max(t_read, 1e-9) * 1e-6
                           << " Mvoxel/s, open + decode + threshold into the int32 field)\n";
            amrex::Print() << "Dims: " << domain.length(0) << " " << domain.length(1) << " " << domain.length(2) << "\n";
            amrex::Print() << "ThresholdMinMax: " << mf.min(0) << " " << mf.max(0) << "\n";
            long long direct1 = mf.sum(0), total = domain.numPts();
            amrex::Print() << "DirectCount1: " << direct1 << " Total: " << total << "\n";
            {   // position-weighted sum: catches a voxel that is right in value but wrong in place
                unsigned long long chk = 0, idx = 0;
                for (int k = domain.smallEnd(2); k <= domain.bigEnd(2); ++k)
                    for (int j = domain.smallEnd(1); j <= domain.bigEnd(1); ++j)
                        for (int i = domain.smallEnd(0); i <= domain.bigEnd(0); ++i, ++idx)
                            chk += (idx % 65521ull + 1ull) * (unsigned long long)(mf(i, j, k, 0) & 0xff);
                amrex::Print() << "WeightedChecksum: " << chk << "\n";
            }
            std::string plotfile;
            if (pp.query("plotfile", plotfile)) {
                // the thresholded field and an analytic ramp as an AMReX plotfile (the writer the
                // TortuosityHypre / EffectiveDiffusivityHypre classes use for write_plotfile = 1)
                amrex::RealBox rb({AMREX_D_DECL(0.0, 0.0, 0.0)},
                                  {AMREX_D_DECL(amrex::Real(domain.length(0)), amrex::Real(domain.length(1)),
                                                amrex::Real(domain.length(2)))});
                const int is_per[3] = {0, 0, 0};
                amrex::Geometry geom(domain, &rb, 0, is_per);
                amrex::MultiFab plot(ba, dm, 2, 0);
                for (int k = domain.smallEnd(2); k <= domain.bigEnd(2); ++k)
                    for (int j = domain.smallEnd(1); j <= domain.bigEnd(1); ++j)
                        for (int i = domain.smallEnd(0); i <= domain.bigEnd(0); ++i) {
                            plot(i, j, k, 0) = amrex::Real(mf(i, j, k, 0));
                            plot(i, j, k, 1) = i + 0.5 * j - 0.25 * k;
                        }
                amrex::WriteSingleLevelPlotfile(plotfile, plot, {"phase_id", "ramp"}, geom, 0.0, 0);
                amrex::Print() << "Plotfile: " << plotfile << "\n";
            }
            if (use_gpu_count) {
                for (int phase = 0; phase <= 1; ++phase) {
                    OpenImpala::VolumeFraction vf(mf, phase);
                    long long pc = 0, tc = 0;
                    vf.value(pc, tc);
                    amrex::Print() << "VolumeFractionCount" << phase << ": " << pc << " of " << tc << "\n";
                    const long long expect = phase == 1 ? direct1 : total - direct1;
                    if (pc != expect || tc != total) fail("VolumeFraction count differs from the direct loop");
                }
            }
        }
        amrex::Print() << (passed ? "TEST PASSED\n" : "TEST FAILED\n");
    }
    amrex::Finalize();
    return passed ? 0 : 1;
}
