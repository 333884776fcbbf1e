// `Diffusion <inputs> [key=value ...]` -- the reference's application driver
// (src/props/Diffusion.cpp): same inputs keys and defaults (:200-223, :605-611),
// same readers by file extension (:262-300).
//   calculation_method = homogenization (the default, :511-589): three periodic
//     corrector solves (EffectiveDiffusivityHypre) and the D_eff tensor (:60-167);
//   calculation_method = flow_through (:590-732): TortuosityHypre per direction and
//     results.txt.
//   rev.do_study = 1 (:317-504): D_eff tensors of random sub-volumes -> rev_study_Deff.csv.
#include <algorithm>
#include <thread>
#include <mutex>
#include <atomic>
#include <filesystem>
#include <fstream>
#include <iomanip>
#include <limits>
#include <map>
#include <memory>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include <sys/wait.h>
#include <unistd.h>

#include <AMReX.H>
#include <AMReX_ParmParse.H>
#include <AMReX_Print.H>

#include <openimpala_b200.h>

#include "../io/HDF5Reader.H"
#include "../io/RawReader.H"
#include "../io/TiffReader.H"
#include "../props/EffectiveDiffusivityHypre.H"
#include "../props/TortuosityHypre.H"
#include "../props/VolumeFraction.H"

namespace {

OpenImpala::TortuosityHypre::SolverType stringToSolverType(const std::string& solver_str) {
    std::string s = solver_str;
    std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)std::tolower(c); });
    using ST = OpenImpala::TortuosityHypre::SolverType;
    if (s == "jacobi") return ST::Jacobi;
    if (s == "gmres") return ST::GMRES;
    if (s == "flexgmres") return ST::FlexGMRES;
    if (s == "pcg") return ST::PCG;
    if (s == "bicgstab") return ST::BiCGSTAB;
    if (s == "smg") return ST::SMG;
    if (s == "pfmg") return ST::PFMG;
    amrex::Abort("Invalid solver string: '" + solver_str + "'.");
    return ST::GMRES;
}

// calculate_Deff_tensor_homogenization, src/props/Diffusion.cpp:60-167: host restatement
// on the corrector fields (1 ghost cell, periodic), used to cross-check the device sums.
void calculate_Deff_tensor_homogenization(amrex::Real Deff_tensor[AMREX_SPACEDIM][AMREX_SPACEDIM],
                                          const amrex::MultiFab& chi_x, const amrex::MultiFab& chi_y,
                                          const amrex::MultiFab& chi_z, const amrex::iMultiFab& active_mask,
                                          const amrex::Geometry& geom, int /*verbose_level*/) {
    const amrex::MultiFab* chi[3] = {&chi_x, &chi_y, &chi_z};
    const amrex::Box& bx = geom.Domain();
    double inv_2dx[3];
    for (int d = 0; d < 3; ++d) inv_2dx[d] = 1.0 / (2.0 * geom.CellSize(d));
    double sum[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int k = bx.smallEnd(2); k <= bx.bigEnd(2); ++k)
        for (int j = bx.smallEnd(1); j <= bx.bigEnd(1); ++j)
            for (int i = bx.smallEnd(0); i <= bx.bigEnd(0); ++i) {
                if (active_mask(i, j, k, 0) != 1) continue;
                for (int c = 0; c < 3; ++c) {          // corrector chi_c -> column c
                    const amrex::MultiFab& f = *chi[c];
                    const double g[3] = {(f(i + 1, j, k) - f(i - 1, j, k)) * inv_2dx[0],
                                         (f(i, j + 1, k) - f(i, j - 1, k)) * inv_2dx[1],
                                         (f(i, j, k + 1) - f(i, j, k - 1)) * inv_2dx[2]};
                    for (int r = 0; r < 3; ++r) sum[r][c] += (r == c ? 1.0 : 0.0) - g[r];
                }
            }
    const long long n = bx.numPts();
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) Deff_tensor[r][c] = n > 0 ? sum[r][c] / (amrex::Real)n : 0.0;
}

OpenImpala::RawDataType rawTypeFromString(const std::string& raw_type) {
    static const std::map<std::string, OpenImpala::RawDataType> types = {
        {"UINT8", OpenImpala::RawDataType::UINT8}, {"INT8", OpenImpala::RawDataType::INT8},
        {"INT16_LE", OpenImpala::RawDataType::INT16_LE}, {"INT16_BE", OpenImpala::RawDataType::INT16_BE},
        {"UINT16_LE", OpenImpala::RawDataType::UINT16_LE}, {"UINT16_BE", OpenImpala::RawDataType::UINT16_BE},
        {"INT32_LE", OpenImpala::RawDataType::INT32_LE}, {"INT32_BE", OpenImpala::RawDataType::INT32_BE},
        {"UINT32_LE", OpenImpala::RawDataType::UINT32_LE}, {"UINT32_BE", OpenImpala::RawDataType::UINT32_BE},
        {"FLOAT32_LE", OpenImpala::RawDataType::FLOAT32_LE}, {"FLOAT32_BE", OpenImpala::RawDataType::FLOAT32_BE},
        {"FLOAT64_LE", OpenImpala::RawDataType::FLOAT64_LE}, {"FLOAT64_BE", OpenImpala::RawDataType::FLOAT64_BE}};
    auto it = types.find(raw_type);
    if (it == types.end()) throw std::runtime_error("Unknown raw datatype: " + raw_type);
    return it->second;
}

template <class Reader>
void loadThresholded(Reader& reader, int box_size, double threshold, amrex::Box& domain, amrex::BoxArray& ba,
                     amrex::DistributionMapping& dm, amrex::iMultiFab& mf_phase) {
    domain = reader.box();
    ba.define(domain);
    ba.maxSize(box_size);
    dm.define(ba);
    mf_phase.define(ba, dm, 1, 1);
    amrex::iMultiFab tmp(ba, dm, 1, 0);
    reader.threshold(threshold, 1, 0, tmp);      // > threshold -> 1, else 0 (reference :259-271)
    amrex::Copy(mf_phase, tmp, 0, 0, 1, 0);
}

}  // namespace

// `Diffusion inputs b200.ranks=N`: the reference is started as N MPI ranks by mpirun; without MPI this
// process starts the N ranks itself (one per GPU): N copies of the same command line with OI_RANK,
// OI_WORLD_SIZE and OI_COMM_FILE in their environment, each owning one z-slab of the image.  Returns -1
// when this process should carry on (a single rank, or already one of the ranks).
int launchRanks(int argc, char* argv[]) {
    if (std::getenv("OI_RANK") || std::getenv("RANK")) return -1;
    int ranks = 1;
    for (int a = 1; a < argc; ++a) {
        const std::string arg = argv[a];
        if (arg.rfind("b200.ranks=", 0) == 0) ranks = std::atoi(arg.c_str() + 11);
    }
    if (ranks <= 1 && argc > 1) {                     // the key may also sit in the inputs file
        std::ifstream in(argv[1]);
        std::string line;
        while (std::getline(in, line)) {
            const size_t h = line.find('#');
            if (h != std::string::npos) line.erase(h);
            std::string key, eq;
            int v = 0;
            std::stringstream ss(line);
            if ((ss >> key >> eq >> v) && key == "b200.ranks" && eq == "=") ranks = v;
        }
    }
    if (ranks <= 1) return -1;
    const std::string comm_file = (std::filesystem::temp_directory_path() / ("oi_comm_" + std::to_string((long)getpid()))).string();
    std::filesystem::remove(comm_file);
    std::vector<pid_t> kids;
    for (int r = 0; r < ranks; ++r) {
        const pid_t pid = fork();
        if (pid < 0) { std::perror("fork"); return 1; }
        if (pid == 0) {
            setenv("OI_RANK", std::to_string(r).c_str(), 1);
            setenv("OI_WORLD_SIZE", std::to_string(ranks).c_str(), 1);
            setenv("OI_COMM_FILE", comm_file.c_str(), 1);
            execv("/proc/self/exe", argv);
            std::perror("execv");
            _exit(127);
        }
        kids.push_back(pid);
    }
    int rc = 0;
    for (pid_t k : kids) {
        int st = 0;
        waitpid(k, &st, 0);
        const int code = WIFEXITED(st) ? WEXITSTATUS(st) : 128 + (WIFSIGNALED(st) ? WTERMSIG(st) : 0);
        if (code != 0) rc = code;
    }
    std::filesystem::remove(comm_file);
    return rc;
}

int main(int argc, char* argv[]) {
    {
        const int launched = launchRanks(argc, argv);
        if (launched >= 0) return launched;
    }
    amrex::Initialize(argc, argv);
    {
        const amrex::Real t_start = amrex::second();
        std::string filename, data_path = "./data/", results_path = "./results_diffusion/";
        std::string hdf5_dataset = "image", solver_str = "FlexGMRES", method = "homogenization";
        std::string output_filename = "results.txt";
        amrex::Real threshold_val = 0.5;
        int phase_id = 1, box_size = 32, verbose = 1, write_plotfile = 0;
        int raw_w = 0, raw_h = 0, raw_d = 0;
        std::string raw_type = "UINT8";
        {
            amrex::ParmParse pp;
            pp.get("filename", filename);
            pp.query("data_path", data_path);
            pp.query("results_path", results_path);
            pp.query("hdf5_dataset", hdf5_dataset);
            pp.query("threshold_val", threshold_val);
            pp.query("phase_id", phase_id);
            pp.query("solver_type", solver_str);
            pp.query("box_size", box_size);
            pp.query("verbose", verbose);
            pp.query("write_plotfile", write_plotfile);
            pp.query("calculation_method", method);
            pp.query("output_filename", output_filename);
            // .raw needs its shape from the inputs (RawReader contract); keys of tRawReader.inputs
            pp.query("width", raw_w); pp.query("height", raw_h); pp.query("depth", raw_d);
            pp.query("datatype", raw_type);
        }
        std::filesystem::path results_dir(results_path);
        if (!std::filesystem::exists(results_dir)) {
            std::filesystem::create_directories(results_dir);
            if (verbose >= 1) amrex::Print() << "Created results directory: " << results_dir.string() << std::endl;
        }
        const std::filesystem::path input = std::filesystem::path(data_path) / filename;

        amrex::Geometry geom_full;
        amrex::BoxArray ba;
        amrex::DistributionMapping dm;
        amrex::iMultiFab mf_phase;
        amrex::Box domain;
        try {
            if (verbose >= 1) amrex::Print() << "Reading full domain data from: " << input.string() << std::endl;
            if (!input.has_extension()) throw std::runtime_error("File has no extension: " + input.string());
            std::string ext = input.extension().string();
            std::transform(ext.begin(), ext.end(), ext.begin(), ::tolower);
            if (ext == ".tif" || ext == ".tiff") {
                OpenImpala::TiffReader reader(input.string());
                if (!reader.isRead()) throw std::runtime_error("TiffReader failed to read metadata.");
                loadThresholded(reader, box_size, threshold_val, domain, ba, dm, mf_phase);
            } else if (ext == ".h5" || ext == ".hdf5") {
                OpenImpala::HDF5Reader reader(input.string(), hdf5_dataset);
                if (!reader.isRead()) throw std::runtime_error("HDF5Reader failed to read metadata.");
                loadThresholded(reader, box_size, threshold_val, domain, ba, dm, mf_phase);
            } else if (ext == ".raw") {
                OpenImpala::RawReader reader(input.string(), raw_w, raw_h, raw_d, rawTypeFromString(raw_type));
                loadThresholded(reader, box_size, threshold_val, domain, ba, dm, mf_phase);
            } else {
                throw std::runtime_error("Unsupported file extension for full domain load: " + ext);
            }
            amrex::RealBox rb({AMREX_D_DECL(0.0, 0.0, 0.0)},
                              {AMREX_D_DECL(amrex::Real(domain.length(0)), amrex::Real(domain.length(1)),
                                            amrex::Real(domain.length(2)))});
            amrex::Array<int, AMREX_SPACEDIM> periodic = {AMREX_D_DECL(1, 1, 1)};   // reference :306-309
            geom_full.define(domain, &rb, 0, periodic.data());
            mf_phase.FillBoundary(geom_full.periodicity());
        } catch (const std::exception& e) {
            amrex::Print() << "Error loading full domain data: " << e.what() << std::endl;
            amrex::Abort("Full domain data loading failed.");
        }

        // ---- REV study (reference :317-504): D_eff tensors of random cubic sub-volumes, each
        // treated as its own periodic box, one CSV row per (sample, size)
        {
            int rev_do_study = 0, rev_num_samples = 3, rev_write_plotfiles = 0, rev_verbose = 1;
            std::string rev_sizes_str = "32 64 96", rev_solver_str = "FlexGMRES", rev_results_filename = "rev_study_Deff.csv";
            amrex::ParmParse ppr("rev");
            ppr.query("do_study", rev_do_study);
            ppr.query("num_samples", rev_num_samples);
            ppr.query("sizes", rev_sizes_str);
            {   // unquoted `rev.sizes = 32 64 96` arrives as several values
                std::vector<std::string> all;
                if (ppr.queryarr("sizes", all) && all.size() > 1) {
                    rev_sizes_str.clear();
                    for (const auto& v : all) rev_sizes_str += v + " ";
                }
            }
            ppr.query("solver_type", rev_solver_str);
            ppr.query("results_file", rev_results_filename);
            ppr.query("write_plotfiles", rev_write_plotfiles);
            ppr.query("verbose", rev_verbose);
            if (rev_do_study && amrex::ParallelDescriptor::NProcs() > 1)
                amrex::Abort("rev.do_study runs on one rank (use b200.rev_workers to spread the sub-volumes over the GPUs).");
            if (rev_do_study) {
                if (verbose >= 1) {
                    amrex::Print() << "\n--- Starting REV Study (Homogenization Method) for Phase ID " << phase_id << " ---\n";
                    amrex::Print() << "  Number of samples per size: " << rev_num_samples << std::endl;
                    amrex::Print() << "  Target REV sizes: " << rev_sizes_str << std::endl;
                    amrex::Print() << "  REV Solver: " << rev_solver_str << std::endl;
                }
                std::vector<int> sizes;
                {
                    std::stringstream ss(rev_sizes_str);
                    int v;
                    while (ss >> v) sizes.push_back(v);
                }
                if (sizes.empty()) {
                    amrex::Warning("REV sizes string is empty or invalid. Skipping REV study.");
                } else {
                    std::ofstream csv((results_dir / rev_results_filename).string());
                    csv << "SampleNo,SeedX,SeedY,SeedZ,REV_Size_Target,ActualSizeX,ActualSizeY,ActualSizeZ,D_xx,D_yy,D_zz,D_xy,D_xz,D_yz\n";
                    std::mt19937 gen(amrex::ParallelDescriptor::MyProc() + 12345 + rev_num_samples);     // :341
                    const auto st_rev = static_cast<OpenImpala::EffectiveDiffusivityHypre::SolverType>(stringToSolverType(rev_solver_str));
                    // (1) the sub-volumes, drawn in the reference's order from its generator
                    struct RevJob { int sample; int target; amrex::IntVect seed_lo; amrex::Box bx; amrex::Real D[3][3]; };
                    std::vector<RevJob> jobs;
                    for (int s_idx = 0; s_idx < rev_num_samples; ++s_idx) {
                        for (int target : sizes) {
                            amrex::IntVect seed_lo;
                            for (int d = 0; d < 3; ++d) {                                               // :346-355
                                const int min_c = domain.smallEnd(d), max_c = domain.bigEnd(d) - (target - 1);
                                if (min_c > max_c || target > domain.length(d)) seed_lo[d] = domain.smallEnd(d);
                                else { std::uniform_int_distribution<> distr(min_c, max_c); seed_lo[d] = distr(gen); }
                            }
                            amrex::Box bx(seed_lo, seed_lo + amrex::IntVect(target - 1, target - 1, target - 1));
                            bx &= domain;
                            const int longside = bx.isEmpty() ? 0 : std::max(bx.length(0), std::max(bx.length(1), bx.length(2)));
                            if (bx.isEmpty() || longside < 8) {                                         // :361-369
                                if (rev_verbose >= 1)
                                    amrex::Warning("Skipping REV for sample " + std::to_string(s_idx + 1) + " target size " +
                                                   std::to_string(target) + " due to small/empty box after intersection");
                                continue;
                            }
                            RevJob job{s_idx + 1, target, seed_lo, bx, {}};
                            for (auto& row : job.D) for (auto& v : row) v = std::numeric_limits<amrex::Real>::quiet_NaN();
                            jobs.push_back(job);
                        }
                    }
                    // (2) the solves.  The sub-volumes are independent ("replicas"): b200.rev_workers = W runs them on
                    // W host threads, worker w on device w % (number of GPUs) -- several GPUs share the study, and
                    // several workers on one GPU overlap the launch-bound solves of small boxes (each solver object
                    // has its own stream).  W = 1 (default) is the reference's serial loop.
                    int rev_workers = 1;
                    {
                        amrex::ParmParse pp_b200("b200");
                        pp_b200.query("rev_workers", rev_workers);
                    }
                    rev_workers = std::max(1, std::min<int>(rev_workers, (int)jobs.size()));
                    int n_devices = 1;
                    if (rev_workers > 1 && (oi_device_count(&n_devices) != 0 || n_devices < 1)) n_devices = 1;
                    std::mutex print_mutex;
                    auto run_job = [&](RevJob& job) {
                        const amrex::Box& bx = job.bx;
                        if (rev_verbose >= 1) {
                            std::lock_guard<std::mutex> lk(print_mutex);
                            amrex::Print() << " REV Sample " << job.sample << ", Target Size " << job.target << ", Seed Lo (global): "
                                           << job.seed_lo << ", Actual REV Box (global): " << bx << std::endl;
                        }
                        // the sub-volume as its own periodic box with origin 0 (:377-420)
                        const amrex::Box rel(amrex::IntVect(0, 0, 0), bx.bigEnd() - bx.smallEnd());
                        amrex::Geometry geom_rev;
                        amrex::RealBox rb_rev({AMREX_D_DECL(0.0, 0.0, 0.0)},
                                              {AMREX_D_DECL(amrex::Real(rel.length(0)), amrex::Real(rel.length(1)), amrex::Real(rel.length(2)))});
                        amrex::Array<int, AMREX_SPACEDIM> per_rev = {AMREX_D_DECL(1, 1, 1)};
                        geom_rev.define(rel, &rb_rev, 0, per_rev.data());
                        amrex::BoxArray ba_rev(rel);
                        ba_rev.maxSize(box_size);
                        amrex::DistributionMapping dm_rev(ba_rev);
                        amrex::iMultiFab mf_rev(ba_rev, dm_rev, 1, 1);
                        mf_rev.setVal(0);
                        for (int k = rel.smallEnd(2); k <= rel.bigEnd(2); ++k)
                            for (int j = rel.smallEnd(1); j <= rel.bigEnd(1); ++j)
                                for (int i = rel.smallEnd(0); i <= rel.bigEnd(0); ++i)
                                    mf_rev(i, j, k, 0) = mf_phase(i + bx.smallEnd(0), j + bx.smallEnd(1), k + bx.smallEnd(2), 0);
                        mf_rev.FillBoundary(geom_rev.periodicity());
                        amrex::Real Dtmp[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
                        bool all_ok = true;
                        const long long n_rev = rel.numPts();
                        const OpenImpala::Direction dirs3[3] = {OpenImpala::Direction::X, OpenImpala::Direction::Y, OpenImpala::Direction::Z};
                        for (int c = 0; c < 3 && all_ok; ++c) {
                            // plotfiles of a sub-volume go to their own directory (reference :434-441)
                            const std::string rev_dir = (results_dir / ("REV_Sample" + std::to_string(job.sample) + "_Size" +
                                                                        std::to_string(bx.length(0)) + "_Dir" + std::to_string(c))).string();
                            OpenImpala::EffectiveDiffusivityHypre solver(geom_rev, ba_rev, dm_rev, mf_rev, phase_id, dirs3[c], st_rev,
                                                                         rev_dir, rev_verbose > 1 ? rev_verbose : 0, rev_write_plotfiles != 0);
                            if (!solver.solve()) {
                                all_ok = false;
                                if (rev_verbose >= 1) amrex::Print() << "    REV Chi solve FAILED for dir " << c << std::endl;
                                break;
                            }
                            amrex::Real sums[3];
                            long long n_active = 0;
                            solver.gradientSums(sums, n_active);
                            for (int r = 0; r < 3; ++r) Dtmp[r][c] = ((r == c ? (amrex::Real)n_active : 0.0) - sums[r]) / (amrex::Real)n_rev;
                        }
                        if (all_ok) for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) job.D[r][c] = Dtmp[r][c];
                    };
                    if (rev_workers == 1) {
                        for (RevJob& job : jobs) run_job(job);
                    } else {
                        if (verbose >= 1)
                            amrex::Print() << "  REV workers: " << rev_workers << " host threads over " << n_devices << " device(s)" << std::endl;
                        std::atomic<size_t> next{0};
                        std::vector<std::thread> pool;
                        for (int w = 0; w < rev_workers; ++w)
                            pool.emplace_back([&, w] {
                                OpenImpala::EffectiveDiffusivityHypre::setThreadDevice(w % n_devices);
                                for (size_t q = next++; q < jobs.size(); q = next++) run_job(jobs[q]);
                            });
                        for (auto& th : pool) th.join();
                    }
                    // (3) one CSV row per sub-volume, in the order they were drawn
                    for (const RevJob& job : jobs) {
                        const auto& D = job.D;
                        csv << job.sample << "," << job.seed_lo[0] << "," << job.seed_lo[1] << "," << job.seed_lo[2] << "," << job.target << ","
                            << job.bx.length(0) << "," << job.bx.length(1) << "," << job.bx.length(2) << "," << std::fixed << std::setprecision(8)
                            << D[0][0] << "," << D[1][1] << "," << D[2][2] << "," << D[0][1] << "," << D[0][2] << "," << D[1][2] << "\n";
                    }
                    csv.flush();
                }
            }
        }
        if (method != "flow_through" && method != "homogenization" && method != "skip_if_rev")
            amrex::Abort("Invalid calculation_method: '" + method + "'. Use homogenization, flow_through or skip_if_rev.");
        if (method == "skip_if_rev") {                      // reference :507: no full-domain calculation
            amrex::Print() << std::endl << "Total run time (seconds) = " << (amrex::second() - t_start) << std::endl;
            amrex::Finalize();
            return 0;
        }

        if (method == "homogenization") {                                           // reference :509-589
            if (verbose >= 1) amrex::Print() << "\n--- Effective Diffusivity via Homogenization (Full Domain) ---\n";
            amrex::MultiFab chi[3] = {amrex::MultiFab(ba, dm, 1, 1), amrex::MultiFab(ba, dm, 1, 1),
                                      amrex::MultiFab(ba, dm, 1, 1)};
            int check_host_tensor = 0;
            amrex::ParmParse pp_b200("b200");
            pp_b200.query("check_host_tensor", check_host_tensor);
            if (check_host_tensor && amrex::ParallelDescriptor::NProcs() > 1)
                amrex::Abort("b200.check_host_tensor needs the whole corrector fields on one rank");
            const auto st_eff = static_cast<OpenImpala::EffectiveDiffusivityHypre::SolverType>(stringToSolverType(solver_str));
            amrex::Real Deff[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
            const long long n_total = domain.numPts();
            bool all_converged = true;
            const OpenImpala::Direction dirs3[3] = {OpenImpala::Direction::X, OpenImpala::Direction::Y, OpenImpala::Direction::Z};
            for (int c = 0; c < 3; ++c) {
                const char* dc = c == 0 ? "X" : (c == 1 ? "Y" : "Z");
                if (verbose >= 1) amrex::Print() << "\n--- Solving for Full Domain chi_" << dc << " ---\n";
                OpenImpala::EffectiveDiffusivityHypre solver(geom_full, ba, dm, mf_phase, phase_id, dirs3[c], st_eff,
                                                             (results_dir / (std::string("FullDomain_chi_") + dc)).string(),
                                                             verbose, write_plotfile != 0);
                if (!solver.solve()) { all_converged = false; break; }              // :546-549
                amrex::Real sums[3];
                long long n_active = 0;
                solver.gradientSums(sums, n_active);
                for (int r = 0; r < 3; ++r)
                    Deff[r][c] = n_total > 0 ? ((r == c ? (amrex::Real)n_active : 0.0) - sums[r]) / (amrex::Real)n_total : 0.0;
                if (check_host_tensor) solver.getChiSolution(chi[c]);
            }
            if (all_converged) {
                if (check_host_tensor) {
                    amrex::iMultiFab active(ba, dm, 1, 0);
                    for (int k = domain.smallEnd(2); k <= domain.bigEnd(2); ++k)
                        for (int j = domain.smallEnd(1); j <= domain.bigEnd(1); ++j)
                            for (int i = domain.smallEnd(0); i <= domain.bigEnd(0); ++i)
                                active(i, j, k, 0) = (mf_phase(i, j, k, 0) == phase_id) ? 1 : 0;
                    amrex::Real Dh[3][3];
                    calculate_Deff_tensor_homogenization(Dh, chi[0], chi[1], chi[2], active, geom_full, verbose);
                    amrex::Real worst = 0.0;
                    for (int r = 0; r < 3; ++r)
                        for (int c = 0; c < 3; ++c) worst = std::max(worst, std::abs(Dh[r][c] - Deff[r][c]));
                    amrex::Print() << "Host tensor check: max |D_host - D_device| = " << std::scientific << worst << "\n";
                    if (!(worst <= 1e-10)) amrex::Abort("host / device D_eff tensors disagree");
                }
                amrex::Print() << "Full Domain Effective Diffusivity Tensor D_eff / D_material:\n";
                for (int r = 0; r < 3; ++r) {
                    amrex::Print() << "  [";
                    for (int c = 0; c < 3; ++c)
                        amrex::Print() << std::scientific << std::setprecision(8) << Deff[r][c] << (c == 2 ? "" : ", ");
                    amrex::Print() << "]\n";
                }
                // (the reference prints the tensor only; the file is an addition of this build)
                const std::filesystem::path out_path = results_dir / output_filename;
                std::ofstream out;
                if (amrex::ParallelDescriptor::IOProcessor()) out.open(out_path);
                if (out.is_open()) {
                    out << "# Effective Diffusivity Results (Homogenization Method)\n";
                    out << "# Input File: " << filename << "\n";
                    out << "# Analysis Phase ID: " << phase_id << "\n";
                    out << "# -----------------------------\n";
                    const char* ax = "xyz";
                    for (int r = 0; r < 3; ++r)
                        for (int c = 0; c < 3; ++c)
                            out << "Deff_" << ax[r] << ax[c] << ": " << std::scientific << std::setprecision(9) << Deff[r][c] << "\n";
                }
            } else {
                amrex::Print() << "Full domain D_eff calculation skipped due to chi_k non-convergence.\n";
            }
            amrex::Print() << std::endl << "Total run time (seconds) = " << (amrex::second() - t_start) << std::endl;
            amrex::Finalize();
            return 0;
        }

        if (verbose >= 1) amrex::Print() << "\n--- Full Domain Calculation: Tortuosity via Flow-Through ---\n";
        amrex::Real vlo = -1.0, vhi = 1.0;                                            // reference :605-611
        amrex::ParmParse pp_tort("tortuosity");
        pp_tort.query("vlo", vlo);
        pp_tort.query("vhi", vhi);

        if (verbose > 0) amrex::Print() << "Calculating Volume Fraction for Phase ID: " << phase_id << "\n";
        OpenImpala::VolumeFraction vf_calc(mf_phase, phase_id);
        long long phase_voxels = 0, total_voxels = 0;
        vf_calc.value(phase_voxels, total_voxels, false);
        const amrex::Real volume_fraction = total_voxels > 0 ? (amrex::Real)phase_voxels / (amrex::Real)total_voxels : 0.0;
        amrex::Print() << "  Volume Fraction = " << std::fixed << std::setprecision(8) << volume_fraction << "\n";

        std::map<std::string, amrex::Real> results;
        std::string direction_str;
        amrex::ParmParse pp;
        pp.get("direction", direction_str);
        {   // unquoted `direction = X Y Z` arrives as several values
            std::vector<std::string> all;
            if (pp.queryarr("direction", all) && all.size() > 1) {
                direction_str.clear();
                for (const auto& s : all) direction_str += s + " ";
            }
        }
        std::string upper = direction_str;
        std::transform(upper.begin(), upper.end(), upper.begin(), ::toupper);
        std::vector<OpenImpala::Direction> dirs;
        if (upper.find("ALL") != std::string::npos) {
            dirs = {OpenImpala::Direction::X, OpenImpala::Direction::Y, OpenImpala::Direction::Z};
        } else {
            std::stringstream ss(upper);
            std::string one;
            while (ss >> one) {
                if (one == "X") dirs.push_back(OpenImpala::Direction::X);
                else if (one == "Y") dirs.push_back(OpenImpala::Direction::Y);
                else if (one == "Z") dirs.push_back(OpenImpala::Direction::Z);
            }
        }
        if (dirs.empty()) amrex::Warning("No valid directions specified in 'direction' input. Skipping tortuosity calculation.");

        // The directions are independent solves: b200.dir_workers = W runs them on W host threads, worker w
        // on device w % (number of GPUs) (one solver object = one handle = one stream, so on a single GPU the
        // launch-bound solves of a small image overlap).  W = 1 (default) is the reference's serial loop.
        int dir_workers = 1;
        {
            amrex::ParmParse pp_b200("b200");
            pp_b200.query("dir_workers", dir_workers);
        }
        dir_workers = std::max(1, std::min<int>(dir_workers, (int)dirs.size()));
        if (amrex::ParallelDescriptor::NProcs() > 1) dir_workers = 1;       // the ranks already share every solve
        int n_devices_dir = 1;
        if (dir_workers > 1 && (oi_device_count(&n_devices_dir) != 0 || n_devices_dir < 1)) n_devices_dir = 1;
        std::vector<amrex::Real> taus(dirs.size(), std::numeric_limits<amrex::Real>::quiet_NaN());
        auto solve_direction = [&](size_t di) {
            const auto dir = dirs[di];
            const std::string dc = dir == OpenImpala::Direction::X ? "X" : dir == OpenImpala::Direction::Y ? "Y" : "Z";
            if (verbose >= 1) amrex::Print() << "\n--- Solving for Tortuosity in Direction: " << dc << " ---\n";
            amrex::Geometry geom_tort;                 // same box, NON-periodic (reference :671-677)
            {
                amrex::RealBox rb = geom_full.ProbDomain();
                amrex::Array<int, AMREX_SPACEDIM> np = {AMREX_D_DECL(0, 0, 0)};
                geom_tort.define(geom_full.Domain(), &rb, 0, np.data());
            }
            // b200.stream_upload = N (TIFF, HDF5 or RAW input): skip the int32 iMultiFab on the way to the GPU and
            // decode N planes at a time straight into the pinned staging buffers
            int stream_upload = 0;
            {
                amrex::ParmParse pp_b200("b200");
                pp_b200.query("stream_upload", stream_upload);
            }
            std::string ext_l = input.extension().string();
            std::transform(ext_l.begin(), ext_l.end(), ext_l.begin(), ::tolower);
            std::unique_ptr<OpenImpala::TortuosityHypre> solver_ptr;
            if (stream_upload > 0 && (ext_l == ".tif" || ext_l == ".tiff")) {
                OpenImpala::TiffReader reader(input.string());
                const double thr = threshold_val;
                solver_ptr = std::make_unique<OpenImpala::TortuosityHypre>(
                    geom_tort, ba, dm,
                    [&](int z0, int nz, unsigned char* out) { reader.thresholdPlanesU8(thr, 1, 0, z0, nz, out); },
                    stream_upload, volume_fraction, phase_id, dir, stringToSolverType(solver_str), results_path, vlo, vhi,
                    verbose, write_plotfile != 0);
                if (verbose >= 1) amrex::Print() << "  (phase field streamed from the TIFF in chunks of " << stream_upload << " planes)\n";
            } else if (stream_upload > 0 && (ext_l == ".h5" || ext_l == ".hdf5")) {
                OpenImpala::HDF5Reader reader(input.string(), hdf5_dataset);
                const double thr = threshold_val;
                solver_ptr = std::make_unique<OpenImpala::TortuosityHypre>(
                    geom_tort, ba, dm,
                    [&](int z0, int nz, unsigned char* out) { reader.thresholdPlanesU8(thr, 1, 0, z0, nz, out); },
                    stream_upload, volume_fraction, phase_id, dir, stringToSolverType(solver_str), results_path, vlo, vhi,
                    verbose, write_plotfile != 0);
                if (verbose >= 1) amrex::Print() << "  (phase field streamed from the HDF5 dataset in chunks of " << stream_upload << " planes)\n";
            } else if (stream_upload > 0 && ext_l == ".raw") {
                OpenImpala::RawReader reader(input.string(), raw_w, raw_h, raw_d, rawTypeFromString(raw_type));
                const double thr = threshold_val;
                solver_ptr = std::make_unique<OpenImpala::TortuosityHypre>(
                    geom_tort, ba, dm,
                    [&](int z0, int nz, unsigned char* out) { reader.thresholdPlanesU8(thr, 1, 0, z0, nz, out); },
                    stream_upload, volume_fraction, phase_id, dir, stringToSolverType(solver_str), results_path, vlo, vhi,
                    verbose, write_plotfile != 0);
                if (verbose >= 1) amrex::Print() << "  (phase field streamed from the RAW volume in chunks of " << stream_upload << " planes)\n";
            } else {
                solver_ptr = std::make_unique<OpenImpala::TortuosityHypre>(
                    geom_tort, ba, dm, mf_phase, volume_fraction, phase_id, dir, stringToSolverType(solver_str),
                    results_path, vlo, vhi, verbose, write_plotfile != 0);
            }
            OpenImpala::TortuosityHypre& solver = *solver_ptr;
            const amrex::Real tau = solver.value();
            taus[di] = tau;
            amrex::Print() << "  >>> Calculated Tortuosity (" << dc << "): " << std::fixed << std::setprecision(8) << tau << " <<<\n";
        };
        if (dir_workers == 1) {
            for (size_t di = 0; di < dirs.size(); ++di) solve_direction(di);
        } else {
            if (verbose >= 1)
                amrex::Print() << "  Direction workers: " << dir_workers << " host threads over " << n_devices_dir << " device(s)" << std::endl;
            std::atomic<size_t> next_dir{0};
            std::vector<std::thread> pool;
            for (int w = 0; w < dir_workers; ++w)
                pool.emplace_back([&, w] {
                    OpenImpala::TortuosityHypre::setThreadDevice(w % n_devices_dir);
                    for (size_t q = next_dir++; q < dirs.size(); q = next_dir++) solve_direction(q);
                });
            for (auto& th : pool) th.join();
        }
        for (size_t di = 0; di < dirs.size(); ++di) {
            const auto dir = dirs[di];
            results[std::string("Tortuosity_") + (dir == OpenImpala::Direction::X ? "X" : dir == OpenImpala::Direction::Y ? "Y" : "Z")] = taus[di];
        }

        const std::filesystem::path out_path = results_dir / output_filename;
        amrex::Print() << "\nWriting final results to: " << out_path << "\n";
        std::ofstream out;
        if (amrex::ParallelDescriptor::IOProcessor()) out.open(out_path);
        if (!amrex::ParallelDescriptor::IOProcessor()) {
            // (only the IO processor writes, as with the reference's ParallelDescriptor::IOProcessor() guard)
        } else if (out.is_open()) {
            out << "# Tortuosity Calculation Results (Flow-Through Method)\n";
            out << "# Input File: " << filename << "\n";
            out << "# Analysis Phase ID: " << phase_id << "\n";
            out << "# -----------------------------\n";
            out << "VolumeFraction: " << std::fixed << std::setprecision(9) << volume_fraction << "\n";
            for (const auto& kv : results) out << kv.first << ": " << std::fixed << std::setprecision(9) << kv.second << "\n";
        } else {
            amrex::Warning("Could not open output file for writing: " + out_path.string());
        }
        amrex::Print() << std::endl << "Total run time (seconds) = " << (amrex::second() - t_start) << std::endl;
    }
    amrex::Finalize();
    return 0;
}
