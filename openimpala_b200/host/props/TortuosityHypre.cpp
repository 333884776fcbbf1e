// TortuosityHypre on the B200 library.
//
// Control flow and result conventions follow the reference class
// (src/props/TortuosityHypre.cpp): all heavy setup in the constructor (:100-191),
// value() solves once and caches (:761-891), numerical failures come back
// in-band as NaN / +Inf, library failures abort.  The numerical work -- mask,
// matrix-free operator, Krylov + multigrid, flux planes -- runs on the GPU
// through include/openimpala_b200.h.
#include "TortuosityHypre.H"

#include <algorithm>
#include <cmath>
#include <iomanip>
#include <vector>

#include <AMReX_ParallelDescriptor.H>
#include <AMReX_MultiFab.H>
#include <AMReX_ParmParse.H>
#include <AMReX_PlotFileUtil.H>
#include <AMReX_Print.H>

#include <openimpala_b200.h>

namespace {
constexpr amrex::Real tiny_flux_threshold = 1.e-15;     // TortuosityHypre.cpp:63

void oi_check(int rc, const char* what) {
    if (rc != OI_OK)
        amrex::Abort(std::string("openimpala_b200 error in ") + what + ": " + oi_last_error() +
                     " - Error Code: " + std::to_string(rc));
}
}  // namespace

namespace OpenImpala {

amrex::Array<int, AMREX_SPACEDIM> TortuosityHypre::loV(const amrex::Box& b) {
    return {b.smallEnd(0), b.smallEnd(1), b.smallEnd(2)};
}
amrex::Array<int, AMREX_SPACEDIM> TortuosityHypre::hiV(const amrex::Box& b) {
    return {b.bigEnd(0), b.bigEnd(1), b.bigEnd(2)};
}

namespace { thread_local int t_thread_device = -1; }

TortuosityHypre::TortuosityHypre(const amrex::Geometry& geom, const amrex::BoxArray& ba,
                                 const amrex::DistributionMapping& dm,
                                 const amrex::iMultiFab& mf_phase_input, const amrex::Real vf,
                                 const int phase, const OpenImpala::Direction dir, const SolverType st,
                                 const std::string& resultspath, const amrex::Real vlo,
                                 const amrex::Real vhi, int verbose, bool write_plotfile)
    : m_solvertype(st), m_resultspath(resultspath), m_phase(phase), m_dir(dir), m_vlo(vlo), m_vhi(vhi),
      m_verbose(verbose), m_vf(vf), m_write_plotfile(write_plotfile), m_geom(geom), m_ba(ba), m_dm(dm) {
    AMREX_ALWAYS_ASSERT_WITH_MESSAGE(mf_phase_input.nGrow() >= 1, "Phase fab needs ghost cells");
    // the phase field is deep-copied (valid cells only: every out-of-domain neighbour is
    // inactive, which is what the reference's non-periodic mask ghosts amount to, :309, :522)
    initialize([&](oi_solver* h) {
        const std::vector<int> cells = mf_phase_input.validCopy(0);
        oi_check(oi_set_phase_i32(h, cells.data()), "oi_set_phase_i32");
        if (m_write_plotfile) m_plot_phase = cells;
    });
}

TortuosityHypre::TortuosityHypre(const amrex::Geometry& geom, const amrex::BoxArray& ba,
                                 const amrex::DistributionMapping& dm, const PhaseStream& phase_stream,
                                 int planes_per_chunk, const amrex::Real vf, const int phase,
                                 const OpenImpala::Direction dir, const SolverType st,
                                 const std::string& resultspath, const amrex::Real vlo, const amrex::Real vhi,
                                 int verbose, bool write_plotfile)
    : m_solvertype(st), m_resultspath(resultspath), m_phase(phase), m_dir(dir), m_vlo(vlo), m_vhi(vhi),
      m_verbose(verbose), m_vf(vf), m_write_plotfile(write_plotfile), m_geom(geom), m_ba(ba), m_dm(dm) {
    // the phase field arrives in z-chunks decoded straight into pinned staging buffers;
    // decoding of chunk k+1 overlaps the upload of chunk k
    initialize([&](oi_solver* h) {
        const amrex::Box lb = m_ba.localBox();                         // planes of this rank's slab
        const int nz = lb.length(2), zb = lb.smallEnd(2) - m_geom.Domain().smallEnd(2);
        const int chunk = std::max(1, std::min(planes_per_chunk, nz));
        oi_check(oi_phase_stream_begin(h, chunk), "oi_phase_stream_begin");
        int which = 0;
        for (int z0 = 0; z0 < nz; z0 += chunk, which ^= 1) {
            const int n = std::min(chunk, nz - z0);
            uint8_t* buf = nullptr;
            oi_check(oi_phase_stream_buffer(h, which, &buf), "oi_phase_stream_buffer");
            phase_stream(zb + z0, n, buf);
            if (m_write_plotfile)
                m_plot_phase.insert(m_plot_phase.end(), buf,
                                    buf + (size_t)n * (size_t)m_geom.Domain().length(0) * (size_t)m_geom.Domain().length(1));
            oi_check(oi_phase_stream_submit(h, which, z0, n), "oi_phase_stream_submit");
        }
        oi_check(oi_phase_stream_end(h), "oi_phase_stream_end");
    });
}

void TortuosityHypre::initialize(const std::function<void(oi_solver*)>& upload_phase) {
    const bool io = amrex::ParallelDescriptor::IOProcessor();
    if (m_verbose > 0 && io) {
        amrex::Print() << "TortuosityHypre: Initializing..." << std::endl;
        amrex::Print() << "  Original Total VF (Phase " << m_phase << "): " << m_vf << std::endl;
    }
    // hard-coded defaults, overridable from the inputs file (reference :142-151)
    m_eps = 1e-9;
    m_maxiter = 200;
    amrex::ParmParse pp("hypre");
    pp.query("eps", m_eps);
    pp.query("maxiter", m_maxiter);
    amrex::ParmParse pp_tort("tortuosity");
    pp_tort.query("verbose", m_verbose);
    if (m_verbose > 0 && io) {
        amrex::Print() << "  HYPRE Params: eps=" << m_eps << ", maxiter=" << m_maxiter << std::endl;
        amrex::Print() << "  Class Verbose Level: " << m_verbose << std::endl;
        amrex::Print() << "  Write Plotfile Flag: " << m_write_plotfile << std::endl;
    }
    AMREX_ALWAYS_ASSERT_WITH_MESSAGE(m_vf >= 0.0 && m_vf <= 1.0, "Original Volume fraction must be between 0 and 1");
    AMREX_ALWAYS_ASSERT_WITH_MESSAGE(m_eps > 0.0, "Solver tolerance (eps) must be positive");
    AMREX_ALWAYS_ASSERT_WITH_MESSAGE(m_maxiter > 0, "Solver max iterations must be positive");

    // device handle
    const amrex::Box& domain = m_geom.Domain();
    oi_params p;
    oi_default_params(&p);
    p.nx = domain.length(0); p.ny = domain.length(1); p.nz = domain.length(2);
    // this rank's z-slab of the BoxArray (reference: every rank owns the boxes its DistributionMapping
    // assigns to it, TortuosityHypre.H:68-80); one rank = the whole box
    const amrex::Box local = m_ba.localBox();
    p.z_begin = local.smallEnd(2) - domain.smallEnd(2); p.nz_local = local.length(2);
    p.comm = amrex::ParallelDescriptor::Communicator();
    if (amrex::ParallelDescriptor::NProcs() > 1) {
        p.device = amrex::ParallelDescriptor::LocalDevice();
        if (m_write_plotfile) amrex::Abort("TortuosityHypre: write_plotfile is not supported on more than one rank in this build");
    }
    p.direction = static_cast<int>(m_dir);
    p.phase_id = m_phase;
    p.vlo = m_vlo; p.vhi = m_vhi;
    for (int d = 0; d < 3; ++d) p.dx[d] = m_geom.CellSize(d);
    p.eps = m_eps; p.maxiter = m_maxiter; p.verbose = m_verbose;
    amrex::ParmParse pp_b200("b200");            // extras of this implementation, all optional
    if (amrex::ParallelDescriptor::NProcs() <= 1) pp_b200.query("device", p.device);
    if (t_thread_device >= 0) p.device = t_thread_device;             // setThreadDevice()
    pp_b200.query("mg_degree", p.mg_degree);
    pp_b200.query("flux_polish", p.flux_polish);
    pp_b200.query("stencil_variant", p.stencil_variant);
    pp_b200.query("precond", p.precond);
    oi_check(oi_create(&m_solver, &p), "oi_create");
    upload_phase(m_solver);

    int num_remspot_passes = 0;                  // reference :254-263
    pp_tort.query("remspot_passes", num_remspot_passes);
    if (num_remspot_passes > 0) oi_check(oi_remspot(m_solver, num_remspot_passes), "oi_remspot");
    else if (m_verbose > 1 && io) amrex::Print() << "  Skipping tortuosity_remspot filter (remspot_passes <= 0)." << std::endl;

    if (m_verbose > 0 && io) amrex::Print() << "TortuosityHypre: Generating activity mask via boundary search..." << std::endl;
    int64_t num_active = 0;
    oi_check(oi_build_mask(m_solver, &num_active), "oi_build_mask");
    const long long total_cells = domain.numPts();
    m_active_vf = (total_cells > 0) ? static_cast<amrex::Real>(num_active) / total_cells : 0.0;   // :552-553
    if (m_verbose > 0 && io)
        amrex::Print() << "  Active Volume Fraction (percolating phase " << m_phase << "): " << m_active_vf << std::endl;

    if (m_active_vf <= std::numeric_limits<amrex::Real>::epsilon()) {                             // :170-178
        if (m_verbose >= 0 && io)
            amrex::Print() << "WARNING: Active volume fraction is zero. Skipping matrix setup and solve." << std::endl;
        m_first_call = false;
        m_value = std::numeric_limits<amrex::Real>::quiet_NaN();
        return;
    }
    if (m_verbose > 0 && io) amrex::Print() << "TortuosityHypre: Initialization complete." << std::endl;
}

void TortuosityHypre::setThreadDevice(int device) { t_thread_device = device; }

TortuosityHypre::~TortuosityHypre() {
    if (m_solver) oi_destroy(m_solver);
    m_solver = nullptr;
}

bool TortuosityHypre::solve() {
    m_num_iterations = -1;
    m_final_res_norm = std::numeric_limits<amrex::Real>::quiet_NaN();
    m_converged = false;
    // The reference aborts on every SolverType but FlexGMRES (:695-697).  All
    // seven names are accepted here and run the same MG-preconditioned CG: the
    // converged answer does not depend on the Krylov method.
    oi_solve_info info;
    oi_check(oi_solve(m_solver, &info), "oi_solve");
    m_num_iterations = info.iterations;
    m_final_res_norm = info.rel_residual;
    m_converged = !(std::isnan(m_final_res_norm) || std::isinf(m_final_res_norm));                 // :687-688
    m_converged = m_converged && (m_final_res_norm >= 0.0) && info.converged;
    const bool io = amrex::ParallelDescriptor::IOProcessor();
    if (!m_converged && m_verbose >= 0) amrex::Warning("Krylov solver did not converge within tolerance!");
    if (m_verbose > 0 && io) {
        amrex::Print() << "  HYPRE Solver iterations: " << m_num_iterations << std::endl;
        amrex::Print() << "  HYPRE Final Relative Residual Norm: " << std::scientific << m_final_res_norm
                       << std::defaultfloat << std::endl;
        amrex::Print() << "  Solver Converged Status: " << (m_converged ? "Yes" : "No") << std::endl;
        amrex::Print() << "  Device time: solve " << info.solve_ms << " ms, multigrid setup " << info.setup_ms << " ms" << std::endl;
    }
    if (m_write_plotfile && m_converged) writeSolutionPlotfile();                                  // :710-745
    else if (m_write_plotfile && m_verbose >= 0 && io)
        amrex::Warning("Skipping plotfile write because solver did not converge.");               // :746-750
    return m_converged;
}

// <resultspath>/tortuosity_solution_<dir> with solution_potential, phase_id and active_mask,
// as the reference writes it (TortuosityHypre.cpp:710-745).  phase_id is the field handed to
// the constructor (the reference plots it after the optional remspot filter).
void TortuosityHypre::writeSolutionPlotfile() {
    const bool io = amrex::ParallelDescriptor::IOProcessor();
    if (m_verbose > 0 && io) amrex::Print() << "  Writing solution plotfile..." << std::endl;
    const amrex::Box& domain = m_geom.Domain();
    const size_t n = (size_t)domain.numPts();
    std::vector<double> x(n);
    std::vector<uint8_t> mask(n);
    oi_check(oi_get_solution(m_solver, x.data()), "oi_get_solution");
    oi_check(oi_get_mask_u8(m_solver, mask.data()), "oi_get_mask_u8");
    amrex::MultiFab mf_plot(m_ba, m_dm, 3, 0);
    size_t q = 0;
    for (int k = domain.smallEnd(2); k <= domain.bigEnd(2); ++k)
        for (int j = domain.smallEnd(1); j <= domain.bigEnd(1); ++j)
            for (int i = domain.smallEnd(0); i <= domain.bigEnd(0); ++i, ++q) {
                mf_plot(i, j, k, 0) = x[q];
                mf_plot(i, j, k, 1) = q < m_plot_phase.size() ? amrex::Real(m_plot_phase[q]) : 0.0;
                mf_plot(i, j, k, 2) = amrex::Real(mask[q]);
            }
    const std::string plotfilename = m_resultspath + "/tortuosity_solution_" + std::to_string(static_cast<int>(m_dir));
    const amrex::Vector<std::string> varnames = {"solution_potential", "phase_id", "active_mask"};
    amrex::WriteSingleLevelPlotfile(plotfilename, mf_plot, varnames, m_geom, 0.0, 0);
    if (m_verbose > 0 && io) amrex::Print() << "  Plotfile written to " << plotfilename << std::endl;
}

void TortuosityHypre::global_fluxes() {
    int64_t n_in = 0, n_out = 0;
    oi_check(oi_fluxes(m_solver, &m_flux_in, &m_flux_out, &n_in, &n_out), "oi_fluxes");           // :1000-1134
    if (m_verbose > 1 && amrex::ParallelDescriptor::IOProcessor())
        amrex::Print() << "  Active boundary cell counts: In=" << n_in << ", Out=" << n_out << "\n";
}

amrex::Real TortuosityHypre::value(const bool refresh) {
    const amrex::Real eps = std::numeric_limits<amrex::Real>::epsilon();
    const amrex::Real nan = std::numeric_limits<amrex::Real>::quiet_NaN();
    const amrex::Real inf = std::numeric_limits<amrex::Real>::infinity();
    const bool io = amrex::ParallelDescriptor::IOProcessor();
    if (m_active_vf <= eps && !m_first_call) return nan;                                          // :764-766
    if (m_first_call || refresh) {
        if (m_active_vf <= eps) {                                                                  // :770-777
            if (m_verbose >= 0 && io) amrex::Print() << "WARNING: Active volume fraction is zero. Tortuosity is NaN or Inf." << std::endl;
            m_value = nan; m_first_call = false;
            return m_value;
        }
        if (m_verbose > 0 && io) amrex::Print() << "Calculating Tortuosity (solve required)..." << std::endl;
        if (!solve()) {                                                                            // :782-787
            if (m_verbose >= 0 && io) amrex::Print() << "WARNING: Solver did not converge or failed. Tortuosity calculation skipped, returning NaN." << std::endl;
            m_value = nan; m_first_call = false;
            return m_value;
        }
        global_fluxes();
        // flux conservation gate, :794-823
        constexpr amrex::Real flux_tol = 1.0e-6;
        const amrex::Real mag_in = std::abs(m_flux_in), mag_out = std::abs(m_flux_out);
        const amrex::Real avg = 0.5 * (mag_in + mag_out);
        amrex::Real rel_diff = 0.0;
        bool conserved = true;
        if (avg > tiny_flux_threshold) {
            rel_diff = std::abs(mag_in - mag_out) / avg;
            conserved = !(rel_diff > flux_tol);
        }
        if (m_verbose > 0 && io) {
            amrex::Print() << "  Flux Conservation Check (|in|-|out|) / avg(|in|,|out|):\n";
            amrex::Print() << "    Flux In  = " << std::fixed << std::setprecision(8) << m_flux_in << "\n";
            amrex::Print() << "    Flux Out = " << std::fixed << std::setprecision(8) << m_flux_out << "\n";
            amrex::Print() << "    Relative Difference = " << std::scientific << rel_diff << std::defaultfloat
                           << " (Tolerance = " << flux_tol << ")\n";
            if (conserved) amrex::Print() << "    Conservation Check Status: PASS\n";
            else amrex::Warning("Flux conservation check failed!");
        }
        if (!conserved) {
            if (m_verbose >= 0 && io) amrex::Print() << "WARNING: Flux not conserved. Tortuosity calculation skipped, returning NaN." << std::endl;
            m_value = nan;
        } else {
            const int d = static_cast<int>(m_dir);
            const amrex::Real L = m_geom.ProbLength(d);                                            // :831-840
            const amrex::Real A = d == 0 ? m_geom.ProbLength(1) * m_geom.ProbLength(2)
                                : d == 1 ? m_geom.ProbLength(0) * m_geom.ProbLength(2)
                                         : m_geom.ProbLength(0) * m_geom.ProbLength(1);
            const amrex::Real gradPhi = (m_vhi - m_vlo) / L;
            amrex::Real Deff = 0.0;
            if (avg < tiny_flux_threshold) {                                                       // :846-851
                m_value = (m_active_vf > eps) ? inf : nan;
            } else if (m_active_vf <= eps) {                                                       // :854-858
                m_value = nan;
            } else if (std::abs(gradPhi) < tiny_flux_threshold) {                                  // :860-864
                m_value = inf;
            } else {
                Deff = (avg / A) / std::abs(gradPhi);                                              // :868
                m_value = (std::abs(Deff) < tiny_flux_threshold) ? inf : m_active_vf / Deff;       // :869-876
            }
            if (m_verbose > 0 && io) {
                amrex::Print() << "  Calculation Details: ActiveVf=" << m_active_vf << ", L=" << L << ", A=" << A
                               << ", gradPhi=" << gradPhi << ", AvgFluxMag=" << avg << ", Deff=" << Deff << std::endl;
                amrex::Print() << "  Calculated Tortuosity (using Active Vf): " << m_value << std::endl;
            }
        }
    }
    m_first_call = false;
    return m_value;
}

bool TortuosityHypre::checkMatrixProperties() {
    if (m_active_vf <= std::numeric_limits<amrex::Real>::epsilon()) return true;   // nothing was assembled
    int32_t ok = 0;
    oi_check(oi_check_matrix_properties(m_solver, &ok), "oi_check_matrix_properties");
    if (m_verbose > 0 && amrex::ParallelDescriptor::IOProcessor())
        amrex::Print() << (ok ? "TortuosityHypre: Matrix/vector property checks passed."
                              : "TortuosityHypre: Matrix/vector property checks FAILED.") << std::endl;
    return ok != 0;
}

}  // namespace OpenImpala
