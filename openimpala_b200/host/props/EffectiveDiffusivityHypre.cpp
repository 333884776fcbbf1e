// EffectiveDiffusivityHypre on the B200 library.
//
// Control flow follows the reference class (src/props/EffectiveDiffusivityHypre.cpp):
// mask and system set up in the constructor (:104-203), solve() runs the Krylov
// solve and reports convergence in-band (:543-676), getChiSolution() hands the
// corrector back with periodic ghosts (:678-745).  The numerical work runs on the
// GPU through include/openimpala_b200.h.
#include "EffectiveDiffusivityHypre.H"

#include <cmath>
#include <vector>

#include <AMReX_ParallelDescriptor.H>
#include <AMReX_ParmParse.H>
#include <AMReX_PlotFileUtil.H>
#include <AMReX_Print.H>

#include <openimpala_b200.h>

namespace {
void oi_check(int rc, const char* what) {
    if (rc != OI_OK)
        amrex::Abort(std::string("openimpala_b200 error in ") + what + ": " + oi_last_error() +
                     " - Error Code: " + std::to_string(rc));
}
}  // namespace

namespace OpenImpala {

amrex::Array<int, AMREX_SPACEDIM> EffectiveDiffusivityHypre::loV(const amrex::Box& b) {
    return {b.smallEnd(0), b.smallEnd(1), b.smallEnd(2)};
}
amrex::Array<int, AMREX_SPACEDIM> EffectiveDiffusivityHypre::hiV(const amrex::Box& b) {
    return {b.bigEnd(0), b.bigEnd(1), b.bigEnd(2)};
}

namespace { thread_local int t_thread_device = -1; }

EffectiveDiffusivityHypre::EffectiveDiffusivityHypre(const amrex::Geometry& geom, const amrex::BoxArray& ba,
                                                     const amrex::DistributionMapping& dm,
                                                     const amrex::iMultiFab& mf_phase_input, const int phase_id,
                                                     const OpenImpala::Direction dir_of_chi_k,
                                                     const SolverType solver_type, const std::string& resultspath,
                                                     int verbose_level, bool write_plotfile_flag)
    : m_solvertype(solver_type), m_resultspath(resultspath), m_phase_id(phase_id), m_dir_solve(dir_of_chi_k),
      m_eps(1e-9), m_maxiter(1000), m_verbose(verbose_level), m_write_plotfile(write_plotfile_flag),
      m_geom(geom), m_ba(ba), m_dm(dm), m_mf_active_mask(ba, dm, 1, 1) {
    const bool io = amrex::ParallelDescriptor::IOProcessor();
    if (m_verbose > 0 && io)
        amrex::Print() << "EffectiveDiffusivityHypre: Initializing for chi_k in direction "
                       << static_cast<int>(m_dir_solve) << "..." << std::endl;
    amrex::ParmParse pp_hypre("hypre");                                   // reference :154-159
    pp_hypre.query("eps", m_eps);
    pp_hypre.query("maxiter", m_maxiter);
    AMREX_ALWAYS_ASSERT_WITH_MESSAGE(m_eps > 0.0, "Solver tolerance (eps) must be positive");
    AMREX_ALWAYS_ASSERT_WITH_MESSAGE(m_maxiter > 0, "Solver max iterations must be positive");
    for (int d = 0; d < AMREX_SPACEDIM; ++d)
        AMREX_ALWAYS_ASSERT_WITH_MESSAGE(m_geom.CellSize(d) > 0.0, "Cell size must be positive.");
    for (int d = 0; d < AMREX_SPACEDIM; ++d)
        if (!m_geom.isPeriodic(d))
            amrex::Abort("EffectiveDiffusivityHypre: this build solves the cell problem on a fully periodic "
                         "geometry (Diffusion.cpp:306-308); a non-periodic direction was requested.");

    // this rank's z-slab of the BoxArray (reference: SPMD over MPI ranks, EffectiveDiffusivityHypre.H:55-63); one rank =
    // the whole box.  The periodic wrap in z between the first and the last slab is the library's (peer / NCCL halo).
    const int n_ranks = amrex::ParallelDescriptor::NProcs();
    if (n_ranks > 1 && m_write_plotfile)
        amrex::Abort("EffectiveDiffusivityHypre: write_plotfile is not supported on more than one rank in this build");
    // generateActiveMask (:213-330): phase == phase_id, ghosts by periodicity
    const amrex::Box& domain = m_geom.Domain();
    const amrex::Box local = m_ba.localBox();
    m_mf_active_mask.setVal(0);
    for (int k = local.smallEnd(2); k <= local.bigEnd(2); ++k)
        for (int j = local.smallEnd(1); j <= local.bigEnd(1); ++j)
            for (int i = local.smallEnd(0); i <= local.bigEnd(0); ++i)
                m_mf_active_mask(i, j, k, 0) = (mf_phase_input(i, j, k, 0) == m_phase_id) ? 1 : 0;
    m_mf_active_mask.FillBoundary(m_geom.periodicity());

    oi_params p;
    oi_default_params(&p);
    p.problem = OI_PROBLEM_CELL;
    p.nx = domain.length(0); p.ny = domain.length(1); p.nz = domain.length(2);
    p.z_begin = local.smallEnd(2) - domain.smallEnd(2); p.nz_local = local.length(2);
    p.comm = amrex::ParallelDescriptor::Communicator();
    if (n_ranks > 1) p.device = amrex::ParallelDescriptor::LocalDevice();
    p.direction = static_cast<int>(m_dir_solve);
    p.phase_id = m_phase_id;
    for (int d = 0; d < 3; ++d) p.dx[d] = m_geom.CellSize(d);
    p.eps = m_eps; p.maxiter = m_maxiter; p.verbose = m_verbose;
    amrex::ParmParse pp_b200("b200");
    if (n_ranks <= 1) pp_b200.query("device", p.device);
    if (t_thread_device >= 0) p.device = t_thread_device;                 // setThreadDevice()
    pp_b200.query("mg_degree", p.mg_degree);
    pp_b200.query("stencil_variant", p.stencil_variant);
    pp_b200.query("precond", p.precond);
    oi_check(oi_create(&m_solver, &p), "oi_create");
    {
        const std::vector<int> cells = mf_phase_input.validCopy(0);
        oi_check(oi_set_phase_i32(m_solver, cells.data()), "oi_set_phase_i32");
    }
    int64_t n_active = 0;
    oi_check(oi_build_mask(m_solver, &n_active), "oi_build_mask");
    m_num_active = n_active;
    if (m_verbose > 0 && io)
        amrex::Print() << "  Active mask generated. Number of active cells (Manually summed): " << m_num_active << std::endl;
    if (m_num_active == 0) {                                              // :186-196
        if (m_verbose >= 0 && io)
            amrex::Print() << "WARNING: No active cells found (manual sum) for phase_id " << m_phase_id
                           << ". HYPRE setup will be skipped." << std::endl;
        m_converged = true;
        m_num_iterations = 0;
        m_final_res_norm = 0.0;
        return;
    }
    if (m_verbose > 0 && io) amrex::Print() << "EffectiveDiffusivityHypre: Initialization complete." << std::endl;
}

void EffectiveDiffusivityHypre::setThreadDevice(int device) { t_thread_device = device; }

EffectiveDiffusivityHypre::~EffectiveDiffusivityHypre() {
    if (m_solver) oi_destroy(m_solver);
    m_solver = nullptr;
}

bool EffectiveDiffusivityHypre::solve() {
    const bool io = amrex::ParallelDescriptor::IOProcessor();
    if (m_num_active == 0) {                                              // :559-572
        if (m_verbose >= 0 && io)
            amrex::Print() << "EffectiveDiffusivityHypre::solve: Skipping HYPRE solve as no active cells were found "
                              "(manual sum) for phase " << m_phase_id << std::endl;
        m_converged = true;
        m_num_iterations = 0;
        m_final_res_norm = 0.0;
        return m_converged;
    }
    if (m_solvertype != SolverType::FlexGMRES)                            // :616-619
        amrex::Abort("Unsupported solver type requested in EffectiveDiffusivityHypre::solve: " +
                     std::to_string(static_cast<int>(m_solvertype)));
    m_num_iterations = -1;
    m_final_res_norm = std::numeric_limits<amrex::Real>::quiet_NaN();
    m_converged = false;
    oi_solve_info info;
    oi_check(oi_solve(m_solver, &info), "oi_solve");
    m_num_iterations = info.iterations;
    m_final_res_norm = info.rel_residual;
    m_converged = !(std::isnan(m_final_res_norm) || std::isinf(m_final_res_norm));      // :607-608
    m_converged = m_converged && (m_final_res_norm >= 0.0) && info.converged;
    if (!m_converged && m_verbose >= 0) amrex::Warning("Cell-problem solver did not converge within tolerance!");
    if (m_verbose > 0 && io) {
        amrex::Print() << "  HYPRE Solver Iterations: " << m_num_iterations << std::endl;
        amrex::Print() << "  HYPRE Final Relative Residual Norm: " << std::scientific << m_final_res_norm
                       << std::defaultfloat << std::endl;
        amrex::Print() << "  Solver Converged Status: " << (m_converged ? "Yes" : "No") << std::endl;
    }
    // <resultspath>/effdiff_chi_dir<k> with chi_k and the mask the solver used (:648-685)
    if (m_write_plotfile && m_converged) {
        if (m_verbose > 0 && io)
            amrex::Print() << "  Writing solution plotfile for chi_k in direction " << static_cast<int>(m_dir_solve)
                           << "..." << std::endl;
        amrex::MultiFab mf_plot(m_ba, m_dm, 2, 0);
        amrex::MultiFab chi(m_ba, m_dm, 1, 0);
        getChiSolution(chi);
        const amrex::Box& domain = m_geom.Domain();
        for (int k = domain.smallEnd(2); k <= domain.bigEnd(2); ++k)
            for (int j = domain.smallEnd(1); j <= domain.bigEnd(1); ++j)
                for (int i = domain.smallEnd(0); i <= domain.bigEnd(0); ++i) {
                    mf_plot(i, j, k, 0) = chi(i, j, k, 0);
                    mf_plot(i, j, k, 1) = amrex::Real(m_mf_active_mask(i, j, k, 0));
                }
        const std::string full_plot_path = m_resultspath + "/effdiff_chi_dir" + std::to_string(static_cast<int>(m_dir_solve));
        const amrex::Vector<std::string> varnames = {"chi_k", "active_mask_from_solver"};
        amrex::WriteSingleLevelPlotfile(full_plot_path, mf_plot, varnames, m_geom, 0.0, 0);
        if (m_verbose > 0 && io) amrex::Print() << "  Plotfile written to " << full_plot_path << std::endl;
    } else if (m_write_plotfile && m_verbose >= 0 && io) {
        amrex::Warning("Skipping plotfile write for chi_k because solver did not converge and had active cells.");
    }
    return m_converged;
}

void EffectiveDiffusivityHypre::getChiSolution(amrex::MultiFab& chi_field) {
    chi_field.setVal(0.0);
    if (m_solver && m_converged && m_num_active > 0) {                    // :631-637: zero if not converged
        // this rank's slab of the corrector field (the whole box on one rank)
        const amrex::Box local = m_ba.localBox();
        std::vector<double> x((size_t)local.numPts());
        oi_check(oi_get_solution(m_solver, x.data()), "oi_get_solution");
        size_t n = 0;
        for (int k = local.smallEnd(2); k <= local.bigEnd(2); ++k)
            for (int j = local.smallEnd(1); j <= local.bigEnd(1); ++j)
                for (int i = local.smallEnd(0); i <= local.bigEnd(0); ++i) chi_field(i, j, k, 0) = x[n++];
    }
    if (chi_field.nGrow() > 0) chi_field.FillBoundary(m_geom.periodicity());           // :742-744
}

void EffectiveDiffusivityHypre::gradientSums(amrex::Real sums[AMREX_SPACEDIM], long long& n_active) {
    for (int a = 0; a < AMREX_SPACEDIM; ++a) sums[a] = 0.0;
    n_active = m_num_active;
    if (!m_solver || !m_converged || m_num_active == 0) return;
    double s3[3];
    int64_t na = 0;
    oi_check(oi_cell_gradient_sums(m_solver, s3, &na), "oi_cell_gradient_sums");
    for (int a = 0; a < 3; ++a) sums[a] = s3[a];
}

}  // namespace OpenImpala
