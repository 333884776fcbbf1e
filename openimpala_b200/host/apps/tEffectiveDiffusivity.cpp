// Stand-alone driver in the mould of the reference's src/props/tEffectiveDiffusivity.cpp:176-478:
// read inputs (tifffile, phase_id, threshold_val, solver, box_size, verbose, write_plotfile,
// resultsdir; hypre.eps / hypre.maxiter are read by the class), threshold the image on a fully
// periodic geometry, solve the three corrector problems chi_x, chi_y, chi_z with
// EffectiveDiffusivityHypre, form D_eff / D by the reference's own host routine
// (calculate_Deff_tensor_homogenization: central differences of chi over the active cells), and
// apply its pass criteria: every solve converged, tensor symmetric to 1e-7, diagonals >= 0.
// The device-side gradient sums (oi_cell_gradient_sums) are checked against that host tensor.
#include <cmath>
#include <iomanip>
#include <memory>
#include <string>

#include <AMReX.H>
#include <AMReX_MultiFab.H>
#include <AMReX_ParmParse.H>
#include <AMReX_Print.H>

#include "../io/TiffReader.H"
#include "../props/EffectiveDiffusivityHypre.H"

namespace {
OpenImpala::EffectiveDiffusivityHypre::SolverType stringToSolverType(const std::string& s) {
    using ST = OpenImpala::EffectiveDiffusivityHypre::SolverType;
    if (s == "Jacobi") return ST::Jacobi;
    if (s == "GMRES") return ST::GMRES;
    if (s == "FlexGMRES") return ST::FlexGMRES;
    if (s == "PCG") return ST::PCG;
    if (s == "BiCGSTAB") return ST::BiCGSTAB;
    if (s == "SMG") return ST::SMG;
    if (s == "PFMG") return ST::PFMG;
    amrex::Abort("Invalid solver string: " + s);
    return ST::FlexGMRES;
}
}  // namespace

int main(int argc, char* argv[]) {
    amrex::Initialize(argc, argv);
    bool passed = true;
    auto fail = [&](const std::string& why) { passed = false; amrex::Print() << "TEST FAILED: " << why << "\n"; };
    {
        const amrex::Real t0 = amrex::second();
        std::string tifffile, resultsdir = "./tEffectiveDiffusivity_results", solver_str = "FlexGMRES";
        int phase_id = 1, box_size = 32, verbose = 1, write_plotfile = 0;
        amrex::Real threshold_val = 0.5;
        {
            amrex::ParmParse pp;                                          // reference :191-198
            pp.get("tifffile", tifffile);
            pp.query("resultsdir", resultsdir);
            pp.query("phase_id", phase_id);
            pp.query("solver", solver_str);
            pp.query("box_size", box_size);
            pp.query("verbose", verbose);
            pp.query("write_plotfile", write_plotfile);
            pp.query("threshold_val", threshold_val);
        }
        amrex::Geometry geom;
        amrex::BoxArray ba;
        amrex::DistributionMapping dm;
        amrex::iMultiFab mf_phase;
        amrex::Box domain;
        try {
            OpenImpala::TiffReader reader(tifffile);
            domain = reader.box();
            amrex::RealBox rb({AMREX_D_DECL(0.0, 0.0, 0.0)},
                              {AMREX_D_DECL(amrex::Real(domain.length(0)), amrex::Real(domain.length(1)),
                                            amrex::Real(domain.length(2)))});
            amrex::Array<int, AMREX_SPACEDIM> is_periodic{AMREX_D_DECL(1, 1, 1)};      // cell problem: periodic box
            geom.define(domain, &rb, 0, is_periodic.data());
            ba.define(domain);
            ba.maxSize(box_size);
            dm.define(ba);
            amrex::iMultiFab no_ghost(ba, dm, 1, 0);
            reader.threshold(threshold_val, 1, 0, no_ghost);
            if (no_ghost.min(0) == no_ghost.max(0)) amrex::Print() << "Warning: phase field uniform after thresholding.\n";
            mf_phase.define(ba, dm, 1, 1);
            amrex::Copy(mf_phase, no_ghost, 0, 0, 1, 0);
            mf_phase.FillBoundary(geom.periodicity());
        } catch (const std::exception& e) {
            fail(std::string("Error during TiffReader/grid setup: ") + e.what());
        }

        if (passed) {
            if (write_plotfile) amrex::UtilCreateDirectory(resultsdir, 0755);
            const OpenImpala::Direction dirs[3] = {OpenImpala::Direction::X, OpenImpala::Direction::Y, OpenImpala::Direction::Z};
            amrex::MultiFab chi[3] = {amrex::MultiFab(ba, dm, 1, 1), amrex::MultiFab(ba, dm, 1, 1), amrex::MultiFab(ba, dm, 1, 1)};
            amrex::Real dev[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
            const long long n_total = domain.numPts();
            bool all_converged = true;
            for (int c = 0; c < 3 && all_converged; ++c) {
                if (verbose >= 1) amrex::Print() << "\n--- Solving for chi_" << "XYZ"[c] << " ---\n";
                OpenImpala::EffectiveDiffusivityHypre solver(geom, ba, dm, mf_phase, phase_id, dirs[c],
                                                             stringToSolverType(solver_str), resultsdir, verbose,
                                                             write_plotfile != 0);
                if (!solver.solve()) { all_converged = false; break; }
                solver.getChiSolution(chi[c]);
                amrex::Real sums[3];
                long long n_active = 0;
                solver.gradientSums(sums, n_active);
                for (int r = 0; r < 3; ++r)
                    dev[r][c] = n_total > 0 ? ((r == c ? (amrex::Real)n_active : 0.0) - sums[r]) / (amrex::Real)n_total : 0.0;
            }
            if (!all_converged) {
                fail("one or more chi_k solver instances FAILED to converge");
            } else {
                // calculate_Deff_tensor_homogenization (src/props/Diffusion.cpp:60-167) on the host
                amrex::Real D[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
                long long n_active = 0;
                const amrex::Real inv2dx[3] = {1.0 / (2.0 * geom.CellSize(0)), 1.0 / (2.0 * geom.CellSize(1)),
                                               1.0 / (2.0 * geom.CellSize(2))};
                for (int k = domain.smallEnd(2); k <= domain.bigEnd(2); ++k)
                    for (int j = domain.smallEnd(1); j <= domain.bigEnd(1); ++j)
                        for (int i = domain.smallEnd(0); i <= domain.bigEnd(0); ++i) {
                            if (mf_phase(i, j, k, 0) != phase_id) continue;
                            ++n_active;
                            for (int c = 0; c < 3; ++c) {
                                const amrex::Real g[3] = {(chi[c](i + 1, j, k) - chi[c](i - 1, j, k)) * inv2dx[0],
                                                          (chi[c](i, j + 1, k) - chi[c](i, j - 1, k)) * inv2dx[1],
                                                          (chi[c](i, j, k + 1) - chi[c](i, j, k - 1)) * inv2dx[2]};
                                for (int r = 0; r < 3; ++r) D[r][c] += (r == c ? 1.0 : 0.0) - g[r];
                            }
                        }
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c) D[r][c] = n_total > 0 ? D[r][c] / (amrex::Real)n_total : 0.0;
                amrex::Print() << "Effective Diffusivity Tensor D_eff / D_material (D_material=1 assumed):\n";
                for (int r = 0; r < 3; ++r)
                    amrex::Print() << "  [" << std::scientific << std::setprecision(8) << D[r][0] << ", " << D[r][1] << ", "
                                   << D[r][2] << "]\n";
                const amrex::Real sym_tol = 1e-7;                                     // reference :424-432
                if (std::abs(D[0][1] - D[1][0]) > sym_tol || std::abs(D[0][2] - D[2][0]) > sym_tol ||
                    std::abs(D[1][2] - D[2][1]) > sym_tol)
                    fail("D_eff tensor is not symmetric within tolerance!");
                else if (verbose >= 1) amrex::Print() << "  D_eff tensor symmetry check: PASS\n";
                for (int d = 0; d < 3; ++d)
                    if (D[d][d] < 0.0) fail("D_eff diagonal component is negative");  // :437-441
                amrex::Real worst = 0.0;
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c) worst = std::max(worst, std::abs(D[r][c] - dev[r][c]));
                amrex::Print() << "  Device gradient sums vs host tensor: max abs difference " << std::scientific << worst
                               << std::defaultfloat << "\n";
                if (!(worst <= 1e-9)) fail("device-side tensor differs from the host tensor");
            }
        }
        amrex::Print() << "\n--- Effective Diffusivity Test Summary ---\n  Total Run Time: " << (amrex::second() - t0) << " sec\n";
        amrex::Print() << (passed ? "  TEST RESULT: PASS\nTEST PASSED\n" : "  TEST RESULT: FAIL\nTEST FAILED\n");
    }
    amrex::Finalize();
    return passed ? 0 : 1;
}
