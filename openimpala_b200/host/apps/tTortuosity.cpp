// Stand-alone driver in the mould of the reference's src/props/tTortuosity.cpp:
// read inputs (tifffile, phase_id, direction, solver, box_size, v_lo, v_hi,
// expected_vf, expected_tau ...), build the fields, VolumeFraction,
// TortuosityHypre, checkMatrixProperties(), value(); exit code = pass/fail.
#include <cmath>
#include <iomanip>
#include <memory>
#include <string>

#include <AMReX.H>
#include <AMReX_ParmParse.H>
#include <AMReX_Print.H>

#include "../io/TiffReader.H"
#include "../props/TortuosityHypre.H"
#include "../props/VolumeFraction.H"

namespace {
OpenImpala::Direction stringToDirection(const std::string& s) {
    if (s == "X" || s == "x") return OpenImpala::Direction::X;
    if (s == "Y" || s == "y") return OpenImpala::Direction::Y;
    if (s == "Z" || s == "z") return OpenImpala::Direction::Z;
    amrex::Abort("Invalid direction string: " + s);
    return OpenImpala::Direction::X;
}
OpenImpala::TortuosityHypre::SolverType stringToSolverType(const std::string& s) {
    using ST = OpenImpala::TortuosityHypre::SolverType;
    if (s == "Jacobi") return ST::Jacobi;
    if (s == "GMRES") return ST::GMRES;
    if (s == "FlexGMRES") return ST::FlexGMRES;
    if (s == "PCG") return ST::PCG;
    if (s == "BiCGSTAB") return ST::BiCGSTAB;
    if (s == "SMG") return ST::SMG;
    if (s == "PFMG") return ST::PFMG;
    amrex::Abort("Invalid solver string: " + s);
    return ST::GMRES;
}
}  // namespace

int main(int argc, char* argv[]) {
    amrex::Initialize(argc, argv);
    bool passed = true;
    auto fail = [&](const std::string& why) { passed = false; amrex::Print() << "TEST FAILED: " << why << "\n"; };
    {
        const amrex::Real t0 = amrex::second();
        std::string tifffile, resultsdir = "./tortuosity_results", direction_str = "X", solver_str = "GMRES";
        int phase_id = 1, box_size = 32, verbose = 1, write_plotfile = 0;
        amrex::Real expected_vf = -1.0, expected_tau = -1.0, vf_tolerance = 1e-9, tau_tolerance = 1e-5;
        amrex::Real threshold_val = 0.5, v_lo = 0.0, v_hi = 1.0;
        {
            amrex::ParmParse pp;
            pp.get("tifffile", tifffile);
            pp.query("resultsdir", resultsdir);
            pp.query("phase_id", phase_id);
            pp.query("direction", direction_str);
            pp.query("solver", solver_str);
            pp.query("box_size", box_size);
            pp.query("verbose", verbose);
            pp.query("write_plotfile", write_plotfile);
            pp.query("expected_vf", expected_vf);
            pp.query("expected_tau", expected_tau);
            pp.query("vf_tolerance", vf_tolerance);
            pp.query("tau_tolerance", tau_tolerance);
            pp.query("threshold_val", threshold_val);
            pp.query("v_lo", v_lo);
            pp.query("v_hi", v_hi);
        }
        const OpenImpala::Direction direction = stringToDirection(direction_str);
        const auto solver_type = stringToSolverType(solver_str);

        amrex::Geometry geom;
        amrex::BoxArray ba;
        amrex::DistributionMapping dm;
        amrex::iMultiFab mf_phase;
        try {
            OpenImpala::TiffReader reader(tifffile);
            const amrex::Box domain = reader.box();
            amrex::RealBox rb({AMREX_D_DECL(0.0, 0.0, 0.0)},
                              {AMREX_D_DECL(amrex::Real(domain.length(0)), amrex::Real(domain.length(1)),
                                            amrex::Real(domain.length(2)))});
            amrex::Array<int, AMREX_SPACEDIM> is_periodic{AMREX_D_DECL(0, 0, 0)};
            geom.define(domain, &rb, 0, is_periodic.data());
            ba.define(domain);
            ba.maxSize(box_size);
            dm.define(ba);
            amrex::iMultiFab no_ghost(ba, dm, 1, 0);
            reader.threshold(threshold_val, 1, 0, no_ghost);
            if (no_ghost.min(0) == no_ghost.max(0)) fail("Phase field uniform after thresholding.");
            mf_phase.define(ba, dm, 1, 1);
            amrex::Copy(mf_phase, no_ghost, 0, 0, 1, 0);
            mf_phase.FillBoundary(geom.periodicity());
        } catch (const std::exception& e) {
            fail(std::string("Error during TiffReader/grid setup: ") + e.what());
        }

        amrex::Real actual_vf = 0.0;
        if (passed) {
            OpenImpala::VolumeFraction vf_calc(mf_phase, phase_id);
            actual_vf = vf_calc.value_vf(false);
            amrex::Print() << " Calculated Volume Fraction (Phase " << phase_id << "): " << std::setprecision(9) << actual_vf << "\n";
            if (expected_vf >= 0.0 && std::abs(actual_vf - expected_vf) > vf_tolerance) fail("Volume fraction mismatch.");
        }
        amrex::Real actual_tau = std::numeric_limits<amrex::Real>::quiet_NaN();
        if (passed && actual_vf > std::numeric_limits<amrex::Real>::epsilon()) {
            amrex::UtilCreateDirectory(resultsdir, 0755);
            auto tort = std::make_unique<OpenImpala::TortuosityHypre>(geom, ba, dm, mf_phase, actual_vf, phase_id, direction,
                                                                      solver_type, resultsdir, v_lo, v_hi, verbose,
                                                                      write_plotfile != 0);
            if (!tort->checkMatrixProperties()) fail("Assembled matrix/vector failed property checks (check log).");
            if (passed) {
                actual_tau = tort->value();
                if (std::isnan(actual_tau) || std::isinf(actual_tau)) fail("Calculated tortuosity is NaN or Inf!");
                amrex::Print() << " Solver iterations: " << tort->getSolverIterations() << "  relres: " << std::scientific
                               << tort->getFinalRelativeResidualNorm() << std::defaultfloat << "  converged: "
                               << tort->getSolverConverged() << "\n";
                amrex::Print() << " Active VF: " << std::setprecision(9) << tort->getActiveVolumeFraction() << "  FluxIn: "
                               << tort->getFluxIn() << "  FluxOut: " << tort->getFluxOut() << "\n";
            }
            if (passed) {
                amrex::Print() << " Final Calculated Tortuosity: " << std::fixed << std::setprecision(9) << actual_tau << "\n";
                if (expected_tau >= 0.0 && std::abs(actual_tau - expected_tau) > tau_tolerance) fail("Tortuosity mismatch.");
            }
        }
        amrex::Print() << " Run time (seconds) = " << (amrex::second() - t0) << "\n";
        amrex::Print() << (passed ? "TEST PASSED\n" : "TEST FAILED\n");
    }
    amrex::Finalize();
    return passed ? 0 : 1;
}
