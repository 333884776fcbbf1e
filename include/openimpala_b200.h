/*
 * openimpala_b200.h -- C-ABI of the B200-native TortuosityHypre path.
 *
 * The reference (kramergroup/openImpala) has no FFI table for this path: the
 * seam is the C++ class surface OpenImpala::TortuosityHypre /
 * OpenImpala::VolumeFraction plus the two Fortran bind(c) kernels.  Every entry
 * point below names the reference interface it replaces (file:line relative to
 * the reference tree).  Plain pointers and sizes only; no C++/torch types.
 *
 * Conventions
 *   - fields are dense boxes, x fastest, z slowest (AMReX/Fortran order,
 *     src/io/TiffReader.cpp:130-134, src/io/RawReader.cpp:310-313);
 *   - one handle = one (image, phase, direction) solve, like one
 *     TortuosityHypre object (src/props/TortuosityHypre.H:68-80);
 *   - a handle owns one z-slab [z_begin, z_begin+nz_local) of the global box;
 *     with n_ranks == 1 the slab is the whole box;
 *   - every function returns OI_OK (0) or a negative oi_status; text of the
 *     last failure on the calling thread: oi_last_error();
 *   - not thread-safe per handle; no exceptions cross the boundary;
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     point fails with OI_ERR_CUDA.
 */
#ifndef OPENIMPALA_B200_H
#define OPENIMPALA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OI_B200_VERSION 100

typedef enum oi_status {
    OI_OK = 0,
    OI_ERR_INVALID = -1,   /* bad argument / call order          */
    OI_ERR_CUDA = -2,      /* CUDA runtime failure or no device  */
    OI_ERR_NCCL = -3,      /* NCCL failure / library not found   */
    OI_ERR_NOMEM = -4
} oi_status;

/* OpenImpala::Direction, src/props/Tortuosity.H:9-13 */
typedef enum oi_direction { OI_DIR_X = 0, OI_DIR_Y = 1, OI_DIR_Z = 2 } oi_direction;

/* Preconditioner selection (the reference hard-wires FlexGMRES + SMG,
 * src/props/TortuosityHypre.cpp:664-678; every SolverType maps to PCG here). */
typedef enum oi_precond { OI_PRECOND_MG = 0, OI_PRECOND_JACOBI = 1 } oi_precond;

/* Ghost-plane exchange between z-slabs (AMReX FillBoundary in the reference,
 * TortuosityHypre.cpp:339, 584-585, 1033).  AUTO = peer-memory stores over
 * NVLink (CUDA IPC) when every rank can map its neighbours, else NCCL
 * send/recv.  The environment variable OI_HALO_MODE=nccl|p2p overrides. */
typedef enum oi_halo_mode { OI_HALO_AUTO = 0, OI_HALO_NCCL = 1, OI_HALO_PEER = 2 } oi_halo_mode;

/* Which linear problem the handle solves.
 * TORTUOSITY: flow-through Laplace solve of TortuosityHypre (Dirichlet inlet /
 *   outlet, percolation mask), src/props/TortuosityHypre.cpp.
 * CELL: the periodic cell problem for the corrector chi_k of the homogenisation
 *   path, EffectiveDiffusivityHypre (src/props/EffectiveDiffusivityHypre.cpp:104-745,
 *   src/props/EffDiffFillMtx.F90:109-258); `direction` is k; vlo/vhi are unused. */
typedef enum oi_problem { OI_PROBLEM_TORTUOSITY = 0, OI_PROBLEM_CELL = 1 } oi_problem;

struct oi_comm;

typedef struct oi_params {
    int32_t nx, ny, nz;        /* global box, cells (Geometry::Domain)            */
    int32_t z_begin, nz_local; /* this rank's z-slab                              */
    int32_t direction;         /* oi_direction                                    */
    int32_t phase_id;          /* m_phase, TortuosityHypre.cpp:115                */
    double  vlo, vhi;          /* Dirichlet values, TortuosityHypre.cpp:118       */
    double  dx[3];             /* Geometry::CellSize (1,1,1 in both drivers)      */
    double  eps;               /* hypre.eps,  default 1e-9  (TortuosityHypre.cpp:142) */
    int32_t maxiter;           /* hypre.maxiter, default 200 (TortuosityHypre.cpp:143) */
    int32_t verbose;
    int32_t device;            /* CUDA device ordinal, -1 = current               */
    int32_t precond;           /* oi_precond                                      */
    int32_t mg_degree;         /* smoother polynomial degree per leg, 0 = default */
    int32_t stencil_variant;   /* 0 = z-plane ring in shared memory (cp.async, default),
                                  2 = register z-march, 1 = simple gather            */
    int32_t flux_polish;       /* 0 (default, the reference's behaviour): stop on the
                                  residual rule alone; the 1e-6 flux gate of value()
                                  (TortuosityHypre.cpp:794-823) then decides NaN.
                                  1: keep iterating (<= maxiter) until the flux
                                  imbalance is 2x inside that gate                  */
    int32_t halo_mode;         /* oi_halo_mode: how ghost planes travel between z-slabs */
    int32_t problem;           /* oi_problem                                      */
    struct oi_comm* comm;      /* z-slab communicator (oi_comm_create) or NULL for
                                  a single slab; must outlive the handle            */
} oi_params;

typedef struct oi_solve_info {
    int32_t iterations;        /* getSolverIterations(), TortuosityHypre.H:116     */
    int32_t converged;         /* getSolverConverged(),  TortuosityHypre.H:114     */
    double  rel_residual;      /* getFinalRelativeResidualNorm(), :115             */
    double  b_norm;            /* ||b||_2 of the un-eliminated rhs (stop rule)     */
    double  solve_ms;          /* device time of the Krylov loop                   */
    double  setup_ms;          /* coarse-operator build                            */
} oi_solve_info;

typedef struct oi_solver oi_solver; /* opaque */
typedef struct oi_comm oi_comm;     /* opaque */

int         oi_version(void);
const char* oi_last_error(void);
int         oi_device_count(int* count);
void        oi_default_params(oi_params* p);

/* ---- z-slab communicator --------------------------------------------------
 * Stands in for MPI_COMM_WORLD of the reference (TortuosityHypre.cpp:211,
 * 463-484; AMReX FillBoundary / ParallelDescriptor::Reduce*): one process per
 * GPU, rank r owns slab r, halo planes and scalar reductions go over NCCL.
 * id128 is a 128-byte ncclUniqueId made by oi_comm_unique_id on one rank and
 * distributed by the host program (torch.distributed, MPI_Bcast, a file ...). */
int oi_comm_unique_id(void* id128_out);
int oi_comm_create(oi_comm** out, int32_t rank, int32_t n_ranks, const void* id128, int32_t device);
int oi_comm_destroy(oi_comm* c);
/* Sum of n int64 values over the ranks of the communicator, in place (host pointer).  Stands in for
 * amrex::ParallelDescriptor::ReduceLongSum of the reference's callers (VolumeFraction.cpp:58-60).
 * A NULL communicator is one rank: the values stay as they are. */
int oi_comm_allreduce_sum_i64(oi_comm* c, int64_t* values, int32_t n);

/* ---- VolumeFraction::value, src/props/VolumeFraction.cpp:22-66 ---------- */
/* Count cells == phase in a host field (copied to the device, counted there).
 * total_count is n (Sum of tile numPts, VolumeFraction.cpp:48). */
int oi_count_phase_i32(const int32_t* host_field, int64_t n, int32_t phase,
                       int64_t* phase_count, int64_t* total_count);
int oi_count_phase_u8(const uint8_t* host_field, int64_t n, int32_t phase,
                      int64_t* phase_count, int64_t* total_count);

/* ---- TortuosityHypre ctor, src/props/TortuosityHypre.cpp:100-191 -------- */
int oi_create(oi_solver** out, const oi_params* p);
int oi_destroy(oi_solver* h);

/* Phase field of the local slab (m_mf_phase deep copy, TortuosityHypre.cpp:132).
 * Host buffers hold nx*ny*nz_local values; the *_device variant takes a device
 * pointer to uint8 values already resident in HBM. */
int oi_set_phase_i32(oi_solver* h, const int32_t* host_phase);
int oi_set_phase_u8(oi_solver* h, const uint8_t* host_phase);
int oi_set_phase_device_u8(oi_solver* h, const void* device_phase);

/* Streamed alternative to oi_set_phase_u8 for images that should never sit whole in host
 * memory (the reference re-opens the TIFF per tile and z-slice, src/io/TiffReader.cpp:320-338):
 * the slab arrives in z-chunks of at most max_planes_per_chunk planes through two
 * library-owned PINNED staging buffers.  The caller decodes a chunk straight into
 * oi_phase_stream_buffer(h, which), submits it (asynchronous H2D copy + count + conversion
 * to the device mask) and decodes the next chunk into the other buffer meanwhile;
 * oi_phase_stream_buffer waits until the upload that last used that buffer has left it.
 * Every plane of the slab must be submitted exactly once before oi_phase_stream_end. */
int oi_phase_stream_begin(oi_solver* h, int32_t max_planes_per_chunk);
int oi_phase_stream_buffer(oi_solver* h, int32_t which, uint8_t** host_buffer);
int oi_phase_stream_submit(oi_solver* h, int32_t which, int32_t z_local_begin, int32_t nz_chunk);
int oi_phase_stream_end(oi_solver* h);

/* Global counts over the resident phase field (VolumeFraction.cpp:22-66). */
int oi_volume_fraction(oi_solver* h, int64_t* phase_count, int64_t* total_count);

/* tortuosity_remspot, src/props/Tortuosity_filcc.F90:88-177 (optional filter,
 * tortuosity.remspot_passes, TortuosityHypre.cpp:248-292). */
int oi_remspot(oi_solver* h, int32_t passes);

/* generateActivityMask + parallelFloodFill, TortuosityHypre.cpp:394-558 /
 * :297-389, then setupMatrixEquation / tortuosity_fillmtx
 * (TortuosityHypre.cpp:562-649, TortuosityHypreFill.F90:44-314) in matrix-free
 * form: per-cell connectivity byte, rhs norm and initial guess.
 * n_active = global sum(mask) (TortuosityHypre.cpp:549). */
int oi_build_mask(oi_solver* h, int64_t* n_active);

/* solve(), TortuosityHypre.cpp:654-756 */
int oi_solve(oi_solver* h, oi_solve_info* info);

/* global_fluxes(), TortuosityHypre.cpp:1000-1134 (already x face area). */
int oi_fluxes(oi_solver* h, double* flux_in, double* flux_out,
              int64_t* n_active_in, int64_t* n_active_out);

/* Cell problem only.  sums3[a] = sum over active cells of the central difference
 * d(chi_k)/dx_a of the solved corrector (periodic box, chi = 0 in the solid):
 * the per-direction ingredient of calculate_Deff_tensor_homogenization
 * (src/props/Diffusion.cpp:60-167): D_eff[a][k] = (delta_ak * n_active - sums3[a]) / N. */
int oi_cell_gradient_sums(oi_solver* h, double* sums3, int64_t* n_active);

/* checkMatrixProperties(), TortuosityHypre.cpp:896-982: device-side check of the
 * same invariants on the matrix-free rows.  ok = 1 when all pass (global). */
int oi_check_matrix_properties(oi_solver* h, int32_t* ok);

/* ---- read-backs / test hooks (local slab, nx*ny*nz_local values) -------- */
int oi_get_mask_u8(oi_solver* h, uint8_t* host_mask);      /* m_mf_active_mask  */
int oi_get_solution(oi_solver* h, double* host_x);          /* HYPRE m_x         */
int oi_set_solution(oi_solver* h, const double* host_x);
int oi_get_initial_guess(oi_solver* h, double* host_x0);    /* xinit, F90:233-262 */
int oi_get_rhs(oi_solver* h, double* host_rhs);             /* rhs,   F90:115-225 */
/* a[7*m+s], slot order C,-x,+x,-y,+y,-z,+z (TortuosityHypreFill.F90:20-26):
 * the rows tortuosity_fillmtx would have produced, rebuilt from the
 * connectivity bytes. */
int oi_get_matrix_rows(oi_solver* h, double* host_a7);
/* y = A_elim * x on the active interior unknowns (0 elsewhere). */
int oi_apply_operator(oi_solver* h, const double* host_x, double* host_y);
/* z = M^-1 r : one application of the preconditioner (test hook). */
int oi_apply_precond(oi_solver* h, const double* host_r, double* host_z);
/* Average device time (ms) of one launch of a named kernel over `reps`
 * launches on the resident problem, CUDA events on the solver stream.
 * name: "apply" | "smooth" | "residual_restrict" | "axpy2_dot" | "xpby" | "dot"
 *       | "count_phase" */
int oi_time_kernel(oi_solver* h, const char* name, int32_t reps, double* avg_ms,
                   int64_t* cells);
/* CUDA-event stopwatch on the solver's own stream (the stream every kernel of
 * this handle is launched on): record slot 0..7, then elapsed ms between two
 * recorded slots (synchronises on the later one). */
int oi_timer_record(oi_solver* h, int32_t slot);
int oi_timer_elapsed_ms(oi_solver* h, int32_t slot_begin, int32_t slot_end, double* ms);
/* Device blocks of destroyed handles are kept in a process-wide cache so that the
 * next handle of the same shape allocates nothing (OI_NO_MEM_CACHE=1 disables it).
 * This returns the idle blocks to the driver; call it only while no multi-slab
 * handle is alive on any rank (neighbours may have the blocks mapped). */
int oi_release_cached_memory(int64_t* bytes_released);
/* Local slab: stats3[0] = unknowns, [1] = aligned 2-cell groups and [2] = aligned 4-cell
 * groups that hold at least one unknown.  The fp64 / fp32 kernels skip a 16-byte group
 * without unknowns, so these are the granules that are actually read and written. */
int oi_sparsity(oi_solver* h, int64_t* stats3);
/* Which halo path the handle uses (oi_halo_mode; AUTO for a single slab) and how
 * many ghost-plane exchanges went through peer memory so far. */
int oi_halo_info(oi_solver* h, int32_t* mode, int64_t* peer_exchanges);
/* Iterations replayed as a CUDA graph.  On a single slab of at most 2^25 cells (the
 * launch-bound regime; OI_GRAPH=1 lifts the limit, OI_GRAPH=0 disables) every PCG
 * iteration after the first is one graph launch: replays = how many so far on this
 * handle, kernels_per_iteration = kernel nodes in one captured iteration (0 before the
 * first capture).  Replaces nothing in the reference: HYPRE's Krylov loop
 * (src/props/TortuosityHypre.cpp:683) is host driven. */
int oi_graph_info(oi_solver* h, int64_t* replays, int64_t* kernels_per_iteration);
/* Number of kernels this handle has launched since creation. */
int oi_launch_count(oi_solver* h, int64_t* launches);

/* ---- the reference's two existing C-ABI kernels, under their own names -----------------
 * Fortran bind(c) convention: every scalar by reference, arrays with their own lo/hi
 * bounds (x fastest), HOST pointers.  Drop-in replacements for the Fortran objects:
 *   tortuosity_fillmtx  src/props/TortuosityHypreFill_F.H:47-68 (TortuosityHypreFill.F90:44-314)
 *   tortuosity_remspot  src/props/Tortuosity_filcc_F.H:65-67    (Tortuosity_filcc.F90:88-177)
 * The box is copied to the device, one kernel fills it, the result is copied back; rows,
 * rhs and xinit are bit-identical to the Fortran's (openimpala_b200/csrc/oi_refabi.cu).
 * They abort (like `error stop`) when no CUDA device is present. */
void tortuosity_fillmtx(double* a, double* rhs, double* xinit, const int* nval, const int* p,
                        const int* p_lo, const int* p_hi, const int* active_mask, const int* mask_lo,
                        const int* mask_hi, const int* bxlo, const int* bxhi, const int* domlo,
                        const int* domhi, const double* dxinv, const double* vlo, const double* vhi,
                        const int* phase, const int* dir, const int* debug_print_level);
void tortuosity_remspot(int* q, const int* q_lo, const int* q_hi, const int* ncomp, const int* bxlo,
                        const int* bxhi, const int* domlo, const int* domhi);

#ifdef __cplusplus
}
#endif
#endif /* OPENIMPALA_B200_H */
