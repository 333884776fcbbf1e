// Internal launcher interface between the solver driver and the .cu kernels.
#pragma once
#include "oi_common.cuh"

namespace oi {

// ---------------------------------------------------------------- level 0
// APPLY works on fp64 fields (Krylov vectors); SMOOTH / RESTRICT / first sweep /
// prolongation work on mg_t fields (multigrid preconditioner), hence the void*.
struct L0Args {
    Grid g;
    const uint8_t* flags;      // connectivity bytes, ghost planes at k=-1, k=nz
    const void* u;             // input field (ghost planes)
    const void* b;             // rhs (SMOOTH / RESTRICT)
    void* out;                 // output field, or coarse rhs for RESTRICT
    double w;                  // Jacobi weight (SMOOTH) or scale (APPLY)
    const void* ec;            // coarse correction (ADDC / prolongation) or nullptr
    int cnx, cny;              // coarse dims (ADDC / RESTRICT)
    int fx, fy, fz;            // coarsening factors to level 1
    double* red_partials;      // reduction scratch (DOT)
    unsigned int* red_counter;
    double* red_out;
    int n_sm;
    // z-slabs: ghost planes of `u` arrive through the flag words of hin (the boundary CTAs wait inside
    // the kernel), boundary planes of `out` are stored into the neighbours' ghost planes (hout).
    // All-null (the default) = no halo work in the kernel.  Ring kernels, prolongation.
    HaloIn hin;
    HaloOut hout;
    // z-slabs, two sweeps per pass: the ring SMOOTH kernel run on the two boundary planes only (bnd_only: nothing
    // is stored locally, the planes go to the neighbours through hout), and the planes received from them, which
    // the pair kernel takes as the intermediate iterate at k = -1 / k = nz (nullptr: outside the box, zero)
    int bnd_only;
    const void* vb_lo;
    const void* vb_hi;
};

long long l0_max_blocks(const Grid& g, int n_sm);
int pick_zchunk(const Grid& g, int n_sm);
// shared-memory ring kernels (oi_level0_ring.cu): mode 0 APPLY, 1 SMOOTH, 2 RESTRICT
bool ring_supported(const L0Args& a, int mode);
void ring_launch(const L0Args& a, int mode, bool dot, cudaStream_t st);
// TMA variant of the ring (oi_level0_tma.cu): APPLY and SMOOTH through cp.async.bulk.tensor + mbarriers.
// tma_launch returns false (nothing launched) when the tensor maps cannot be encoded.
bool tma_supported(const L0Args& a, int mode);
bool tma_launch(const L0Args& a, int mode, bool dot, cudaStream_t st);
// z += P*ec on unknown cells (prolongation + correction)
void l0_prolong_add(const L0Args& a, cudaStream_t st);
bool prolong_halo_supported(const L0Args& a);
void l0_apply(const L0Args& a, bool dot, int variant, cudaStream_t st);
void l0_smooth(const L0Args& a, bool addc, bool dot, int variant, cudaStream_t st);
void l0_residual_restrict(const L0Args& a, int variant, cudaStream_t st);
void l0_residual(const L0Args& a, cudaStream_t st);
void l0_jacobi_first(const L0Args& a, cudaStream_t st);
// two smoothing sweeps in one pass (oi_level0_pair.cu): out = S_w2(S_w1(u)); needs one z-slab,
// a non-periodic box, nx % 4 == 0 and fp32 multigrid vectors
bool pair_supported(const L0Args& a);
// the same kernel on one z-slab of several (needs the neighbours' intermediate boundary planes, vb_lo / vb_hi)
bool pair_supported_slab(const L0Args& a);
void l0_smooth_pair(const L0Args& a, double w1, double w2, bool dot, int variant, cudaStream_t st);

// ---------------------------------------------------------------- coarse levels
// 7-point operator with stored face couplings (Galerkin sums of fine faces):
//   (A x)_I = dg_I x_I - cxp_I x_{I+ex} - cxp_{I-ex} x_{I-ex} - ... (y, z)
struct CoarseLevel {
    int nx, ny, nz;            // local dims (nz local planes)
    int z0;                    // global index of local plane 0 at this level
    int nzg;                   // global planes at this level
    long long plane;
    int fx, fy, fz;            // coarsening factors from this level to the next
    int periodic;              // PER_X | PER_Y | PER_Z of the box (cell problem), else 0
    int replicated;            // multi-rank: this level holds the whole box on every rank (no halo)
    // Half-precision copies of cxp, cyp, czp, dg for the bandwidth-bound smoothing / residual kernel (20 instead
    // of 28 bytes per cell and sweep).  Only set on levels where every coefficient survives the conversion
    // exactly (small-integer multiples of 2^-level: levels 1-3 for unit cells), so both sets describe the
    // SAME operator; nullptr otherwise.  Same layout and ghost planes as the float arrays.
    const unsigned short *hx, *hy, *hz, *hd;
    float *cxp, *cyp, *czp;    // coupling to +x,+y,+z neighbour (>= 0), ghost planes
    float* dg;                 // diagonal (0 = empty aggregate)
    float *dgx, *dgy, *dgz;    // its shares by axis (build time only: each is scaled by 1/f_axis)
    mg_t *x, *b, *t;           // solution, rhs, scratch (ghost planes)
};

void coarse_build_from_flags(const Grid& g, const uint8_t* flags, int dir_axis, int n_dir_global,
                             const CoarseLevel& c, int fx, int fy, int fz, cudaStream_t st);
void coarse_build_from_coarse(const CoarseLevel& f, const CoarseLevel& c, cudaStream_t st);
// dst[i] = half(src[i]) for i in [0, n); *mismatch += number of values the conversion changed
bool coarse_half_applicable(const CoarseLevel& L);       // this level runs the kernel that can read them
void coarse_to_half(const float* src, unsigned short* dst, long long n, unsigned long long* mismatch, cudaStream_t st);
void coarse_jacobi_first(const CoarseLevel& L, const mg_t* b, mg_t* out, double w, cudaStream_t st);
// out = x + w (b - A x) / dg
// hin / hout (z-slabs, optional): wait for the ghost planes of x inside the kernel / store the boundary planes of
// out into the neighbours' ghost planes (only where coarse_halo_supported(L))
void coarse_smooth(const CoarseLevel& L, const mg_t* x, const mg_t* b, mg_t* out, double w,
                   cudaStream_t st, const HaloIn* hin = nullptr, const HaloOut* hout = nullptr);
void coarse_residual(const CoarseLevel& L, const mg_t* x, const mg_t* b, mg_t* out, cudaStream_t st,
                     const HaloIn* hin = nullptr);
bool coarse_halo_supported(const CoarseLevel& L);
// out = S_w2(S_w1(x)) in one pass (single z-slab, non-periodic, half coefficient copies, >= 2^21 cells)
bool coarse_pair_supported(const CoarseLevel& L);
void coarse_smooth_pair(const CoarseLevel& L, const mg_t* x, const mg_t* b, mg_t* out, double w1, double w2,
                        cudaStream_t st);
// x += P * ec  (ec lives on level `next`)
void coarse_prolong_add(const CoarseLevel& L, mg_t* x, const CoarseLevel& next, const mg_t* ec,
                        cudaStream_t st);
void coarse_restrict(const CoarseLevel& f, const mg_t* res, const CoarseLevel& c, mg_t* bc,
                     cudaStream_t st);

// One-CTA V-cycle over the last levels of the hierarchy (each holding the whole box in z):
// L[0] is the first tail level -- rhs in L[0].b, correction left in L[0].x -- L[n_levels-1]
// the coarsest level of the hierarchy.  w/deg: smoothing weights of a level with a coarser
// one below; wc/deg_c: the coarsest level's sweeps.
constexpr int TAIL_MAX_LEVELS = 6;
struct TailArgs {
    int n_levels, deg, deg_c;
    CoarseLevel L[TAIL_MAX_LEVELS];
    mg_t w[16], wc[16];
};
// staged: run on shared-memory copies of the fields when they fit (else, and for fp64
// multigrid vectors, in global memory)
void coarse_tail_cycle(const TailArgs& a, bool staged, cudaStream_t st);

// ---------------------------------------------------------------- vector ops (K4)
// x += a p ; r -= a q ; out[0] = r.r      with a = num[0]/den[0] read on device
void vec_axpy2_dot(long long n, double* x, double* r, const double* p, const double* q,
                   const double* num, const double* den, double* partials, unsigned int* counter,
                   double* out, int n_sm, cudaStream_t st);
// same, plus what the next preconditioner application needs: r32 <- mg_t(r_new) and
// z1 <- w0 * r_new / diag on unknowns (its first smoothing sweep from a zero guess).
// x == nullptr: the solution update is deferred to vec_xpby / vec_axpy (x and p untouched)
void vec_axpy2_dot_first(const Grid& g, const uint8_t* flags, long long n, double* x, double* r,
                         const double* p, const double* q, mg_t* r32, mg_t* z1, const double* num,
                         const double* den, double w0, double* partials, unsigned int* counter,
                         double* out, int n_sm, cudaStream_t st, const HaloOut* ho = nullptr);
// p = z + (num/den) p      (z is a multigrid-precision vector; both zero off the unknowns)
// x != nullptr: first x += (anum/aden) p with the old p (deferred solution update)
// ho (z-slabs): also store the new p's boundary planes into the neighbours' ghost planes
void vec_xpby(long long n, const uint8_t* flags, double* p, const mg_t* z, const double* num,
              const double* den, double* x, const double* anum, const double* aden, int n_sm, cudaStream_t st,
              long long plane = 0, const HaloOut* ho = nullptr);
bool vec_halo_supported(long long plane, long long n);
// x += (num/den) p on the unknowns
void vec_axpy(long long n, const uint8_t* flags, double* x, const double* p, const double* num, const double* den,
              int n_sm, cudaStream_t st);
// conversions between Krylov (fp64) and multigrid precision
void vec_to_mg(long long n, mg_t* dst, const double* src, int n_sm, cudaStream_t st);
void vec_from_mg(long long n, double* dst, const mg_t* src, int n_sm, cudaStream_t st);
// out[0] = a.b
void vec_dot(long long n, const double* a, const double* b, double* partials,
             unsigned int* counter, double* out, int n_sm, cudaStream_t st);
void vec_copy(long long n, double* dst, const double* src, int n_sm, cudaStream_t st);
// z = r / diag on unknowns (Jacobi preconditioner), out[0] = r.z
void l0_jacobi_precond_dot(const Grid& g, const uint8_t* flags, const double* r, mg_t* z,
                           double* partials, unsigned int* counter, double* out, int n_sm,
                           cudaStream_t st);
int vec_max_blocks(int n_sm);

// ---------------------------------------------------------------- mask / setup (K1, K2, K7, K8)
void count_phase_u8(const uint8_t* f, long long n, int phase, unsigned long long* out, int n_sm,
                    cudaStream_t st);
void count_phase_i32(const int32_t* f, long long n, int phase, unsigned long long* out, int n_sm,
                     cudaStream_t st);
void phase_i32_to_u8(const int32_t* in, uint8_t* is_phase, long long n, int phase, int n_sm,
                     cudaStream_t st);
void phase_u8_to_isphase(const uint8_t* in, uint8_t* is_phase, long long n, int phase, int n_sm,
                         cudaStream_t st);

// connected components of {is_phase}: labels = min linear index of the component
// (returns the number of kernels it launched)
int ccl_label(const uint8_t* is_phase, int* labels, int nx, int ny, int nz, int n_sm, cudaStream_t st);
// reach[root] |= 1 (touches inlet plane) | 2 (touches outlet plane); planes given
// in local coordinates, -1 = not on this slab
void ccl_mark_planes(const uint8_t* is_phase, const int* labels, unsigned int* reach, int nx,
                     int ny, int nz, int dir, int lo_local, int hi_local, int n_sm,
                     cudaStream_t st);
// cross-slab propagation: for cells of local plane k (0 or nz-1) whose neighbour
// across the slab boundary is phase, OR the neighbour's reach bits into the
// local root; sets *changed when any bit was added.
void ccl_export_plane(const uint8_t* is_phase, const int* labels, const unsigned int* reach,
                      uint8_t* plane_bits, int nx, int ny, int k, int n_sm, cudaStream_t st);
void ccl_import_plane(const uint8_t* is_phase, const int* labels, unsigned int* reach,
                      const uint8_t* nbr_bits, int nx, int ny, int k, int* changed, int n_sm,
                      cudaStream_t st);
// mask = phase && reach[label]==3 ; writes active u8 (with ghost planes untouched)
void build_active(const uint8_t* is_phase, const int* labels, const unsigned int* reach,
                  uint8_t* active, int nx, int ny, int nz, unsigned long long* n_active, int n_sm,
                  cudaStream_t st);
// connectivity bytes from the active mask (ghost planes of `active` must be valid)
// counts[0] = active cells on the inlet plane, counts[1] = on the outlet plane
void build_flags(const Grid& g, const uint8_t* active, uint8_t* flags, int dir,
                 unsigned long long* counts, cudaStream_t st);
// x0 (F90:233-262).  mirror_quirk = 1 reproduces the reference's xinit exactly
// (cells with diagonal == 1 keep 0, F90:233); 0 = plain ramp with exact
// Dirichlet values, which is what the solver starts from.
void fill_initial_guess(const Grid& g, const uint8_t* flags, double* x, int dir, int n_dir_global,
                        double vlo, double vhi, int mirror_quirk, cudaStream_t st);
// isolated-voxel filter with the reference's in-place (sequential) semantics, as a
// fixed-point iteration on flip flags: one round, then the final xor
void remspot_round(const uint8_t* v, const uint8_t* fcur, uint8_t* fnext, int nx, int ny, int nz,
                   int* changed, int n_sm, cudaStream_t st);
void remspot_apply(uint8_t* v, const uint8_t* f, long long n, unsigned long long* flips, int n_sm,
                   cudaStream_t st);
void count_nonbinary_u8(const uint8_t* f, long long n, unsigned long long* out, int n_sm, cudaStream_t st);
void count_nonbinary_i32(const int32_t* f, long long n, unsigned long long* out, int n_sm, cudaStream_t st);
// boundary fluxes (TortuosityHypre.cpp:1052-1105): out[0]=sum_in, out[1]=sum_out
void flux_planes(const Grid& g, const uint8_t* flags, const double* x, int dir, int n_dir_global,
                 double* partials, unsigned int* counter, double* out, cudaStream_t st);
// matrix rows / rhs as tortuosity_fillmtx would have produced them
void export_rows(const Grid& g, const uint8_t* flags, const uint8_t* active, int dir,
                 int n_dir_global, double vlo, double vhi, double* a7, double* rhs,
                 cudaStream_t st);
// cell problem of the homogenisation path (EffectiveDiffusivityHypre):
// r += sign * b (r may be null) and out[0] = ||b||^2 ; out[0..2] = sum over active
// cells of d(chi)/dx, d(chi)/dy, d(chi)/dz by central differences
void cellp_rhs(const Grid& g, const uint8_t* flags, double* r, int dir, double sign, double* partials,
               unsigned int* counter, double* out, cudaStream_t st);
void cellp_grad_sums(const Grid& g, const uint8_t* flags, const double* x, double* partials,
                     unsigned int* counter, double* out, cudaStream_t st);
// out[0..2] += unknowns, aligned 2-cell groups and aligned 4-cell groups holding an unknown
// (n rounded down to a multiple of 4)
void flag_stats(const uint8_t* flags, long long n, unsigned long long* out, cudaStream_t st);
// device-side checkMatrixProperties; bad[0] incremented per violation
void check_rows(const Grid& g, const uint8_t* flags, const uint8_t* active, int dir,
                int n_dir_global, unsigned long long* bad, cudaStream_t st);


// ---------------------------------------------------------------- peer halo (oi_halo.cu)
// Store this slab's bottom plane into the lower neighbour's top ghost plane and its
// top plane into the upper neighbour's bottom ghost plane (peer pointers, either
// may be null), then write `seq` into the neighbours' flag words.
void halo_push(const void* src_lo, void* dst_lo, const void* src_hi, void* dst_hi, size_t plane_bytes,
               unsigned int* flag_lo, unsigned int* flag_hi, unsigned int seq, unsigned int* counter,
               int n_sm, cudaStream_t st);
// Spin until the local flag words reach `seq` (used when stream memory operations
// are not available; the default wait is cuStreamWaitValue32).
void halo_wait_spin(const unsigned int* flag_a, const unsigned int* flag_b, unsigned int seq,
                    cudaStream_t st);

}  // namespace oi
