// The two entry points of the reference that already are C-ABI (Fortran bind(c)), under their
// reference names and signatures, running on the device:
//
//   tortuosity_fillmtx   src/props/TortuosityHypreFill_F.H:47-68  (body: TortuosityHypreFill.F90:44-314)
//   tortuosity_remspot   src/props/Tortuosity_filcc_F.H:65-67     (body: Tortuosity_filcc.F90:88-177)
//
// A maintainer of the reference who links this library instead of the two Fortran objects gets
// the same arrays back (rows, rhs bit-exact; xinit bit-exact: the ramp is evaluated in the
// Fortran's operation order without fused multiply-adds).  Both take HOST pointers in the
// Fortran calling convention (every scalar by reference, arrays with their own lo/hi bounds,
// x fastest), copy the box to the device, run one kernel and copy the result back; they are the
// per-tile seam of the reference (MFIter loops, TortuosityHypre.cpp:270-290, 590-632), not the
// fast path -- the fast path never materialises the 7 coefficients (oi_build_mask + oi_solve).
// No CPU fallback: without a CUDA device they print the error and abort, as the Fortran's
// `error stop` would.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "../../include/openimpala_b200.h"

namespace {

struct Box3 { int lo[3], hi[3]; };
__host__ __device__ inline long long box_len(const Box3& b, int d) { return (long long)b.hi[d] - b.lo[d] + 1; }
__host__ __device__ inline long long box_pts(const Box3& b) {
    return (b.hi[0] < b.lo[0] || b.hi[1] < b.lo[1] || b.hi[2] < b.lo[2]) ? 0 : box_len(b, 0) * box_len(b, 1) * box_len(b, 2);
}
__device__ __forceinline__ bool box_has(const Box3& b, int i, int j, int k) {
    return i >= b.lo[0] && i <= b.hi[0] && j >= b.lo[1] && j <= b.hi[1] && k >= b.lo[2] && k <= b.hi[2];
}
__device__ __forceinline__ long long box_idx(const Box3& b, int i, int j, int k) {
    return (long long)(i - b.lo[0]) + box_len(b, 0) * ((long long)(j - b.lo[1]) + box_len(b, 1) * (long long)(k - b.lo[2]));
}

[[noreturn]] void die(const char* who, const char* what) {
    std::fprintf(stderr, "%s: %s (the B200 build has no CPU fallback)\n", who, what);
    std::abort();
}
#define REF_CUDA(who, expr)                                             \
    do {                                                                \
        cudaError_t e_ = (expr);                                        \
        if (e_ != cudaSuccess) die(who, cudaGetErrorString(e_));        \
    } while (0)

// ---------------------------------------------------------------- tortuosity_fillmtx
struct FillArgs {
    Box3 pb, mb, bx, dom;
    double c[3], vlo, vhi;
    int phase, dir;
};

// One thread per cell of bx.  A cell outside the bounds of `p` / `mask` reads as inactive (the
// reference would read out of bounds there; its callers always pass one ghost cell).
__global__ void __launch_bounds__(256)
fillmtx_box_kernel(FillArgs A, const int* __restrict__ p, const int* __restrict__ mask, double* __restrict__ a,
                   double* __restrict__ rhs, double* __restrict__ xinit, long long n) {
    const long long m = (long long)blockIdx.x * 256 + threadIdx.x;
    if (m >= n) return;
    const long long lx = box_len(A.bx, 0), ly = box_len(A.bx, 1);
    const int i = A.bx.lo[0] + (int)(m % lx);
    const int j = A.bx.lo[1] + (int)((m / lx) % ly);
    const int k = A.bx.lo[2] + (int)(m / (lx * ly));
    auto act = [&](int ii, int jj, int kk) -> bool {            // phase == id and mask == cell_active (F90:111, 126)
        if (!box_has(A.pb, ii, jj, kk) || !box_has(A.mb, ii, jj, kk)) return false;
        return p[box_idx(A.pb, ii, jj, kk)] == A.phase && mask[box_idx(A.mb, ii, jj, kk)] == 1;
    };
    double row[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    double* out = a + 7 * m;
    // inactive: mask == cell_inactive (0) or another phase (F90:111-118)
    const bool self = box_has(A.pb, i, j, k) && box_has(A.mb, i, j, k) && p[box_idx(A.pb, i, j, k)] == A.phase &&
                      mask[box_idx(A.mb, i, j, k)] != 0;
    if (!self) {
        row[0] = 1.0;
#pragma unroll
        for (int s = 0; s < 7; ++s) out[s] = row[s];
        rhs[m] = 0.0;
        xinit[m] = 0.0;
        return;
    }
    double diag = 0.0;                                           // F90:126-166, slots C,-x,+x,-y,+y,-z,+z
    if (act(i - 1, j, k)) { row[1] = -A.c[0]; diag = __dadd_rn(diag, A.c[0]); }
    if (act(i + 1, j, k)) { row[2] = -A.c[0]; diag = __dadd_rn(diag, A.c[0]); }
    if (act(i, j - 1, k)) { row[3] = -A.c[1]; diag = __dadd_rn(diag, A.c[1]); }
    if (act(i, j + 1, k)) { row[4] = -A.c[1]; diag = __dadd_rn(diag, A.c[1]); }
    if (act(i, j, k - 1)) { row[5] = -A.c[2]; diag = __dadd_rn(diag, A.c[2]); }
    if (act(i, j, k + 1)) { row[6] = -A.c[2]; diag = __dadd_rn(diag, A.c[2]); }
    row[0] = diag;
    const double small_real = 1.0e-15;
    if (fabs(diag) < small_real) {                               // F90:172-181: decouple, skip the Dirichlet overwrite
        out[0] = 1.0;
#pragma unroll
        for (int s = 1; s < 7; ++s) out[s] = 0.0;
        rhs[m] = 0.0;
        xinit[m] = 0.0;
        return;
    }
    double b = 0.0;
    bool on_dirichlet = false;
    const int idx = A.dir == 0 ? i : (A.dir == 1 ? j : k);
    if (A.dir >= 0 && A.dir <= 2) {                              // F90:192-228
        if (idx == A.dom.lo[A.dir]) { on_dirichlet = true; b = A.vlo; }
        else if (idx == A.dom.hi[A.dir]) { on_dirichlet = true; b = A.vhi; }
    }
    if (on_dirichlet) {
        row[0] = 1.0;
#pragma unroll
        for (int s = 1; s < 7; ++s) row[s] = 0.0;
    }
#pragma unroll
    for (int s = 0; s < 7; ++s) out[s] = row[s];
    rhs[m] = b;
    // F90:233-262: ramp unless the (non-Dirichlet) diagonal is exactly 1; such a cell keeps what the
    // caller put into xinit
    if (fabs(row[0] - 1.0) > small_real || on_dirichlet) {
        if (A.dir >= 0 && A.dir <= 2) {
            const int extent = A.dom.hi[A.dir] - A.dom.lo[A.dir];
            const double factor = (fabs((double)extent) < small_real) ? 0.0 : 1.0 / (double)extent;
            // vlo + (vhi - vlo) * (idx - domlo) * factor, left to right, no contraction
            const double t = __dmul_rn(__dmul_rn(__dadd_rn(A.vhi, -A.vlo), (double)(idx - A.dom.lo[A.dir])), factor);
            xinit[m] = __dadd_rn(A.vlo, t);
        } else {
            xinit[m] = 0.5 * (A.vlo + A.vhi);
        }
    }
}

// ---------------------------------------------------------------- tortuosity_remspot
// The Fortran loop updates q in place in (k, j, i) order, so a voxel sees its -x/-y/-z neighbours
// inside bx already filtered.  Same fixed point on flip flags as oi_remspot (oi_mask.cu), on an
// arbitrary box: flip_c = no in-domain neighbour equals q_c, earlier neighbours inside bx read as
// their filtered value (current estimate), everything else as stored.
struct SpotArgs { Box3 qb, bx, dom; };

__device__ __forceinline__ int flipped(int v) { return v == 0 ? 1 : 0; }      // F90:163-167

__global__ void __launch_bounds__(256)
remspot_box_round_kernel(SpotArgs A, const int* __restrict__ q, const unsigned char* __restrict__ fcur,
                         unsigned char* __restrict__ fnext, long long n, int* changed) {
    const long long m = (long long)blockIdx.x * 256 + threadIdx.x;
    bool any = false;
    if (m < n) {
        const long long lx = box_len(A.bx, 0), ly = box_len(A.bx, 1);
        const int i = A.bx.lo[0] + (int)(m % lx);
        const int j = A.bx.lo[1] + (int)((m / lx) % ly);
        const int k = A.bx.lo[2] + (int)(m / (lx * ly));
        const int c = q[box_idx(A.qb, i, j, k)];
        auto val = [&](int ii, int jj, int kk, bool earlier) -> int {
            int v = q[box_idx(A.qb, ii, jj, kk)];
            if (earlier && box_has(A.bx, ii, jj, kk) && fcur[box_idx(A.bx, ii, jj, kk)]) v = flipped(v);
            return v;
        };
        bool connected = false;                                   // a neighbour outside the domain never matches (F90:113-152)
        if (i != A.dom.lo[0] && box_has(A.qb, i - 1, j, k)) connected |= (val(i - 1, j, k, true) == c);
        if (i != A.dom.hi[0] && box_has(A.qb, i + 1, j, k)) connected |= (val(i + 1, j, k, false) == c);
        if (j != A.dom.lo[1] && box_has(A.qb, i, j - 1, k)) connected |= (val(i, j - 1, k, true) == c);
        if (j != A.dom.hi[1] && box_has(A.qb, i, j + 1, k)) connected |= (val(i, j + 1, k, false) == c);
        if (k != A.dom.lo[2] && box_has(A.qb, i, j, k - 1)) connected |= (val(i, j, k - 1, true) == c);
        if (k != A.dom.hi[2] && box_has(A.qb, i, j, k + 1)) connected |= (val(i, j, k + 1, false) == c);
        const unsigned char f = connected ? 0 : 1;
        any = (f != fcur[m]);
        fnext[m] = f;
    }
    if (__any_sync(0xffffffffu, any) && (threadIdx.x & 31) == 0) *changed = 1;
}

__global__ void __launch_bounds__(256)
remspot_box_apply_kernel(SpotArgs A, int* __restrict__ q, const unsigned char* __restrict__ f, long long n) {
    const long long m = (long long)blockIdx.x * 256 + threadIdx.x;
    if (m >= n || !f[m]) return;
    const long long lx = box_len(A.bx, 0), ly = box_len(A.bx, 1);
    const int i = A.bx.lo[0] + (int)(m % lx);
    const int j = A.bx.lo[1] + (int)((m / lx) % ly);
    const int k = A.bx.lo[2] + (int)(m / (lx * ly));
    const long long at = box_idx(A.qb, i, j, k);
    q[at] = flipped(q[at]);
}

Box3 make_box(const int* lo, const int* hi) {
    Box3 b;
    for (int d = 0; d < 3; ++d) { b.lo[d] = lo[d]; b.hi[d] = hi[d]; }
    return b;
}

template <typename T>
struct DevBuf {
    T* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
};

}  // namespace

extern "C" {

void tortuosity_fillmtx(double* a, double* rhs, double* xinit, const int* nval, const int* p, const int* p_lo,
                        const int* p_hi, const int* active_mask, const int* mask_lo, const int* mask_hi,
                        const int* bxlo, const int* bxhi, const int* domlo, const int* domhi, const double* dxinv,
                        const double* vlo, const double* vhi, const int* phase, const int* dir,
                        const int* debug_print_level) {
    static const char* who = "tortuosity_fillmtx";
    (void)debug_print_level;
    FillArgs A{};
    A.pb = make_box(p_lo, p_hi); A.mb = make_box(mask_lo, mask_hi);
    A.bx = make_box(bxlo, bxhi); A.dom = make_box(domlo, domhi);
    for (int d = 0; d < 3; ++d) A.c[d] = dxinv[d];
    A.vlo = *vlo; A.vhi = *vhi; A.phase = *phase; A.dir = *dir;
    const long long n = box_pts(A.bx);
    if (nval && (long long)*nval != n) die(who, "nval does not match the number of cells of the box");   // F90:84-88
    if (n == 0) return;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) die(who, "no CUDA device");
    const long long np = box_pts(A.pb), nm = box_pts(A.mb);
    DevBuf<int> dp, dm;
    DevBuf<double> da, drhs, dx;
    REF_CUDA(who, cudaMalloc(&dp.p, sizeof(int) * (size_t)np));
    REF_CUDA(who, cudaMalloc(&dm.p, sizeof(int) * (size_t)nm));
    REF_CUDA(who, cudaMalloc(&da.p, sizeof(double) * 7 * (size_t)n));
    REF_CUDA(who, cudaMalloc(&drhs.p, sizeof(double) * (size_t)n));
    REF_CUDA(who, cudaMalloc(&dx.p, sizeof(double) * (size_t)n));
    REF_CUDA(who, cudaMemcpy(dp.p, p, sizeof(int) * (size_t)np, cudaMemcpyHostToDevice));
    REF_CUDA(who, cudaMemcpy(dm.p, active_mask, sizeof(int) * (size_t)nm, cudaMemcpyHostToDevice));
    REF_CUDA(who, cudaMemcpy(dx.p, xinit, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
    fillmtx_box_kernel<<<(unsigned)((n + 255) / 256), 256>>>(A, dp.p, dm.p, da.p, drhs.p, dx.p, n);
    REF_CUDA(who, cudaGetLastError());
    REF_CUDA(who, cudaMemcpy(a, da.p, sizeof(double) * 7 * (size_t)n, cudaMemcpyDeviceToHost));
    REF_CUDA(who, cudaMemcpy(rhs, drhs.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
    REF_CUDA(who, cudaMemcpy(xinit, dx.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
}

void tortuosity_remspot(int* q, const int* q_lo, const int* q_hi, const int* ncomp, const int* bxlo,
                        const int* bxhi, const int* domlo, const int* domhi) {
    static const char* who = "tortuosity_remspot";
    if (*ncomp < 1) die(who, "Input array q must have at least comp_phase components.");                // F90:108
    SpotArgs A{};
    A.qb = make_box(q_lo, q_hi); A.bx = make_box(bxlo, bxhi); A.dom = make_box(domlo, domhi);
    const long long n = box_pts(A.bx), nq = box_pts(A.qb);
    if (n == 0) return;
    for (int d = 0; d < 3; ++d)
        if (A.bx.lo[d] < A.qb.lo[d] || A.bx.hi[d] > A.qb.hi[d]) die(who, "the box is not inside the bounds of q");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) die(who, "no CUDA device");
    DevBuf<int> dq, dchg;
    DevBuf<unsigned char> fa, fb;
    REF_CUDA(who, cudaMalloc(&dq.p, sizeof(int) * (size_t)nq));       // component comp_phase = 1 is the first block
    REF_CUDA(who, cudaMalloc(&dchg.p, sizeof(int)));
    REF_CUDA(who, cudaMalloc(&fa.p, (size_t)n));
    REF_CUDA(who, cudaMalloc(&fb.p, (size_t)n));
    REF_CUDA(who, cudaMemcpy(dq.p, q, sizeof(int) * (size_t)nq, cudaMemcpyHostToDevice));
    REF_CUDA(who, cudaMemset(fa.p, 0, (size_t)n));
    const unsigned nb = (unsigned)((n + 255) / 256);
    // exact after as many rounds as the longest chain of mutually dependent isolated voxels
    const long long max_rounds = n + 2;
    bool settled = false;
    for (long long r = 0; r < max_rounds; ++r) {
        REF_CUDA(who, cudaMemset(dchg.p, 0, sizeof(int)));
        remspot_box_round_kernel<<<nb, 256>>>(A, dq.p, fa.p, fb.p, n, dchg.p);
        REF_CUDA(who, cudaGetLastError());
        int changed = 0;
        REF_CUDA(who, cudaMemcpy(&changed, dchg.p, sizeof(int), cudaMemcpyDeviceToHost));
        unsigned char* t = fa.p; fa.p = fb.p; fb.p = t;
        if (!changed) { settled = true; break; }
    }
    if (!settled) die(who, "flip flags did not reach their fixed point");
    remspot_box_apply_kernel<<<nb, 256>>>(A, dq.p, fa.p, n);
    REF_CUDA(who, cudaGetLastError());
    REF_CUDA(who, cudaMemcpy(q, dq.p, sizeof(int) * (size_t)nq, cudaMemcpyDeviceToHost));
}

}  // extern "C"
