// Level-0 (voxel grid) matrix-free stencil kernels for sm_100a: fallback variants.
//
// The operator is the reference's 7-point finite-volume matrix
// (src/props/TortuosityHypreFill.F90:96-228) with the identity rows (inactive
// cells, Dirichlet planes) eliminated; coefficients are rebuilt from one
// connectivity byte per cell instead of 7 stored doubles.
//
// The default kernels live in oi_level0_ring.cu (shared-memory ring).  This file
// holds the two independent cross-check / fallback variants, which make no
// alignment assumptions and select every face by its connectivity bit:
//   * z-march: 64x8 xy tile per CTA marching along z; every thread owns one (x,y)
//     column and keeps planes k-1,k,k+1 of its column in registers, the +-x,+-y
//     neighbours come from a double-buffered shared-memory copy of plane k.  Used
//     when nx is not a multiple of 4 and for the fused-prolongation sweep.
//   * gather: one thread per cell straight from global/L1/L2.
// All kernels are templated on the element type: double for the Krylov operator
// apply, mg_t for the multigrid sweeps.
#include <cstdlib>

#include "oi_kernels.h"

namespace oi {

namespace {

constexpr int TX = 64, TY = 8;           // CTA tile (cells)
constexpr int NTHREADS = TX * TY;        // 512

template <typename T>
__device__ __forceinline__ T diag_of(uint8_t f, const Grid& g) { return row_diag<T>(f, g); }

// A*u at one cell from centre + 6 neighbour values.  Written as a sum of
// differences (row sum 0, checkMatrixProperties TortuosityHypre.cpp:969-971).
template <typename T>
__device__ __forceinline__ T stencil_au(uint8_t f, const Grid& g, T c, T xm, T xp, T ym, T yp, T zm, T zp) {
    const T z0 = (T)0;
    T ax = ((f & F_XM) ? (c - xm) : z0) + ((f & F_XP) ? (c - xp) : z0);
    T ay = ((f & F_YM) ? (c - ym) : z0) + ((f & F_YP) ? (c - yp) : z0);
    T az = ((f & F_ZM) ? (c - zm) : z0) + ((f & F_ZP) ? (c - zp) : z0);
    T au = (T)g.cx * ax + (T)g.cy * ay + (T)g.cz * az;
    // cell problem: faces without a coupling still carry their 1/dx^2 on the diagonal
    if (g.diag_full > 0.0)
        au += ((T)g.diag_full - ((T)g.cx * (T)__popc(f & 0x03u) + (T)g.cy * (T)__popc(f & 0x0cu) +
                                 (T)g.cz * (T)__popc(f & 0x30u))) * c;
    return au;
}

template <typename T>
struct CoarseRef {       // coarse correction added on the fly (prolongation)
    const T* ec;         // coarse vector (ghost planes allowed), pointer at plane 0
    int cnx, cny;        // coarse dims
    int fx, fy, fz;      // coarsening factors (1 or 2)
    int z0;              // local plane 0 global index (for fz alignment)
};

template <typename T, bool ADDC>
__device__ __forceinline__ T load_val(const T* __restrict__ u, const uint8_t* __restrict__ flags,
                                      long long idx, int i, int j, int k, const CoarseRef<T>& cr) {
    T v = u[idx];
    if (ADDC) {
        if (flags[idx] & F_UNK) {
            int ci = (cr.fx == 2) ? (i >> 1) : i;
            int cj = (cr.fy == 2) ? (j >> 1) : j;
            // k may be -1 (ghost): floor division for fz == 2
            int ck = (cr.fz == 2) ? ((k + 2) >> 1) - 1 : k;
            v += cr.ec[((long long)ck * cr.cny + cj) * cr.cnx + ci];
        }
    }
    return v;
}

}  // namespace

// MODE: 0 = APPLY   out = scale * A u                      (K3)
//       1 = SMOOTH  out = u + w * (b - A u) / diag         (K5, weighted Jacobi)
//       2 = RESTRICT rc[I] = sum_children (b - A u)        (K6, fused residual)
// ADDC: u is read as u + P*ec (fused prolongation/correction)
// DOT : APPLY -> sum u*out ; SMOOTH -> sum b*out  (fused PCG dot, K4)
template <typename T, int MODE, bool ADDC, bool DOT>
__global__ void __launch_bounds__(NTHREADS)
l0_zmarch_kernel(Grid g, const uint8_t* __restrict__ flags, const T* __restrict__ u,
                 const T* __restrict__ b, T* __restrict__ out, T w, CoarseRef<T> cr, int zchunk,
                 double* red_partials, unsigned int* red_counter, double* red_out) {
    __shared__ T tile[2][TY + 2][TX + 2];

    // warp = 16(x) x 2(y) cells; 4 x 4 warps per CTA
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int tx = (warp & 3) * 16 + (lane & 15);
    const int ty = (warp >> 2) * 2 + (lane >> 4);
    const int i = blockIdx.x * TX + tx;
    const int j = blockIdx.y * TY + ty;
    const bool inb = (i < g.nx) && (j < g.ny);
    const int k0 = blockIdx.z * zchunk;
    const int k1 = min(k0 + zchunk, g.nz);

    // halo duty: each thread may own one x-halo and one y-halo cell of the tile
    const bool hx_l = (tx == 0), hx_r = (tx == TX - 1);
    const bool hy_t = (ty == 0), hy_b = (ty == TY - 1);
    // (a partial edge tile of a periodic box is handled below: the last in-box column /
    // row takes its wrapped neighbour straight from global memory)
    const int hi_x = (i < g.nx) ? (hx_l ? wrap_lo(i, g.nx, g.periodic & PER_X) : wrap_hi(i, g.nx, g.periodic & PER_X)) : -1;
    const int hj_y = (j < g.ny) ? (hy_t ? wrap_lo(j, g.ny, g.periodic & PER_Y) : wrap_hi(j, g.ny, g.periodic & PER_Y)) : -1;
    const bool hx_ok = (hx_l || hx_r) && (j < g.ny) && hi_x >= 0;
    const bool hy_ok = (hy_t || hy_b) && (i < g.nx) && hj_y >= 0;
    // last in-box column / row of a partial tile: its +x / +y neighbour is not a halo slot
    const bool edge_x = (g.periodic & PER_X) && i == g.nx - 1 && !hx_r && j < g.ny;
    const bool edge_y = (g.periodic & PER_Y) && j == g.ny - 1 && !hy_b && i < g.nx;

    const long long col = (long long)j * g.nx + i;
    const long long colhx = (long long)j * g.nx + hi_x;
    const long long colhy = (long long)hj_y * g.nx + i;

    // register pipeline over planes
    T v_m = 0, v_c = 0, v_p = 0, v_n = 0;
    T hx_c = 0, hy_c = 0, hx_n = 0, hy_n = 0;
    uint8_t f_c = 0, f_n = 0;
    T b_c = 0, b_n = 0;

    if (inb) {
        v_m = load_val<T, ADDC>(u, flags, (long long)(k0 - 1) * g.plane + col, i, j, k0 - 1, cr);
        v_c = load_val<T, ADDC>(u, flags, (long long)k0 * g.plane + col, i, j, k0, cr);
        v_p = load_val<T, ADDC>(u, flags, (long long)(k0 + 1) * g.plane + col, i, j, k0 + 1, cr);
        f_c = flags[(long long)k0 * g.plane + col];
        if (MODE != 0) b_c = b[(long long)k0 * g.plane + col];
    }
    if (hx_ok) hx_c = load_val<T, ADDC>(u, flags, (long long)k0 * g.plane + colhx, hi_x, j, k0, cr);
    if (hy_ok) hy_c = load_val<T, ADDC>(u, flags, (long long)k0 * g.plane + colhy, i, hj_y, k0, cr);

    double dot_acc = 0.0;
    T zpair = 0;         // RESTRICT: residual carried from the even plane
    int buf = 0;

    for (int k = k0; k < k1; ++k) {
        // 1. publish plane k
        tile[buf][ty + 1][tx + 1] = v_c;
        if (hx_l || hx_r) tile[buf][ty + 1][hx_l ? 0 : TX + 1] = hx_c;
        if (hy_t || hy_b) tile[buf][hy_t ? 0 : TY + 1][tx + 1] = hy_c;

        // 2. prefetch plane k+2 (own column) and plane k+1 (halo, flags, rhs)
        const bool more = (k + 1 < k1);
        if (inb) {
            // plane k+2 <= nz is inside the ghost-padded allocation when k+1 < nz
            if (more) {
                v_n = load_val<T, ADDC>(u, flags, (long long)(k + 2) * g.plane + col, i, j, k + 2, cr);
                f_n = flags[(long long)(k + 1) * g.plane + col];
                if (MODE != 0) b_n = b[(long long)(k + 1) * g.plane + col];
            }
        }
        if (more) {
            if (hx_ok) hx_n = load_val<T, ADDC>(u, flags, (long long)(k + 1) * g.plane + colhx, hi_x, j, k + 1, cr);
            if (hy_ok) hy_n = load_val<T, ADDC>(u, flags, (long long)(k + 1) * g.plane + colhy, i, hj_y, k + 1, cr);
        }
        __syncthreads();

        // 3. compute plane k
        T res = 0;  // MODE 2 residual
        if (inb) {
            T o = 0;
            if (f_c & F_UNK) {
                const T xm = tile[buf][ty + 1][tx];
                T xp = tile[buf][ty + 1][tx + 2];
                const T ym = tile[buf][ty][tx + 1];
                T yp = tile[buf][ty + 2][tx + 1];
                if (edge_x) xp = load_val<T, ADDC>(u, flags, (long long)k * g.plane + (long long)j * g.nx, 0, j, k, cr);
                if (edge_y) yp = load_val<T, ADDC>(u, flags, (long long)k * g.plane + i, i, 0, k, cr);
                const T au = stencil_au<T>(f_c, g, v_c, xm, xp, ym, yp, v_m, v_p);
                if (MODE == 0) {
                    o = w * au;
                    if (DOT) dot_acc += (double)v_c * (double)o;
                } else if (MODE == 1) {
                    o = v_c + w * (b_c - au) / diag_of<T>(f_c, g);
                    if (DOT) dot_acc += (double)b_c * (double)o;
                } else {
                    res = b_c - au;
                }
            }
            if (MODE != 2) out[(long long)k * g.plane + col] = o;
        }
        if (MODE == 2) {
            // 2x2x2 (or fx x fy x fz) sum: x pair = lanes l, l^1 ; y pair = l, l^16
            T s = res;
            if (cr.fx == 2) s += __shfl_xor_sync(0xffffffffu, s, 1);
            if (cr.fy == 2) s += __shfl_xor_sync(0xffffffffu, s, 16);
            const int kg = g.z0 + k;  // global plane: pairs are aligned globally
            bool flush = true;
            if (cr.fz == 2) {
                if ((kg & 1) == 0) { zpair = s; flush = (k + 1 == k1); }
                else { s += zpair; zpair = 0; }
            }
            const bool writer = inb && ((cr.fx == 1) || ((i & 1) == 0)) &&
                                ((cr.fy == 1) || ((j & 1) == 0));
            if (flush && writer) {
                int ci = (cr.fx == 2) ? (i >> 1) : i;
                int cj = (cr.fy == 2) ? (j >> 1) : j;
                int ck = (cr.fz == 2) ? (((kg) >> 1) - (cr.z0 >> 1)) : k;
                out[((long long)ck * cr.cny + cj) * cr.cnx + ci] = s;
            }
        }

        // 4. rotate
        v_m = v_c; v_c = v_p; v_p = v_n;
        hx_c = hx_n; hy_c = hy_n;
        f_c = f_n; b_c = b_n;
        buf ^= 1;
    }

    if (DOT) {
        double v[1] = {dot_acc};
        grid_reduce<1>(v, red_partials, red_counter, red_out);
    }
}

// Simple gather variant (one thread per cell, neighbours straight from
// global/L1/L2).  Kept as an independent cross-check of the staged kernels.
template <typename T, int MODE, bool ADDC, bool DOT>
__global__ void __launch_bounds__(256)
l0_gather_kernel(Grid g, const uint8_t* __restrict__ flags, const T* __restrict__ u,
                 const T* __restrict__ b, T* __restrict__ out, T w, CoarseRef<T> cr,
                 double* red_partials, unsigned int* red_counter, double* red_out) {
    const int i = blockIdx.x * 64 + (threadIdx.x & 63);
    const int j = blockIdx.y * 4 + (threadIdx.x >> 6);
    const int k = blockIdx.z;
    double dot_acc = 0.0;
    if (i < g.nx && j < g.ny) {
        const long long idx = (long long)k * g.plane + (long long)j * g.nx + i;
        const uint8_t f = flags[idx];
        T o = 0;
        if (f & F_UNK) {
            const T z0 = 0;
            const T c = load_val<T, ADDC>(u, flags, idx, i, j, k, cr);
            // (set bits imply an in-box or wrapped neighbour)
            const int im = wrap_lo(i, g.nx, true), ip = wrap_hi(i, g.nx, true);
            const int jm = wrap_lo(j, g.ny, true), jp = wrap_hi(j, g.ny, true);
            const long long row = idx - i, colk = idx - (long long)j * g.nx;
            const T xm = (f & F_XM) ? load_val<T, ADDC>(u, flags, row + im, im, j, k, cr) : z0;
            const T xp = (f & F_XP) ? load_val<T, ADDC>(u, flags, row + ip, ip, j, k, cr) : z0;
            const T ym = (f & F_YM) ? load_val<T, ADDC>(u, flags, colk + (long long)jm * g.nx, i, jm, k, cr) : z0;
            const T yp = (f & F_YP) ? load_val<T, ADDC>(u, flags, colk + (long long)jp * g.nx, i, jp, k, cr) : z0;
            const T zm = (f & F_ZM) ? load_val<T, ADDC>(u, flags, idx - g.plane, i, j, k - 1, cr) : z0;
            const T zp = (f & F_ZP) ? load_val<T, ADDC>(u, flags, idx + g.plane, i, j, k + 1, cr) : z0;
            const T au = stencil_au<T>(f, g, c, xm, xp, ym, yp, zm, zp);
            if (MODE == 0) {
                o = w * au;
                if (DOT) dot_acc = (double)c * (double)o;
            } else if (MODE == 1) {
                const T bb = b[idx];
                o = c + w * (bb - au) / diag_of<T>(f, g);
                if (DOT) dot_acc = (double)bb * (double)o;
            } else {
                o = b[idx] - au;   // plain residual (restriction is a second kernel)
            }
        }
        out[idx] = o;
    }
    if (DOT) {
        double v[1] = {dot_acc};
        grid_reduce<1>(v, red_partials, red_counter, red_out);
    }
}

// First smoothing sweep from a zero guess: out = w * b / diag  (no stencil).
template <typename T>
__global__ void __launch_bounds__(256)
l0_jacobi_first_kernel(Grid g, const uint8_t* __restrict__ flags, const T* __restrict__ b,
                       T* __restrict__ out, T w, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += stride) {
        const uint8_t f = flags[idx];
        out[idx] = (f & F_UNK) ? w * b[idx] / diag_of<T>(f, g) : (T)0;
    }
}

// ---------------------------------------------------------------- launchers
int pick_zchunk(const Grid& g, int n_sm) {
    // enough CTAs for ~8 waves (tail effect), chunks of 16..64 planes (each chunk
    // re-reads 2 halo planes and refills its pipeline)
    const long long tiles = (long long)((g.nx + TX - 1) / TX) * ((g.ny + TY - 1) / TY);
    const long long target = (long long)n_sm * 32;
    long long chunks = (target + tiles - 1) / tiles;
    if (chunks < 1) chunks = 1;
    int zc = (int)((g.nz + chunks - 1) / chunks);
    // chunks of at most 128 planes on a single slab (919 vs 928 ms per 1024^3 step against 64), 64 on z-slabs, where
    // the first and last chunk are the ones that wait for the neighbours and should stay a small share
    int cap = (g.nz == g.nzg) ? 128 : 64;
    if (const char* e = getenv("OI_ZCHUNK")) { const int v = atoi(e); if (v >= 8 && v <= 1024) cap = v; }   // experiments
    if (zc < 16) zc = 16;
    if (zc > cap) zc = cap;
    zc = (zc + 1) & ~1;                            // even: restriction pairs stay in one chunk
    return zc;
}

template <typename T, int MODE, bool ADDC, bool DOT>
static void launch_zmarch(const L0Args& a, cudaStream_t st) {
    CoarseRef<T> cr{static_cast<const T*>(a.ec), a.cnx, a.cny, a.fx, a.fy, a.fz, a.g.z0};
    const int zc = pick_zchunk(a.g, a.n_sm);
    dim3 grid((a.g.nx + TX - 1) / TX, (a.g.ny + TY - 1) / TY, (a.g.nz + zc - 1) / zc);
    l0_zmarch_kernel<T, MODE, ADDC, DOT><<<grid, NTHREADS, 0, st>>>(
        a.g, a.flags, static_cast<const T*>(a.u), static_cast<const T*>(a.b), static_cast<T*>(a.out), (T)a.w,
        cr, zc, a.red_partials, a.red_counter, a.red_out);
}

template <typename T, int MODE, bool ADDC, bool DOT>
static void launch_gather(const L0Args& a, cudaStream_t st) {
    CoarseRef<T> cr{static_cast<const T*>(a.ec), a.cnx, a.cny, a.fx, a.fy, a.fz, a.g.z0};
    dim3 grid((a.g.nx + 63) / 64, (a.g.ny + 3) / 4, a.g.nz);
    l0_gather_kernel<T, MODE, ADDC, DOT><<<grid, 256, 0, st>>>(
        a.g, a.flags, static_cast<const T*>(a.u), static_cast<const T*>(a.b), static_cast<T*>(a.out), (T)a.w,
        cr, a.red_partials, a.red_counter, a.red_out);
}

long long l0_max_blocks(const Grid& g, int n_sm) {
    const int zc = pick_zchunk(g, n_sm);
    long long a = (long long)((g.nx + TX - 1) / TX) * ((g.ny + TY - 1) / TY) * ((g.nz + zc - 1) / zc);
    long long b = (long long)((g.nx + 63) / 64) * ((g.ny + 3) / 4) * g.nz;
    return a > b ? a : b;
}

// OI_TMA=1: the ring kernels stage their planes with cp.async.bulk.tensor (oi_level0_tma.cu) where that applies
static bool want_tma() {
    const char* e = getenv("OI_TMA");
    return e && e[0] == '1';
}

// variant: 0 = shared-memory ring (cp.async), 2 = register z-march, 1 = gather
void l0_apply(const L0Args& a, bool dot, int variant, cudaStream_t st) {           // fp64 fields
    if (variant == 0 && ring_supported(a, 0)) {
        if (want_tma() && tma_supported(a, 0) && tma_launch(a, 0, dot, st)) return;
        ring_launch(a, 0, dot, st);
    } else if (variant == 0 || variant == 2) {
        if (dot) launch_zmarch<double, 0, false, true>(a, st); else launch_zmarch<double, 0, false, false>(a, st);
    } else {
        if (dot) launch_gather<double, 0, false, true>(a, st); else launch_gather<double, 0, false, false>(a, st);
    }
}

void l0_smooth(const L0Args& a, bool addc, bool dot, int variant, cudaStream_t st) {   // mg_t fields
    if (variant == 0 && !addc && ring_supported(a, 1)) {
        if (want_tma() && tma_supported(a, 1) && tma_launch(a, 1, dot, st)) return;
        ring_launch(a, 1, dot, st);
    } else if (variant == 0 || variant == 2) {
        if (addc) { if (dot) launch_zmarch<mg_t, 1, true, true>(a, st); else launch_zmarch<mg_t, 1, true, false>(a, st); }
        else      { if (dot) launch_zmarch<mg_t, 1, false, true>(a, st); else launch_zmarch<mg_t, 1, false, false>(a, st); }
    } else {
        if (addc) { if (dot) launch_gather<mg_t, 1, true, true>(a, st); else launch_gather<mg_t, 1, true, false>(a, st); }
        else      { if (dot) launch_gather<mg_t, 1, false, true>(a, st); else launch_gather<mg_t, 1, false, false>(a, st); }
    }
}

void l0_residual_restrict(const L0Args& a, int variant, cudaStream_t st) {   // out = coarse rhs (mg_t)
    if (variant == 0 && ring_supported(a, 2)) ring_launch(a, 2, false, st);
    else launch_zmarch<mg_t, 2, false, false>(a, st);
}

void l0_residual(const L0Args& a, cudaStream_t st) {            // out = fine residual (mg_t)
    launch_gather<mg_t, 2, false, false>(a, st);
}

void l0_jacobi_first(const L0Args& a, cudaStream_t st) {        // mg_t fields
    const long long n = (long long)a.g.nz * a.g.plane;
    int blocks = a.n_sm * 8;
    long long need = (n + 255) / 256;
    if (need < blocks) blocks = (int)need;
    if (blocks < 1) blocks = 1;
    l0_jacobi_first_kernel<mg_t><<<blocks, 256, 0, st>>>(a.g, a.flags, static_cast<const mg_t*>(a.b),
                                                         static_cast<mg_t*>(a.out), (mg_t)a.w, n);
}

}  // namespace oi
