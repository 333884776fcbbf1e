// Coarse-level kernels of the geometric V-cycle (levels >= 1).
//
// Coarse operators are aggregation-Galerkin: a coarse cell is a (fx x fy x fz)
// block of finer cells, the coupling across a coarse face is the sum of the
// finer couplings crossing it, the diagonal is the sum of outward couplings plus
// the children's sink terms (Dirichlet neighbours, solid faces of the cell problem).
// Couplings and sinks along axis a are scaled by 1/f_a: the aggregate sums
// (area) faces but the cell distance grows by f_a, so this is what turns
// piecewise-constant aggregation into the rediscretised cell-centred operator (1/2 for
// every axis under full coarsening).  The diagonal is kept split by axis (dgx, dgy,
// dgz; build-time only) so that the next level can scale each share on its own:
// anisotropic cells are semicoarsened (only the strongly coupled axes) until the
// couplings even out.  This replaces HYPRE SMG/PFMG's coarse
// operator build (reference call site src/props/TortuosityHypre.cpp:671-681).
#include <cuda_fp16.h>

#include "oi_kernels.h"
#include "oi_coarse_tail.cuh"

namespace oi {

namespace {

// diagonal of a fine row (couplings + sink terms: Dirichlet neighbours, and for the
// cell problem the faces towards the solid)
__device__ __forceinline__ void diag0_by_axis(uint8_t f, const Grid& g, double& dx, double& dy, double& dz) {
    if (g.diag_full > 0.0) { dx = 2.0 * g.cx; dy = 2.0 * g.cy; dz = 2.0 * g.cz; return; }
    dx = g.cx * (double)__popc(f & 0x03u); dy = g.cy * (double)__popc(f & 0x0cu); dz = g.cz * (double)__popc(f & 0x30u);
}

// level 1 from connectivity bytes
__global__ void __launch_bounds__(256)
build_from_flags_kernel(Grid g, const uint8_t* __restrict__ flags, CoarseLevel c, int fx, int fy,
                        int fz, double scx, double scy, double scz) {
    const long long nc = (long long)c.nz * c.plane;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long I = (long long)blockIdx.x * blockDim.x + threadIdx.x; I < nc; I += stride) {
        const int ci = (int)(I % c.nx);
        const int cj = (int)((I / c.nx) % c.ny);
        const int ck = (int)(I / c.plane);
        const int i0 = ci * fx, j0 = cj * fy, k0 = ck * fz;
        const int i1 = min(i0 + fx, g.nx), j1 = min(j0 + fy, g.ny), k1 = min(k0 + fz, g.nz);
        double sx = 0.0, sy = 0.0, sz = 0.0, sdx = 0.0, sdy = 0.0, sdz = 0.0, inx = 0.0, iny = 0.0, inz = 0.0;
        for (int k = k0; k < k1; ++k)
            for (int j = j0; j < j1; ++j)
                for (int i = i0; i < i1; ++i) {
                    const long long idx = (long long)k * g.plane + (long long)j * g.nx + i;
                    const uint8_t f = flags[idx];
                    if (!(f & F_UNK)) continue;
                    double ddx, ddy, ddz;
                    diag0_by_axis(f, g, ddx, ddy, ddz);
                    sdx += ddx; sdy += ddy; sdz += ddz;
                    // +x, +y, +z couplings to UNKNOWN neighbours (Dirichlet
                    // neighbours stay in the diagonal as sink terms)
                    // (a set bit at the box edge means a periodic neighbour on the far side)
                    const long long ixp = (i + 1 < g.nx) ? idx + 1 : idx - i;
                    const long long iyp = (j + 1 < g.ny) ? idx + g.nx : idx - (long long)j * g.nx;
                    if ((f & F_XP) && (flags[ixp] & F_UNK)) {
                        if (i + 1 < i1) inx += g.cx; else sx += g.cx;
                    }
                    if ((f & F_YP) && (flags[iyp] & F_UNK)) {
                        if (j + 1 < j1) iny += g.cy; else sy += g.cy;
                    }
                    if ((f & F_ZP) && (flags[idx + g.plane] & F_UNK)) {
                        if (k + 1 < k1) inz += g.cz; else sz += g.cz;
                    }
                }
        const float dx = (float)(scx * (sdx - 2.0 * inx)), dy = (float)(scy * (sdy - 2.0 * iny)),
                    dz = (float)(scz * (sdz - 2.0 * inz));
        c.cxp[I] = (float)(scx * sx);
        c.cyp[I] = (float)(scy * sy);
        c.czp[I] = (float)(scz * sz);
        c.dgx[I] = dx; c.dgy[I] = dy; c.dgz[I] = dz;
        c.dg[I] = dx + dy + dz;
    }
}

__global__ void __launch_bounds__(256)
build_from_coarse_kernel(CoarseLevel f, CoarseLevel c, double scx, double scy, double scz) {
    const long long nc = (long long)c.nz * c.plane;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const int fx = f.fx, fy = f.fy, fz = f.fz;
    for (long long I = (long long)blockIdx.x * blockDim.x + threadIdx.x; I < nc; I += stride) {
        const int ci = (int)(I % c.nx);
        const int cj = (int)((I / c.nx) % c.ny);
        const int ck = (int)(I / c.plane);
        const int i0 = ci * fx, j0 = cj * fy, k0 = ck * fz;
        const int i1 = min(i0 + fx, f.nx), j1 = min(j0 + fy, f.ny), k1 = min(k0 + fz, f.nz);
        double sx = 0.0, sy = 0.0, sz = 0.0, sdx = 0.0, sdy = 0.0, sdz = 0.0, inx = 0.0, iny = 0.0, inz = 0.0;
        for (int k = k0; k < k1; ++k)
            for (int j = j0; j < j1; ++j)
                for (int i = i0; i < i1; ++i) {
                    const long long idx = (long long)k * f.plane + (long long)j * f.nx + i;
                    sdx += (double)f.dgx[idx]; sdy += (double)f.dgy[idx]; sdz += (double)f.dgz[idx];
                    const double ax = (double)f.cxp[idx], ay = (double)f.cyp[idx], az = (double)f.czp[idx];
                    if (i + 1 < i1) inx += ax; else sx += ax;
                    if (j + 1 < j1) iny += ay; else sy += ay;
                    if (k + 1 < k1) inz += az; else sz += az;
                }
        const float dx = (float)(scx * (sdx - 2.0 * inx)), dy = (float)(scy * (sdy - 2.0 * iny)),
                    dz = (float)(scz * (sdz - 2.0 * inz));
        c.cxp[I] = (float)(scx * sx);
        c.cyp[I] = (float)(scy * sy);
        c.czp[I] = (float)(scz * sz);
        c.dgx[I] = dx; c.dgy[I] = dy; c.dgz[I] = dz;
        c.dg[I] = dx + dy + dz;
    }
}

// MODE 1: out = x + w (b - A x)/dg ; MODE 2: out = b - A x      (all in mg_t arithmetic)
// 64 x 4 cells per CTA in (x, y), blockIdx.z strides over planes (32-bit index math only).
template <int MODE>
__global__ void __launch_bounds__(256)
coarse_stencil_kernel(CoarseLevel L, const mg_t* __restrict__ x, const mg_t* __restrict__ b,
                      mg_t* __restrict__ out, mg_t w) {
    const int i = blockIdx.x * 64 + (threadIdx.x & 63);
    const int j = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (i >= L.nx || j >= L.ny) return;
    const long long col = (long long)j * L.nx + i;
    for (int k = blockIdx.z; k < L.nz; k += gridDim.z) {
        const long long idx = (long long)k * L.plane + col;
        const float d = L.dg[idx];
        mg_t o = 0;
        if (d > 0.f) {
            const mg_t c = x[idx];
            mg_t acc = (mg_t)d * c;
            // couplings are zero across domain faces, so guarded loads suffice
            const float cxp = L.cxp[idx], cyp = L.cyp[idx], czp = L.czp[idx];
            // (a coupling across a box face exists only when the box is periodic)
            if (cxp != 0.f) acc -= (mg_t)cxp * x[(i + 1 < L.nx) ? idx + 1 : idx - i];
            if (cyp != 0.f) acc -= (mg_t)cyp * x[(j + 1 < L.ny) ? idx + L.nx : idx - (long long)j * L.nx];
            if (czp != 0.f) acc -= (mg_t)czp * x[idx + L.plane];
            if (i > 0 || (L.periodic & PER_X)) {
                const long long im = (i > 0) ? idx - 1 : idx + (L.nx - 1);
                const float cm = L.cxp[im];
                if (cm != 0.f) acc -= (mg_t)cm * x[im];
            }
            if (j > 0 || (L.periodic & PER_Y)) {
                const long long jm = (j > 0) ? idx - L.nx : idx + (long long)(L.ny - 1) * L.nx;
                const float cm = L.cyp[jm];
                if (cm != 0.f) acc -= (mg_t)cm * x[jm];
            }
            {   // k-1 may be the ghost plane (coefficients exchanged at setup)
                const float cm = L.czp[idx - L.plane];
                if (cm != 0.f) acc -= (mg_t)cm * x[idx - L.plane];
            }
            if (MODE == 1) o = c + w * (b[idx] - acc) / (mg_t)d;
            else           o = b[idx] - acc;
        }
        out[idx] = o;
    }
}

// Vectorised variant for fp32 levels whose nx is a multiple of 4: each thread owns
// four x-adjacent cells (16-byte loads of x, b and of every coefficient array), a
// CTA covers 64 x 16 cells of one plane.  Same arithmetic as the scalar kernel.
// HALF: the four coefficient arrays are read from their exact half-precision copies (8-byte loads).
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4h(const unsigned short* p) {
    const uint2 r = *reinterpret_cast<const uint2*>(p);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float ld1h(const unsigned short* p) {
    return __half2float(*reinterpret_cast<const __half*>(p));
}

// HALO (z-slabs): the boundary planes are visited last; the blocks that own plane 0 / nz-1 wait for the
// neighbour's plane on the flag words of `hin` before they read the ghost plane, store their part of the
// new plane into the neighbour's ghost plane (MODE 1) and publish `hout.seq` -- no push kernel and no
// stream wait between two sweeps of a distributed level.
template <int MODE, bool HALF, bool HALO>
__global__ void __launch_bounds__(256)
coarse_stencil_vec4_kernel(CoarseLevel L, const float* __restrict__ x, const float* __restrict__ b,
                           float* __restrict__ out, float w, HaloIn hin, HaloOut hout) {
    const int i = blockIdx.x * 64 + (threadIdx.x & 15) * 4;
    const int j = blockIdx.y * 16 + (threadIdx.x >> 4);
    const bool valid = i < L.nx && j < L.ny;
    if (!HALO && !valid) return;
    const long long col = (long long)j * L.nx + i;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int kk = blockIdx.z; kk < L.nz; kk += gridDim.z) {
        int k = kk;
        if (HALO) {
            if (L.nz >= 2) k = (kk < L.nz - 2) ? kk + 1 : (kk == L.nz - 2 ? 0 : L.nz - 1);
            const bool wlo = hin.flag_lo && k == 0, whi = hin.flag_hi && k == L.nz - 1;
            if (wlo || whi) {
                if (threadIdx.x == 0) {
                    if (wlo) halo_spin(hin.flag_lo, hin.seq);
                    if (whi) halo_spin(hin.flag_hi, hin.seq);
                }
                __syncthreads();
            }
        }
        if (valid) {
            const long long idx = (long long)k * L.plane + col;
            const float4 d = HALF ? ld4h(L.hd + idx) : ld4(L.dg + idx);
            float4 o = zero4;
            if (d.x > 0.f || d.y > 0.f || d.z > 0.f || d.w > 0.f) {
                const float4 c = ld4(x + idx);
                const float4 cxp = HALF ? ld4h(L.hx + idx) : ld4(L.cxp + idx);
                const float4 cyp = HALF ? ld4h(L.hy + idx) : ld4(L.cyp + idx);
                const float4 czp = HALF ? ld4h(L.hz + idx) : ld4(L.czp + idx);
                const bool px = (L.periodic & PER_X) != 0, py = (L.periodic & PER_Y) != 0;
                const long long iw = (i > 0) ? idx - 1 : idx + (L.nx - 1);
                const long long js = (j > 0) ? idx - L.nx : idx + (long long)(L.ny - 1) * L.nx;
                const long long jn = (j + 1 < L.ny) ? idx + L.nx : idx - (long long)j * L.nx;
                const float cxw = (i > 0 || px) ? (HALF ? ld1h(L.hx + iw) : L.cxp[iw]) : 0.f;
                const float xw = (i > 0 || px) ? x[iw] : 0.f;
                const float xe = (i + 4 < L.nx) ? x[idx + 4] : (px ? x[idx - i] : 0.f);
                const float4 cym = (j > 0 || py) ? (HALF ? ld4h(L.hy + js) : ld4(L.cyp + js)) : zero4;
                const float4 ys = (j > 0 || py) ? ld4(x + js) : zero4;
                const float4 yn = (j + 1 < L.ny || py) ? ld4(x + jn) : zero4;
                const float4 czm = HALF ? ld4h(L.hz + idx - L.plane) : ld4(L.czp + idx - L.plane);   // k-1 may be the ghost plane
                const float4 zd = ld4(x + idx - L.plane);
                const float4 zu = ld4(x + idx + L.plane);
                float4 acc;
                acc.x = d.x * c.x - cxp.x * c.y - cxw * xw - cyp.x * yn.x - cym.x * ys.x - czp.x * zu.x - czm.x * zd.x;
                acc.y = d.y * c.y - cxp.y * c.z - cxp.x * c.x - cyp.y * yn.y - cym.y * ys.y - czp.y * zu.y - czm.y * zd.y;
                acc.z = d.z * c.z - cxp.z * c.w - cxp.y * c.y - cyp.z * yn.z - cym.z * ys.z - czp.z * zu.z - czm.z * zd.z;
                acc.w = d.w * c.w - cxp.w * xe - cxp.z * c.z - cyp.w * yn.w - cym.w * ys.w - czp.w * zu.w - czm.w * zd.w;
                const float4 bb = ld4(b + idx);
                if (MODE == 1) {
                    o.x = d.x > 0.f ? c.x + w * (bb.x - acc.x) / d.x : 0.f;
                    o.y = d.y > 0.f ? c.y + w * (bb.y - acc.y) / d.y : 0.f;
                    o.z = d.z > 0.f ? c.z + w * (bb.z - acc.z) / d.z : 0.f;
                    o.w = d.w > 0.f ? c.w + w * (bb.w - acc.w) / d.w : 0.f;
                } else {
                    o.x = d.x > 0.f ? bb.x - acc.x : 0.f;
                    o.y = d.y > 0.f ? bb.y - acc.y : 0.f;
                    o.z = d.z > 0.f ? bb.z - acc.z : 0.f;
                    o.w = d.w > 0.f ? bb.w - acc.w : 0.f;
                }
            }
            *reinterpret_cast<float4*>(out + idx) = o;
            if (HALO && MODE == 1) {
                if (hout.dst_lo && k == 0) *reinterpret_cast<float4*>(static_cast<float*>(hout.dst_lo) + col) = o;
                if (hout.dst_hi && k == L.nz - 1) *reinterpret_cast<float4*>(static_cast<float*>(hout.dst_hi) + col) = o;
            }
        }
        if (HALO && MODE == 1) {
            const unsigned int tiles = gridDim.x * gridDim.y;
            if (hout.flag_lo && k == 0) halo_publish(hout.counter + 0, tiles, hout.flag_lo, nullptr, hout.seq);
            if (hout.flag_hi && k == L.nz - 1) halo_publish(hout.counter + 1, tiles, nullptr, hout.flag_hi, hout.seq);
        }
    }
}

// ---------------------------------------------------------------------------------------
// Two smoothing sweeps of a stored-coefficient level in one pass (the level-1 counterpart of
// oi_level0_pair.cu):  v = x + w1 (b - A x)/d ,  out = v + w2 (b - A v)/d.
// Level 1 is 1/8 of the cells but, with degree 8 per leg, a fifth of a PCG iteration; a sweep moves
// x 4 + b 4 + four half coefficients 8 + out 4 = 20 B per cell, so two sweeps per pass halve that.
// A CTA owns a 64 x 16 tile and marches along z.  Per plane kk it
//   A) forms v(kk) on its tile (4 cells per thread, operands straight from global memory / L1 as in the
//      single-sweep kernel, own z column in registers) and on the one-cell rim around it (164 cells, one per
//      thread of the first 164), and keeps the last three planes of v in shared memory;
//   B) forms out(kk-1) from v(kk-2 .. kk): own column from registers, x / y neighbours from shared memory,
//      coefficients of plane kk-1 re-read (L1 hits: the same thread loaded them one trip earlier).
// One CTA barrier per plane.  Single z-slab (or a level every rank holds whole), non-periodic box,
// nx % 4 == 0, fp32 vectors, half coefficient copies present.
constexpr int CPX = 64, CPY = 16, CPW = CPX + 8, CVH = CPY + 2;     // tile, smem pitch [4 | 64 | 4], rows with rim

// A x at one cell from its seven values and the couplings (same expression order as coarse_stencil_kernel)
__device__ __forceinline__ float coarse_ax(float d, float c, float cxp, float xe, float cxw, float xw, float cyp, float yn,
                                           float cym, float ys, float czp, float zu, float czm, float zd) {
    return d * c - cxp * xe - cxw * xw - cyp * yn - cym * ys - czp * zu - czm * zd;
}

__global__ void __launch_bounds__(256)
coarse_pair_kernel(CoarseLevel L, const float* __restrict__ x, const float* __restrict__ b, float* __restrict__ out,
                   float w1, float w2, int zchunk) {
    __shared__ __align__(16) float vs[3][CVH][CPW];
    const int tid = threadIdx.x;
    const int i0 = blockIdx.x * CPX, j0 = blockIdx.y * CPY;
    const int i = i0 + (tid & 15) * 4, j = j0 + (tid >> 4);
    const int k0 = blockIdx.z * zchunk, k1 = min(k0 + zchunk, L.nz);
    const bool inb = i < L.nx && j < L.ny;
    const long long col = inb ? (long long)j * L.nx + i : 0;
    const int vrow = (tid >> 4) + 1, vcol = 4 + (tid & 15) * 4;
    // rim duty: rows j0-1 and j0+16 (2 x 64 cells), then columns i0-1 and i0+64 (2 x 18 cells)
    int ri = 0, rj = 0, rrow = 0, rcol = 0;
    const bool rim = tid < 2 * CPX + 2 * CVH;
    if (tid < 2 * CPX) { rrow = (tid >= CPX) ? CVH - 1 : 0; rcol = 4 + (tid % CPX); }
    else if (rim) { const int t = tid - 2 * CPX; rrow = t % CVH; rcol = (t >= CVH) ? 4 + CPX : 3; }
    ri = i0 - 4 + rcol; rj = j0 - 1 + rrow;
    const bool rin = rim && ri >= 0 && ri < L.nx && rj >= 0 && rj < L.ny;
    const long long rcolg = rin ? (long long)rj * L.nx + ri : 0;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    // v at one cell of plane k from global memory (scalar path, for the rim)
    auto v_cell = [&](long long idx, int ci, int cj) -> float {
        const float d = ld1h(L.hd + idx);
        if (!(d > 0.f)) return 0.f;
        const float c = x[idx];
        const float xe = (ci + 1 < L.nx) ? x[idx + 1] : 0.f, xw = (ci > 0) ? x[idx - 1] : 0.f;
        const float yn = (cj + 1 < L.ny) ? x[idx + L.nx] : 0.f, ys = (cj > 0) ? x[idx - L.nx] : 0.f;
        const float cxw = (ci > 0) ? ld1h(L.hx + idx - 1) : 0.f, cym = (cj > 0) ? ld1h(L.hy + idx - L.nx) : 0.f;
        const float ax = coarse_ax(d, c, ld1h(L.hx + idx), xe, cxw, xw, ld1h(L.hy + idx), yn, cym, ys, ld1h(L.hz + idx),
                                   x[idx + L.plane], ld1h(L.hz + idx - L.plane), x[idx - L.plane]);
        return c + w1 * (b[idx] - ax) / d;
    };

    float4 x_m = zero4, x_c = zero4, v3 = zero4, v2 = zero4;
    if (inb) {
        x_m = ld4(x + (long long)(k0 - 2) * L.plane + col + (k0 - 2 < -1 ? L.plane : 0));   // plane k0-2 (or the ghost plane again: unused)
        x_c = ld4(x + (long long)(k0 - 1) * L.plane + col);
    }
    for (int kk = k0 - 1; kk <= k1; ++kk) {
        __syncthreads();                       // v(kk-1) complete; the slot of v(kk-3) is free
        float (*V)[CPW] = vs[((kk % 3) + 3) % 3];
        const bool plane_in = kk >= 0 && kk < L.nz;
        float4 x_p = zero4, vA = zero4;
        if (inb && kk + 1 <= L.nz) x_p = ld4(x + (long long)(kk + 1) * L.plane + col);
        // ---- A: v(kk) on the tile (own four cells) ...
        if (plane_in && inb) {
            const long long idx = (long long)kk * L.plane + col;
            const float4 d = ld4h(L.hd + idx);
            if (d.x > 0.f || d.y > 0.f || d.z > 0.f || d.w > 0.f) {
                const float4 cxp = ld4h(L.hx + idx), cyp = ld4h(L.hy + idx), czp = ld4h(L.hz + idx);
                const float4 czm = ld4h(L.hz + idx - L.plane);
                const float cxw = (i > 0) ? ld1h(L.hx + idx - 1) : 0.f;
                const float xw = (i > 0) ? x[idx - 1] : 0.f;
                const float xe = (i + 4 < L.nx) ? x[idx + 4] : 0.f;
                const float4 cym = (j > 0) ? ld4h(L.hy + idx - L.nx) : zero4;
                const float4 ys = (j > 0) ? ld4(x + idx - L.nx) : zero4;
                const float4 yn = (j + 1 < L.ny) ? ld4(x + idx + L.nx) : zero4;
                const float4 bb = ld4(b + idx);
                const float4 c = x_c;
                const float ax0 = coarse_ax(d.x, c.x, cxp.x, c.y, cxw, xw, cyp.x, yn.x, cym.x, ys.x, czp.x, x_p.x, czm.x, x_m.x);
                const float ax1 = coarse_ax(d.y, c.y, cxp.y, c.z, cxp.x, c.x, cyp.y, yn.y, cym.y, ys.y, czp.y, x_p.y, czm.y, x_m.y);
                const float ax2 = coarse_ax(d.z, c.z, cxp.z, c.w, cxp.y, c.y, cyp.z, yn.z, cym.z, ys.z, czp.z, x_p.z, czm.z, x_m.z);
                const float ax3 = coarse_ax(d.w, c.w, cxp.w, xe, cxp.z, c.z, cyp.w, yn.w, cym.w, ys.w, czp.w, x_p.w, czm.w, x_m.w);
                vA.x = d.x > 0.f ? c.x + w1 * (bb.x - ax0) / d.x : 0.f;
                vA.y = d.y > 0.f ? c.y + w1 * (bb.y - ax1) / d.y : 0.f;
                vA.z = d.z > 0.f ? c.z + w1 * (bb.z - ax2) / d.z : 0.f;
                vA.w = d.w > 0.f ? c.w + w1 * (bb.w - ax3) / d.w : 0.f;
            }
        }
        *reinterpret_cast<float4*>(&V[vrow][vcol]) = vA;
        // ... and on the rim
        if (rim) V[rrow][rcol] = (plane_in && rin) ? v_cell((long long)kk * L.plane + rcolg, ri, rj) : 0.f;

        // ---- B: out(kk-1) from v(kk-2), v(kk-1) (shared memory: written one trip ago), v(kk) = vA
        const int k = kk - 1;
        if (k >= k0 && k < k1 && inb) {
            const float (*Vc)[CPW] = vs[((k % 3) + 3) % 3];
            const long long idx = (long long)k * L.plane + col;
            const float4 d = ld4h(L.hd + idx);
            float4 o = zero4;
            if (d.x > 0.f || d.y > 0.f || d.z > 0.f || d.w > 0.f) {
                const float4 cxp = ld4h(L.hx + idx), cyp = ld4h(L.hy + idx), czp = ld4h(L.hz + idx);
                const float4 czm = ld4h(L.hz + idx - L.plane);
                const float cxw = (i > 0) ? ld1h(L.hx + idx - 1) : 0.f;
                const float4 cym = (j > 0) ? ld4h(L.hy + idx - L.nx) : zero4;
                const float4 bb = ld4(b + idx);
                const float4 c = v2;
                const float xw = Vc[vrow][vcol - 1], xe = Vc[vrow][vcol + 4];
                const float4 ys = *reinterpret_cast<const float4*>(&Vc[vrow - 1][vcol]);
                const float4 yn = *reinterpret_cast<const float4*>(&Vc[vrow + 1][vcol]);
                const float ax0 = coarse_ax(d.x, c.x, cxp.x, c.y, cxw, xw, cyp.x, yn.x, cym.x, ys.x, czp.x, vA.x, czm.x, v3.x);
                const float ax1 = coarse_ax(d.y, c.y, cxp.y, c.z, cxp.x, c.x, cyp.y, yn.y, cym.y, ys.y, czp.y, vA.y, czm.y, v3.y);
                const float ax2 = coarse_ax(d.z, c.z, cxp.z, c.w, cxp.y, c.y, cyp.z, yn.z, cym.z, ys.z, czp.z, vA.z, czm.z, v3.z);
                const float ax3 = coarse_ax(d.w, c.w, cxp.w, xe, cxp.z, c.z, cyp.w, yn.w, cym.w, ys.w, czp.w, vA.w, czm.w, v3.w);
                o.x = d.x > 0.f ? c.x + w2 * (bb.x - ax0) / d.x : 0.f;
                o.y = d.y > 0.f ? c.y + w2 * (bb.y - ax1) / d.y : 0.f;
                o.z = d.z > 0.f ? c.z + w2 * (bb.z - ax2) / d.z : 0.f;
                o.w = d.w > 0.f ? c.w + w2 * (bb.w - ax3) / d.w : 0.f;
            }
            *reinterpret_cast<float4*>(out + idx) = o;
        }
        // rotate the register queues: x column and the own v column (v2 = v(kk-1) must be the value stage A stored)
        x_m = x_c; x_c = x_p;
        v3 = v2; v2 = vA;
    }
}

__global__ void __launch_bounds__(256)
to_half_kernel(const float* __restrict__ src, unsigned short* __restrict__ dst, long long n, unsigned long long* mismatch) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long bad = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float v = src[i];
        const __half h = __float2half_rn(v);
        if (__half2float(h) != v) ++bad;
        dst[i] = *reinterpret_cast<const unsigned short*>(&h);
    }
    bad = warp_sum_ll(bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(mismatch, (unsigned long long)bad);
}

// x += P * ec on non-empty cells (prolongation + correction between coarse levels)
__global__ void __launch_bounds__(256)
coarse_prolong_add_kernel(CoarseLevel L, mg_t* __restrict__ x, const mg_t* __restrict__ ec,
                          int cnx, int cny) {
    const int i = blockIdx.x * 64 + (threadIdx.x & 63);
    const int j = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (i >= L.nx || j >= L.ny) return;
    const int ci = (L.fx == 2) ? (i >> 1) : i, cj = (L.fy == 2) ? (j >> 1) : j;
    const long long col = (long long)j * L.nx + i;
    const long long ccol = (long long)cj * cnx + ci;
    const long long cplane = (long long)cnx * cny;
    for (int k = blockIdx.z; k < L.nz; k += gridDim.z) {
        const long long idx = (long long)k * L.plane + col;
        if (L.dg[idx] > 0.f) {
            const int ck = (L.fz == 2) ? (k >> 1) : k;
            x[idx] += ec[(long long)ck * cplane + ccol];
        }
    }
}

__global__ void __launch_bounds__(256)
coarse_jacobi_first_kernel(CoarseLevel L, const mg_t* __restrict__ b, mg_t* __restrict__ out, mg_t w) {
    const long long n = (long long)L.nz * L.plane;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += stride) {
        const float d = L.dg[idx];
        out[idx] = (d > 0.f) ? w * b[idx] / (mg_t)d : (mg_t)0;
    }
}

__global__ void __launch_bounds__(256)
coarse_restrict_kernel(CoarseLevel f, const mg_t* __restrict__ res, CoarseLevel c, mg_t* __restrict__ bc) {
    const int ci = blockIdx.x * 64 + (threadIdx.x & 63);
    const int cj = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (ci >= c.nx || cj >= c.ny) return;
    const int i0 = ci * f.fx, j0 = cj * f.fy;
    const int i1 = min(i0 + f.fx, f.nx), j1 = min(j0 + f.fy, f.ny);
    for (int ck = blockIdx.z; ck < c.nz; ck += gridDim.z) {
        const int k0 = ck * f.fz, k1 = min(k0 + f.fz, f.nz);
        mg_t s = 0;
        for (int k = k0; k < k1; ++k)
            for (int j = j0; j < j1; ++j)
                for (int i = i0; i < i1; ++i)
                    s += res[(long long)k * f.plane + (long long)j * f.nx + i];
        bc[(long long)ck * c.plane + (long long)cj * c.nx + ci] = s;
    }
}

// ---------------------------------------------------------------- coarse tail
// (the cycle itself is in oi_coarse_tail.cuh so that tests can run it on the host)
constexpr int TAIL_NT = 1024;

// one data-parallel step of the CTA: f(idx) for every idx in [0, n), then a barrier
struct CtaStep {
    template <class F>
    __device__ __forceinline__ void operator()(int n, F f) const {
        for (int idx = threadIdx.x; idx < n; idx += TAIL_NT) f(idx);
        __syncthreads();
    }
};

__global__ void __launch_bounds__(TAIL_NT, 1)
coarse_tail_kernel(TailArgs a) {
    tail_cycle(a, CtaStep());
}

// fields staged in dynamic shared memory (fp32 multigrid vectors only)
__global__ void __launch_bounds__(TAIL_NT, 1)
coarse_tail_staged_kernel(TailArgs a) {
    extern __shared__ float4 tail_smem[];
    tail_cycle_staged(a, CtaStep(), reinterpret_cast<float*>(tail_smem));
}

// the 16-byte path needs fp32 vectors and every row / plane start 16-byte aligned
static bool vec4_ok(const CoarseLevel& L) {
    return sizeof(mg_t) == 4 && (L.nx & 3) == 0 && L.nx >= 16;
}

inline int blocks_for(long long n) {
    long long b = (n + 255) / 256;
    if (b > 148 * 8) b = 148 * 8;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace

void coarse_build_from_flags(const Grid& g, const uint8_t* flags, int, int, const CoarseLevel& c,
                             int fx, int fy, int fz, cudaStream_t st) {
    build_from_flags_kernel<<<blocks_for((long long)c.nz * c.plane), 256, 0, st>>>(g, flags, c, fx, fy, fz,
                                                                                  1.0 / fx, 1.0 / fy, 1.0 / fz);
}

void coarse_build_from_coarse(const CoarseLevel& f, const CoarseLevel& c, cudaStream_t st) {
    build_from_coarse_kernel<<<blocks_for((long long)c.nz * c.plane), 256, 0, st>>>(f, c, 1.0 / f.fx, 1.0 / f.fy,
                                                                                   1.0 / f.fz);
}

void coarse_to_half(const float* src, unsigned short* dst, long long n, unsigned long long* mismatch, cudaStream_t st) {
    to_half_kernel<<<blocks_for(n), 256, 0, st>>>(src, dst, n, mismatch);
}

bool coarse_half_applicable(const CoarseLevel& L) { return vec4_ok(L); }

void coarse_jacobi_first(const CoarseLevel& L, const mg_t* b, mg_t* out, double w, cudaStream_t st) {
    coarse_jacobi_first_kernel<<<blocks_for((long long)L.nz * L.plane), 256, 0, st>>>(L, b, out, (mg_t)w);
}

static dim3 grid3(const CoarseLevel& L) {
    int gz = L.nz < 128 ? L.nz : 128;
    return dim3((L.nx + 63) / 64, (L.ny + 3) / 4, gz > 0 ? gz : 1);
}

static dim3 grid_vec4(const CoarseLevel& L) {
    int gz = L.nz < 128 ? L.nz : 128;
    return dim3((L.nx + 63) / 64, (L.ny + 15) / 16, gz > 0 ? gz : 1);
}

template <int MODE>
static void launch_vec4(const CoarseLevel& L, const mg_t* x, const mg_t* b, mg_t* out, float w, const HaloIn* hin,
                        const HaloOut* hout, cudaStream_t st) {
    const float* xf = reinterpret_cast<const float*>(x);
    const float* bf = reinterpret_cast<const float*>(b);
    float* of = reinterpret_cast<float*>(out);
    const HaloIn hi = hin ? *hin : HaloIn{};
    const HaloOut ho = hout ? *hout : HaloOut{};
    const bool halo = hi.flag_lo || hi.flag_hi || ho.flag_lo || ho.flag_hi;
    const dim3 g = grid_vec4(L);
    if (L.hd) {
        if (halo) coarse_stencil_vec4_kernel<MODE, true, true><<<g, 256, 0, st>>>(L, xf, bf, of, w, hi, ho);
        else coarse_stencil_vec4_kernel<MODE, true, false><<<g, 256, 0, st>>>(L, xf, bf, of, w, hi, ho);
    } else {
        if (halo) coarse_stencil_vec4_kernel<MODE, false, true><<<g, 256, 0, st>>>(L, xf, bf, of, w, hi, ho);
        else coarse_stencil_vec4_kernel<MODE, false, false><<<g, 256, 0, st>>>(L, xf, bf, of, w, hi, ho);
    }
}

// whether coarse_smooth / coarse_residual run the kernel that can wait for / push the ghost planes itself
bool coarse_halo_supported(const CoarseLevel& L) { return vec4_ok(L); }

void coarse_smooth(const CoarseLevel& L, const mg_t* x, const mg_t* b, mg_t* out, double w, cudaStream_t st,
                   const HaloIn* hin, const HaloOut* hout) {
    if (vec4_ok(L)) launch_vec4<1>(L, x, b, out, (float)w, hin, hout, st);
    else coarse_stencil_kernel<1><<<grid3(L), 256, 0, st>>>(L, x, b, out, (mg_t)w);
}

void coarse_prolong_add(const CoarseLevel& L, mg_t* x, const CoarseLevel& next, const mg_t* ec, cudaStream_t st) {
    coarse_prolong_add_kernel<<<grid3(L), 256, 0, st>>>(L, x, ec, next.nx, next.ny);
}

void coarse_residual(const CoarseLevel& L, const mg_t* x, const mg_t* b, mg_t* out, cudaStream_t st, const HaloIn* hin) {
    if (vec4_ok(L)) launch_vec4<2>(L, x, b, out, 0.f, hin, nullptr, st);
    else coarse_stencil_kernel<2><<<grid3(L), 256, 0, st>>>(L, x, b, out, (mg_t)0);
}

// two sweeps per pass (coarse_pair_kernel): big single-slab levels with half coefficient copies only
bool coarse_pair_supported(const CoarseLevel& L) {
    return vec4_ok(L) && L.hd != nullptr && L.periodic == 0 && (long long)L.plane * L.nz >= (1LL << 21);
}
void coarse_smooth_pair(const CoarseLevel& L, const mg_t* x, const mg_t* b, mg_t* out, double w1, double w2,
                        cudaStream_t st) {
    const int zc = L.nz >= 256 ? 64 : 32;
    dim3 grid((L.nx + CPX - 1) / CPX, (L.ny + CPY - 1) / CPY, (L.nz + zc - 1) / zc);
    coarse_pair_kernel<<<grid, 256, 0, st>>>(L, reinterpret_cast<const float*>(x), reinterpret_cast<const float*>(b),
                                             reinterpret_cast<float*>(out), (float)w1, (float)w2, zc);
}

void coarse_tail_cycle(const TailArgs& a, bool staged, cudaStream_t st) {
    // 227 KB of shared memory per CTA on sm_100a; stay a little below
    constexpr size_t SMEM_LIMIT = 220 * 1024;
    const size_t bytes = tail_staged_bytes(a);
    if (staged && sizeof(mg_t) == sizeof(float) && bytes <= SMEM_LIMIT) {
        static unsigned long long configured = 0;
        if (first_use_on_this_device(configured))
            cudaFuncSetAttribute(coarse_tail_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
        coarse_tail_staged_kernel<<<1, TAIL_NT, bytes, st>>>(a);
    } else {
        coarse_tail_kernel<<<1, TAIL_NT, 0, st>>>(a);
    }
}

void coarse_restrict(const CoarseLevel& f, const mg_t* res, const CoarseLevel& c, mg_t* bc, cudaStream_t st) {
    coarse_restrict_kernel<<<grid3(c), 256, 0, st>>>(f, res, c, bc);
}

}  // namespace oi
