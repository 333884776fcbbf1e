// One-CTA V-cycle over the smallest levels of the multigrid hierarchy.
//
// Levels of a few thousand cells cost one kernel launch (~4.5 us even inside a CUDA
// graph) per sweep, residual, restriction and prolongation -- ~11 launches per level and
// cycle for microseconds of arithmetic.  tail_cycle() runs ALL the levels below a size
// limit in one CTA: same operators, same sweep order and weights as coarse_cycle() in
// oi_solver.cu, a barrier where that has a kernel boundary, the fields in global memory
// (L1/L2 resident at this size; plain loads and stores, which a CTA sees coherently
// across its barriers).  Every tail level holds the whole box in z (single slab, or a
// replicated level), so periodic neighbours are wrapped indices instead of ghost planes.
//
// The cycle is written against a `step(n, f)` primitive -- run f(idx) for every idx in
// [0, n), then a barrier -- so that the CUDA kernel (oi_coarse.cu: CtaStep) and the host
// emulation used by the CPU tests (tests/cpu_emul/tail_emul.cu: a plain loop) execute the
// same code.  Every step reads fields it does not write (or updates cells independently),
// so the two executions are equivalent.
// Replaces the coarse end of HYPRE SMG/PFMG's V-cycle (reference call site
// src/props/TortuosityHypre.cpp:671-683).
#pragma once
#include "oi_kernels.h"

namespace oi {

__host__ __device__ __forceinline__ int tail_imin(int a, int b) { return a < b ? a : b; }

// MODE 1: out = x + w (b - A x)/dg ; MODE 2: out = b - A x   (coarse_stencil_kernel's arithmetic)
template <int MODE>
__host__ __device__ __forceinline__ void tail_stencil_cell(const CoarseLevel& L, const mg_t* x, const mg_t* b,
                                                           mg_t* out, mg_t w, int idx) {
    const int nx = L.nx, ny = L.ny, nz = L.nz, plane = (int)L.plane;
    const bool px = (L.periodic & PER_X) != 0, py = (L.periodic & PER_Y) != 0, pz = (L.periodic & PER_Z) != 0;
    const float d = L.dg[idx];
    mg_t o = 0;
    if (d > 0.f) {
        const int i = idx % nx, j = (idx / nx) % ny, k = idx / plane;
        const mg_t c = x[idx];
        mg_t acc = (mg_t)d * c;
        const float cxp = L.cxp[idx], cyp = L.cyp[idx], czp = L.czp[idx];
        if (cxp != 0.f && (i + 1 < nx || px)) acc -= (mg_t)cxp * x[(i + 1 < nx) ? idx + 1 : idx - i];
        if (cyp != 0.f && (j + 1 < ny || py)) acc -= (mg_t)cyp * x[(j + 1 < ny) ? idx + nx : idx - j * nx];
        if (czp != 0.f && (k + 1 < nz || pz)) acc -= (mg_t)czp * x[(k + 1 < nz) ? idx + plane : idx - k * plane];
        if (i > 0 || px) {
            const int im = (i > 0) ? idx - 1 : idx + (nx - 1);
            const float cm = L.cxp[im];
            if (cm != 0.f) acc -= (mg_t)cm * x[im];
        }
        if (j > 0 || py) {
            const int jm = (j > 0) ? idx - nx : idx + (ny - 1) * nx;
            const float cm = L.cyp[jm];
            if (cm != 0.f) acc -= (mg_t)cm * x[jm];
        }
        if (k > 0 || pz) {
            const int km = (k > 0) ? idx - plane : idx + (nz - 1) * plane;
            const float cm = L.czp[km];
            if (cm != 0.f) acc -= (mg_t)cm * x[km];
        }
        if (MODE == 1) o = c + w * (b[idx] - acc) / (mg_t)d;
        else           o = b[idx] - acc;
    }
    out[idx] = o;
}

// rhs of coarse cell I = sum of the finer residual over its (fx x fy x fz) block
__host__ __device__ __forceinline__ void tail_restrict_cell(const CoarseLevel& f, const mg_t* res, const CoarseLevel& c,
                                                            mg_t* bc, int I) {
    const int fplane = (int)f.plane, cplane = (int)c.plane;
    const int ci = I % c.nx, cj = (I / c.nx) % c.ny, ck = I / cplane;
    const int i0 = ci * f.fx, j0 = cj * f.fy, k0 = ck * f.fz;
    const int i1 = tail_imin(i0 + f.fx, f.nx), j1 = tail_imin(j0 + f.fy, f.ny), k1 = tail_imin(k0 + f.fz, f.nz);
    mg_t s = 0;
    for (int k = k0; k < k1; ++k)
        for (int j = j0; j < j1; ++j)
            for (int i = i0; i < i1; ++i)
                s += res[k * fplane + j * f.nx + i];
    bc[I] = s;
}

// x += P ec on non-empty cells (piecewise-constant prolongation)
__host__ __device__ __forceinline__ void tail_prolong_cell(const CoarseLevel& L, mg_t* x, const mg_t* ec, int cnx,
                                                           int cny, int idx) {
    if (L.dg[idx] > 0.f) {
        const int plane = (int)L.plane;
        const int i = idx % L.nx, j = (idx / L.nx) % L.ny, k = idx / plane;
        const int ci = (L.fx == 2) ? (i >> 1) : i, cj = (L.fy == 2) ? (j >> 1) : j, ck = (L.fz == 2) ? (k >> 1) : k;
        x[idx] += ec[(ck * cny + cj) * cnx + ci];
    }
}

// The cycle over levels Lv[0 .. n_levels-1] (fields wherever their pointers say: global
// memory, or the shared-memory copies of tail_cycle_staged).  Leaves the correction of the
// first level in Lv[0].x.
template <class Step>
__host__ __device__ __forceinline__ void tail_cycle_levels(const CoarseLevel* Lv, int n_levels, const mg_t* wv, int wdeg,
                                                           const mg_t* wcv, int wcdeg, Step step) {
    mg_t* res[TAIL_MAX_LEVELS];          // where each level's current iterate lives (x or t)
    // down: pre-smooth from a zero guess, residual, restrict
    for (int l = 0; l < n_levels; ++l) {
        const CoarseLevel& L = Lv[l];
        const int n = L.nz * (int)L.plane;
        const bool last = (l + 1 == n_levels);
        const mg_t* w = last ? wcv : wv;
        const int deg = last ? wcdeg : wdeg;
        mg_t* cur = L.t;
        mg_t* oth = L.x;
        {
            const mg_t w0 = w[0];
            step(n, [&](int idx) {
                const float d = L.dg[idx];
                cur[idx] = (d > 0.f) ? w0 * L.b[idx] / (mg_t)d : (mg_t)0;
            });
        }
        for (int s = 1; s < deg; ++s) {
            const mg_t ws = w[s];
            step(n, [&](int idx) { tail_stencil_cell<1>(L, cur, L.b, oth, ws, idx); });
            mg_t* t = cur; cur = oth; oth = t;
        }
        if (!last) {
            const CoarseLevel& C = Lv[l + 1];
            step(n, [&](int idx) { tail_stencil_cell<2>(L, cur, L.b, oth, (mg_t)0, idx); });
            step(C.nz * (int)C.plane, [&](int I) { tail_restrict_cell(L, oth, C, C.b, I); });
        }
        res[l] = cur;
    }
    // up: add the correction, post-smooth with the mirrored weights
    for (int l = n_levels - 2; l >= 0; --l) {
        const CoarseLevel& L = Lv[l];
        const CoarseLevel& C = Lv[l + 1];
        const int n = L.nz * (int)L.plane;
        mg_t* cur = res[l];
        mg_t* oth = (cur == L.x) ? L.t : L.x;
        const mg_t* ec = res[l + 1];
        step(n, [&](int idx) { tail_prolong_cell(L, cur, ec, C.nx, C.ny, idx); });
        for (int s = 0; s < wdeg; ++s) {
            const mg_t ws = wv[wdeg - 1 - s];
            step(n, [&](int idx) { tail_stencil_cell<1>(L, cur, L.b, oth, ws, idx); });
            mg_t* t = cur; cur = oth; oth = t;
        }
        res[l] = cur;
    }
    // the caller reads the first tail level's correction from its x
    if (res[0] != Lv[0].x) {
        const CoarseLevel& L = Lv[0];
        const mg_t* src = res[0];
        step(L.nz * (int)L.plane, [&](int idx) { L.x[idx] = src[idx]; });
    }
}

// Fields in global memory (any mg_t, any size the level limit allows).
template <class Step>
__host__ __device__ __forceinline__ void tail_cycle(const TailArgs& a, Step step) {
    tail_cycle_levels(a.L, a.n_levels, a.w, a.deg, a.wc, a.deg_c, step);
}

// Fields staged in shared memory: a dependent global load costs the CTA an L2 round trip
// (~0.7 us; ~3.7 us per step measured with the global variant, 110 us per cycle on the
// sample image's tail) where shared memory answers in tens of nanoseconds.  `buf` holds
// TAIL_MAX_LEVELS level descriptors followed by 7 float arrays per level (cxp, cyp, czp,
// dg, x, b, t): tail_staged_bytes().  Couplings, diagonals and the first level's rhs are
// copied in, the cycle runs on the copies, the first level's correction is copied out.
__host__ __device__ __forceinline__ size_t tail_desc_floats() {
    return (sizeof(CoarseLevel) * TAIL_MAX_LEVELS + 15) / 16 * 4;
}
__host__ __device__ __forceinline__ size_t tail_staged_bytes(const TailArgs& a) {
    size_t cells = 0;
    for (int l = 0; l < a.n_levels; ++l) cells += (size_t)a.L[l].nz * (size_t)a.L[l].plane;
    return (tail_desc_floats() + 7 * cells) * sizeof(float);
}

template <class Step>
__host__ __device__ __forceinline__ void tail_cycle_staged(const TailArgs& a, Step step, float* buf) {
    static_assert(sizeof(mg_t) == sizeof(float) || sizeof(mg_t) == sizeof(double), "mg_t");
    CoarseLevel* Ls = reinterpret_cast<CoarseLevel*>(buf);
    float* fld = buf + tail_desc_floats();
    step(a.n_levels, [&](int l) {
        CoarseLevel L = a.L[l];
        size_t off = 0;
        for (int q = 0; q < l; ++q) off += 7 * (size_t)a.L[q].nz * (size_t)a.L[q].plane;
        const size_t n = (size_t)L.nz * (size_t)L.plane;
        float* f = fld + off;
        L.cxp = f; L.cyp = f + n; L.czp = f + 2 * n; L.dg = f + 3 * n;
        L.x = reinterpret_cast<mg_t*>(f + 4 * n); L.b = reinterpret_cast<mg_t*>(f + 5 * n);
        L.t = reinterpret_cast<mg_t*>(f + 6 * n);
        Ls[l] = L;
    });
    for (int l = 0; l < a.n_levels; ++l) {
        const CoarseLevel& G = a.L[l];
        const CoarseLevel& L = Ls[l];
        const bool first = (l == 0);
        step(G.nz * (int)G.plane, [&](int idx) {
            L.cxp[idx] = G.cxp[idx]; L.cyp[idx] = G.cyp[idx]; L.czp[idx] = G.czp[idx]; L.dg[idx] = G.dg[idx];
            if (first) L.b[idx] = G.b[idx];
        });
    }
    tail_cycle_levels(Ls, a.n_levels, a.w, a.deg, a.wc, a.deg_c, step);
    {
        const CoarseLevel& G = a.L[0];
        const CoarseLevel& L = Ls[0];
        step(G.nz * (int)G.plane, [&](int idx) { G.x[idx] = L.x[idx]; });
    }
}

}  // namespace oi
