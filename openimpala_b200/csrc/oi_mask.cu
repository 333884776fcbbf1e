// Setup-side kernels: phase counting (K1), percolation mask by connected-
// component labelling (K2), connectivity bytes / rhs / initial guess (K7),
// boundary-flux reduction (K8), matrix-row export and invariant check.
//
// Reference restated here:
//   VolumeFraction::value          src/props/VolumeFraction.cpp:22-66
//   generateActivityMask           src/props/TortuosityHypre.cpp:394-558
//   parallelFloodFill              src/props/TortuosityHypre.cpp:297-389
//   tortuosity_fillmtx             src/props/TortuosityHypreFill.F90:44-314
//   global_fluxes                  src/props/TortuosityHypre.cpp:1000-1134
//   checkMatrixProperties          src/props/TortuosityHypre.cpp:896-982
//
// The reference floods the phase from every inlet-plane cell and from every
// outlet-plane cell by repeated sweeps and keeps the intersection.  The exact
// fixed point of that is "union of 6-connected components of {phase == id} that
// touch both planes", which is what the label-equivalence (union-find) kernels
// below compute in a constant number of passes.
#include <cstdlib>

#include "oi_kernels.h"

namespace oi {

namespace {

constexpr int BT = 256;

inline int nblocks(long long n, int n_sm, int per_sm = 8) {
    long long b = (n + BT - 1) / BT;
    const long long cap = (long long)n_sm * per_sm;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

// ------------------------------------------------------------------ K1
template <typename T>
__global__ void __launch_bounds__(BT)
count_kernel(const T* __restrict__ f, long long n, int phase, unsigned long long* out) {
    const long long stride = (long long)gridDim.x * BT;
    long long acc = 0;
    for (long long i = (long long)blockIdx.x * BT + threadIdx.x; i < n; i += stride)
        acc += ((int)f[i] == phase) ? 1 : 0;
    acc = warp_sum_ll(acc);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, (unsigned long long)acc);  // integer: exact
}

// 16 bytes per thread per step for the uint8 field
__global__ void __launch_bounds__(BT)
count_u8_vec_kernel(const uint8_t* __restrict__ f, long long n, int phase, unsigned long long* out) {
    const long long n16 = n >> 4;
    const long long stride = (long long)gridDim.x * BT;
    long long acc = 0;
    if (phase >= 0 && phase <= 255) {
        const unsigned int pat = 0x01010101u * (unsigned int)phase;
        const uint4* v = reinterpret_cast<const uint4*>(f);
        for (long long i = (long long)blockIdx.x * BT + threadIdx.x; i < n16; i += stride) {
            const uint4 w = v[i];
            // bytes equal to phase: xor -> zero byte; count zero bytes
            unsigned int a = w.x ^ pat, b = w.y ^ pat, c = w.z ^ pat, d = w.w ^ pat;
            // zero-byte mask: __vcmpeq4 returns 0xff per equal byte
            acc += (__popc(__vcmpeq4(a, 0u)) + __popc(__vcmpeq4(b, 0u)) +
                    __popc(__vcmpeq4(c, 0u)) + __popc(__vcmpeq4(d, 0u))) >> 3;
        }
        for (long long i = (n16 << 4) + (long long)blockIdx.x * BT + threadIdx.x; i < n; i += stride)
            acc += ((int)f[i] == phase) ? 1 : 0;
    }
    acc = warp_sum_ll(acc);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, (unsigned long long)acc);
}

template <typename T>
__global__ void __launch_bounds__(BT)
to_isphase_kernel(const T* __restrict__ in, uint8_t* __restrict__ out, long long n, int phase) {
    const long long stride = (long long)gridDim.x * BT;
    for (long long i = (long long)blockIdx.x * BT + threadIdx.x; i < n; i += stride)
        out[i] = ((int)in[i] == phase) ? 1 : 0;
}

// ------------------------------------------------------------------ K2: union-find CCL
// Read-only root search: safe while other threads hook roots with atomicMin
// (every value ever stored in L[i] is a smaller index of the same final set).
// Parents are always smaller indices of the same set, so a chain is strictly
// decreasing; on the way up every visited node is re-pointed at its grandparent
// (path halving with plain L2 stores, as in ECL-CC): a racing store can only replace
// one valid ancestor by another, never leave the set.  Without it the giant pore
// component of a packing builds parent chains thousands of links long.
__device__ __forceinline__ int uf_find(int* L, int i) {
    int curr = __ldcg(&L[i]);          // L2: other SMs hook roots concurrently
    if (curr != i) {
        int prev = i, next;
        while (curr > (next = __ldcg(&L[curr]))) {
            __stcg(&L[prev], next);
            prev = curr;
            curr = next;
        }
    }
    return curr;
}

// Label-equivalence union (Komura / Playne-Hawick): hook the larger root under
// the smaller one with atomicMin and retry when the root moved meanwhile.
__device__ __forceinline__ void uf_union(int* L, int a, int b) {
    while (true) {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a == b) return;
        if (a > b) { const int t = a; a = b; b = t; }   // a < b: hook b under a
        const int old = atomicMin(&L[b], a);
        if (old == b) return;
        b = old;                                         // b had been hooked to `old`: merge a with it
    }
}

// one warp per x-row: label = linear index of the first cell of the x-run
__global__ void __launch_bounds__(BT)
ccl_rows_kernel(const uint8_t* __restrict__ ph, int* __restrict__ L, int nx, long long nrows) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * BT + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * BT) >> 5;
    for (long long row = warp; row < nrows; row += nwarps) {
        const long long base = row * nx;
        int carry = -1;  // run start (x index) continuing from the previous chunk
        for (int x0 = 0; x0 < nx; x0 += 32) {
            const int x = x0 + lane;
            const bool on = (x < nx) && ph[base + x];
            const unsigned int m = __ballot_sync(0xffffffffu, on);
            const unsigned int below = (lane == 0) ? 0u : (~m & ((1u << lane) - 1u));
            int start;
            if (below == 0u) start = (carry >= 0) ? carry : x0;
            else start = x0 + (32 - __clz(below));
            if (x < nx) L[base + x] = on ? (int)(base + start) : -1;
            const int s31 = __shfl_sync(0xffffffffu, start, 31);
            carry = (m >> 31) ? s31 : -1;
        }
    }
}

// merge runs across -y (AXES & 1) and -z (AXES & 2); one union per start of an overlap segment
template <int AXES>
__global__ void __launch_bounds__(BT)
ccl_merge_kernel(const uint8_t* __restrict__ ph, int* L, int nx, int ny, int nz) {
    const long long n = (long long)nx * ny * nz;
    const long long plane = (long long)nx * ny;
    const long long stride = (long long)gridDim.x * BT;
    for (long long idx = (long long)blockIdx.x * BT + threadIdx.x; idx < n; idx += stride) {
        if (!ph[idx]) continue;
        const int i = (int)(idx % nx);
        const int j = (int)((idx / nx) % ny);
        const int k = (int)(idx / plane);
        const bool left = (i > 0) && ph[idx - 1];
        if ((AXES & 1) && j > 0 && ph[idx - nx]) {
            if (!(left && ph[idx - nx - 1])) uf_union(L, (int)idx, (int)(idx - nx));
        }
        if ((AXES & 2) && k > 0 && ph[idx - plane]) {
            if (!(left && ph[idx - plane - 1])) uf_union(L, (int)idx, (int)(idx - plane));
        }
    }
}

// Merge slices of the box with the slice before them along AXIS (1: row j of every plane with row j-1,
// 2: plane k with plane k-1) for every slice s = span, 3 span, 5 span, ... (< extent): the boundaries between
// groups of `span` slices that earlier launches have already merged internally.  One union per start of an
// overlap segment along x, one thread per cell.
template <int AXIS>
__global__ void __launch_bounds__(BT)
ccl_merge_slices_kernel(const uint8_t* __restrict__ ph, int* L, int nx, int ny, long long plane, long long slice_cells,
                        long long total, int span) {
    const long long t = (long long)blockIdx.x * BT + threadIdx.x;
    if (t >= total) return;
    const long long w = t % slice_cells;
    const int s = span + (int)(t / slice_cells) * 2 * span;
    const int i = (int)(w % nx);
    long long idx, back;
    if (AXIS == 2) { idx = (long long)s * plane + w; back = plane; }
    else { idx = (w / nx) * plane + (long long)s * nx + i; back = nx; }
    if (!ph[idx] || !ph[idx - back]) return;
    if (i > 0 && ph[idx - 1] && ph[idx - back - 1]) return;          // not the start of the overlap segment
    uf_union(L, (int)idx, (int)(idx - back));
}

__global__ void __launch_bounds__(BT)
ccl_flatten_kernel(const uint8_t* __restrict__ ph, int* L, long long n) {
    const long long stride = (long long)gridDim.x * BT;
    for (long long idx = (long long)blockIdx.x * BT + threadIdx.x; idx < n; idx += stride) {
        if (!ph[idx]) continue;
        int r = L[idx];
        while (true) { const int p = L[r]; if (p == r) break; r = p; }
        L[idx] = r;
    }
}

__device__ __forceinline__ void reach_or(unsigned int* reach_words, int root, unsigned int bits) {
    atomicOr(&reach_words[root >> 2], bits << ((root & 3) * 8));
}
__device__ __forceinline__ unsigned int reach_get(const unsigned int* reach_words, int root) {
    return (reach_words[root >> 2] >> ((root & 3) * 8)) & 0xffu;
}

__global__ void __launch_bounds__(BT)
ccl_mark_kernel(const uint8_t* __restrict__ ph, const int* __restrict__ L, unsigned int* reach,
                int nx, int ny, int nz, int dir, int lo_local, int hi_local) {
    // threads enumerate the cells of one plane perpendicular to dir
    const int na = (dir == 0) ? ny : nx;
    const int nb = (dir == 2) ? ny : nz;
    const long long np = (long long)na * nb;
    const long long stride = (long long)gridDim.x * BT;
    const long long plane = (long long)nx * ny;
    for (long long t = (long long)blockIdx.x * BT + threadIdx.x; t < np; t += stride) {
        const int a = (int)(t % na), b = (int)(t / na);
        for (int side = 0; side < 2; ++side) {
            const int p = side ? hi_local : lo_local;
            if (p < 0) continue;
            long long idx;
            if (dir == 0) idx = (long long)b * plane + (long long)a * nx + p;
            else if (dir == 1) idx = (long long)b * plane + (long long)p * nx + a;
            else idx = (long long)p * plane + (long long)b * nx + a;
            if (ph[idx]) reach_or(reach, L[idx], side ? 2u : 1u);
        }
    }
}

__global__ void __launch_bounds__(BT)
ccl_export_kernel(const uint8_t* __restrict__ ph, const int* __restrict__ L,
                  const unsigned int* __restrict__ reach, uint8_t* __restrict__ bits, long long np,
                  long long off) {
    const long long stride = (long long)gridDim.x * BT;
    for (long long t = (long long)blockIdx.x * BT + threadIdx.x; t < np; t += stride)
        bits[t] = ph[off + t] ? (uint8_t)reach_get(reach, L[off + t]) : (uint8_t)0;
}

__global__ void __launch_bounds__(BT)
ccl_import_kernel(const uint8_t* __restrict__ ph, const int* __restrict__ L, unsigned int* reach,
                  const uint8_t* __restrict__ nbr_bits, long long np, long long off, int* changed) {
    const long long stride = (long long)gridDim.x * BT;
    for (long long t = (long long)blockIdx.x * BT + threadIdx.x; t < np; t += stride) {
        if (!ph[off + t]) continue;
        const unsigned int nb = nbr_bits[t] & 3u;
        if (!nb) continue;
        const int root = L[off + t];
        const unsigned int have = reach_get(reach, root);
        if (nb & ~have) { reach_or(reach, root, nb); *changed = 1; }
    }
}

__global__ void __launch_bounds__(BT)
build_active_kernel(const uint8_t* __restrict__ ph, const int* __restrict__ L,
                    const unsigned int* __restrict__ reach, uint8_t* __restrict__ active,
                    long long n, unsigned long long* n_active) {
    const long long stride = (long long)gridDim.x * BT;
    long long acc = 0;
    for (long long idx = (long long)blockIdx.x * BT + threadIdx.x; idx < n; idx += stride) {
        uint8_t a = 0;
        if (ph[idx]) a = (reach_get(reach, L[idx]) == 3u) ? 1 : 0;
        active[idx] = a;
        acc += a;
    }
    acc = warp_sum_ll(acc);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(n_active, (unsigned long long)acc);
}

// ------------------------------------------------------------------ K7
__device__ __forceinline__ int dir_index(int dir, int i, int j, int kglob) {
    return dir == 0 ? i : (dir == 1 ? j : kglob);
}

__global__ void __launch_bounds__(BT)
build_flags_kernel(Grid g, const uint8_t* __restrict__ active, uint8_t* __restrict__ flags, int dir,
                   int n_dir, unsigned long long* counts) {
    const long long n = (long long)g.nz * g.plane;
    const long long stride = (long long)gridDim.x * BT;
    long long c_in = 0, c_out = 0;
    for (long long idx = (long long)blockIdx.x * BT + threadIdx.x; idx < n; idx += stride) {
        uint8_t f = 0;
        if (active[idx]) {
            const int i = (int)(idx % g.nx);
            const int j = (int)((idx / g.nx) % g.ny);
            const int k = (int)(idx / g.plane);
            const int d = dir_index(dir, i, j, g.z0 + k);
            // cell problem (diag_full > 0): no Dirichlet planes, every active cell is a row
            const bool cellp = g.diag_full > 0.0;
            if (!cellp && d == 0) { f = F_DIR; ++c_in; }             // F90:193-197
            else if (!cellp && d == n_dir - 1) { f = F_DIR; ++c_out; }   // F90:198-202
            else {
                f = F_UNK;                                           // F90:126-166 / EffDiffFillMtx.F90:150-220
                const int im = wrap_lo(i, g.nx, g.periodic & PER_X), ip = wrap_hi(i, g.nx, g.periodic & PER_X);
                const int jm = wrap_lo(j, g.ny, g.periodic & PER_Y), jp = wrap_hi(j, g.ny, g.periodic & PER_Y);
                const long long row = idx - i, colk = idx - (long long)j * g.nx;
                if (im >= 0 && active[row + im]) f |= F_XM;
                if (ip >= 0 && active[row + ip]) f |= F_XP;
                if (jm >= 0 && active[colk + (long long)jm * g.nx]) f |= F_YM;
                if (jp >= 0 && active[colk + (long long)jp * g.nx]) f |= F_YP;
                // z neighbours: ghost planes of `active` hold the neighbour slab's
                // plane, or 0 outside the global box
                if (active[idx - g.plane]) f |= F_ZM;
                if (active[idx + g.plane]) f |= F_ZP;
            }
        }
        flags[idx] = f;
    }
    c_in = warp_sum_ll(c_in);
    c_out = warp_sum_ll(c_out);
    if ((threadIdx.x & 31) == 0) {
        if (c_in) atomicAdd(&counts[0], (unsigned long long)c_in);
        if (c_out) atomicAdd(&counts[1], (unsigned long long)c_out);
    }
}

// Same, four x-adjacent cells per thread (nx % 4 == 0): the active bytes of the own quad and of its four y / z
// neighbour quads come as one 32-bit load each, the index split happens once per quad.
__global__ void __launch_bounds__(BT)
build_flags_vec4_kernel(Grid g, const uint8_t* __restrict__ active, uint8_t* __restrict__ flags, int dir,
                        int n_dir, unsigned long long* counts) {
    const long long nq = ((long long)g.nz * g.plane) >> 2;
    const long long stride = (long long)gridDim.x * BT;
    const int qx = g.nx >> 2;                          // quads per row
    const bool cellp = g.diag_full > 0.0;
    long long c_in = 0, c_out = 0;
    for (long long q = (long long)blockIdx.x * BT + threadIdx.x; q < nq; q += stride) {
        const long long idx = q << 2;
        const unsigned int a = *reinterpret_cast<const unsigned int*>(active + idx);
        unsigned int out = 0u;
        if (a) {
            const int i = (int)(q % qx) << 2;
            const long long rowq = q / qx;
            const int j = (int)(rowq % g.ny);
            const int k = (int)(rowq / g.ny);
            const long long row = idx - i, colk = idx - (long long)j * g.nx;
            const int jm = wrap_lo(j, g.ny, g.periodic & PER_Y), jp = wrap_hi(j, g.ny, g.periodic & PER_Y);
            const unsigned int ym = jm >= 0 ? *reinterpret_cast<const unsigned int*>(active + colk + (long long)jm * g.nx) : 0u;
            const unsigned int yp = jp >= 0 ? *reinterpret_cast<const unsigned int*>(active + colk + (long long)jp * g.nx) : 0u;
            const unsigned int zm = *reinterpret_cast<const unsigned int*>(active + idx - g.plane);   // ghost planes are valid
            const unsigned int zp = *reinterpret_cast<const unsigned int*>(active + idx + g.plane);
            const int iw = wrap_lo(i, g.nx, g.periodic & PER_X);
            const int ie = (i + 4 < g.nx) ? i + 4 : ((g.periodic & PER_X) ? 0 : -1);
            const unsigned int west = iw >= 0 ? active[row + iw] : 0u;
            const unsigned int east = ie >= 0 ? active[row + ie] : 0u;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (!((a >> (8 * c)) & 0xffu)) continue;
                const int d = dir == 0 ? i + c : (dir == 1 ? j : g.z0 + k);
                unsigned int f;
                if (!cellp && d == 0) { f = F_DIR; ++c_in; }
                else if (!cellp && d == n_dir - 1) { f = F_DIR; ++c_out; }
                else {
                    f = F_UNK;
                    const unsigned int xm = c > 0 ? (a >> (8 * (c - 1))) & 0xffu : west;
                    const unsigned int xp = c < 3 ? (a >> (8 * (c + 1))) & 0xffu : east;
                    if (xm) f |= F_XM;
                    if (xp) f |= F_XP;
                    if ((ym >> (8 * c)) & 0xffu) f |= F_YM;
                    if ((yp >> (8 * c)) & 0xffu) f |= F_YP;
                    if ((zm >> (8 * c)) & 0xffu) f |= F_ZM;
                    if ((zp >> (8 * c)) & 0xffu) f |= F_ZP;
                }
                out |= f << (8 * c);
            }
        }
        *reinterpret_cast<unsigned int*>(flags + idx) = out;
    }
    c_in = warp_sum_ll(c_in);
    c_out = warp_sum_ll(c_out);
    if ((threadIdx.x & 31) == 0) {
        if (c_in) atomicAdd(&counts[0], (unsigned long long)c_in);
        if (c_out) atomicAdd(&counts[1], (unsigned long long)c_out);
    }
}

__global__ void __launch_bounds__(BT)
initial_guess_kernel(Grid g, const uint8_t* __restrict__ flags, double* __restrict__ x, int dir,
                     int n_dir, double vlo, double vhi, int mirror_quirk) {
    const long long n = (long long)g.nz * g.plane;
    const long long stride = (long long)gridDim.x * BT;
    const double ext = (double)(n_dir - 1);
    const double factor = (fabs(ext) < 1e-15) ? 0.0 : 1.0 / ext;     // F90:236-241
    for (long long idx = (long long)blockIdx.x * BT + threadIdx.x; idx < n; idx += stride) {
        const uint8_t f = flags[idx];
        double v = 0.0;
        if (f & (F_UNK | F_DIR)) {
            const int i = (int)(idx % g.nx);
            const int j = (int)((idx / g.nx) % g.ny);
            const int k = (int)(idx / g.plane);
            const int d = dir_index(dir, i, j, g.z0 + k);
            if (f & F_DIR) {
                v = mirror_quirk ? vlo + (vhi - vlo) * (double)d * factor : ((d == 0) ? vlo : vhi);
            } else {
                v = vlo + (vhi - vlo) * (double)d * factor;          // F90:242-258
                if (mirror_quirk) {
                    // F90:233: cells whose diagonal is exactly 1 keep the
                    // caller's value-initialised 0 (TortuosityHypre.cpp:606)
                    const double dg = g.cx * (double)__popc(f & 0x03u) +
                                      g.cy * (double)__popc(f & 0x0cu) +
                                      g.cz * (double)__popc(f & 0x30u);
                    if (!(fabs(dg - 1.0) > 1e-15)) v = 0.0;
                }
            }
        }
        x[idx] = v;
    }
}

// ------------------------------------------------------------------ K8
__global__ void __launch_bounds__(BT)
flux_kernel(Grid g, const uint8_t* __restrict__ flags, const double* __restrict__ x, int dir,
            int n_dir, double inv_dx, double* partials, unsigned int* counter, double* out) {
    const int na = (dir == 0) ? g.ny : g.nx;
    const int nb = (dir == 2) ? g.ny : g.nz;
    const long long np = (long long)na * nb;
    const long long stride = (long long)gridDim.x * BT;
    const long long sd = (dir == 0) ? 1 : (dir == 1 ? (long long)g.nx : g.plane);
    // local position of the global planes 0 and n_dir-1 (z: only on the owning slab)
    const int lo = (dir == 2) ? (0 - g.z0) : 0;
    const int hi = (dir == 2) ? (n_dir - 1 - g.z0) : (n_dir - 1);
    const int nloc = (dir == 2) ? g.nz : n_dir;
    double fin = 0.0, fout = 0.0;
    for (long long t = (long long)blockIdx.x * BT + threadIdx.x; t < np; t += stride) {
        const int a = (int)(t % na), b = (int)(t / na);
        long long base;   // index of the cell with dir-index 0 (local)
        if (dir == 0) base = (long long)b * g.plane + (long long)a * g.nx;
        else if (dir == 1) base = (long long)b * g.plane + a;
        else base = (long long)b * g.nx + a;
        if (n_dir >= 2) {
            if (lo >= 0 && lo < nloc) {
                const long long c = base + lo * sd, in = c + sd;     // :1067-1083
                if ((flags[c] & F_DIR) && (flags[in] & (F_UNK | F_DIR)))
                    fin += -((x[in] - x[c]) * inv_dx);
            }
            if (hi >= 0 && hi < nloc) {
                const long long c = base + hi * sd, in = c - sd;     // :1086-1103
                if ((flags[c] & F_DIR) && (flags[in] & (F_UNK | F_DIR)))
                    fout += -((x[c] - x[in]) * inv_dx);
            }
        }
    }
    double v[2] = {fin, fout};
    grid_reduce<2>(v, partials, counter, out);
}

// ------------------------------------------------------------------ rows / checks
// rhs of the cell problem for chi_dir at an active cell (EffDiffFillMtx.F90:150-234,
// same operation order): -(D_p - D_m)/(2 dx) plus +-1/dx for each face of that
// direction that looks at the solid.
__device__ __forceinline__ double cellp_rhs_at(const Grid& g, uint8_t f, int dir) {
    if (!(f & F_UNK)) return 0.0;
    const double h = dir == 0 ? g.hx : (dir == 1 ? g.hy : g.hz);
    const uint8_t bm = dir == 0 ? F_XM : (dir == 1 ? F_YM : F_ZM);
    const uint8_t bp = dir == 0 ? F_XP : (dir == 1 ? F_YP : F_ZP);
    const double dm = (f & bm) ? 1.0 : 0.0, dp = (f & bp) ? 1.0 : 0.0;
    double flux = 0.0;
    if (!(f & bm)) flux = flux + (1.0 / h);
    if (!(f & bp)) flux = flux - (1.0 / h);
    const double div = -(dp - dm) * (1.0 / (2.0 * h));
    return div + flux;
}

__device__ __forceinline__ void make_row(const Grid& g, uint8_t f, int d, int n_dir, double vlo,
                                         double vhi, double (&a)[7], double& rhs, int dir = 0) {
#pragma unroll
    for (int s = 0; s < 7; ++s) a[s] = 0.0;
    rhs = 0.0;
    if (g.diag_full > 0.0) {                         // EffDiffFillMtx.F90:124-236
        if (f & F_UNK) {
            if (f & F_XM) a[1] = -g.cx;
            if (f & F_XP) a[2] = -g.cx;
            if (f & F_YM) a[3] = -g.cy;
            if (f & F_YP) a[4] = -g.cy;
            if (f & F_ZM) a[5] = -g.cz;
            if (f & F_ZP) a[6] = -g.cz;
            a[0] = g.cx + g.cx + g.cy + g.cy + g.cz + g.cz;
            rhs = cellp_rhs_at(g, f, dir);
        } else {
            a[0] = 1.0;
        }
        return;
    }
    if (f & F_UNK) {
        double dg = 0.0;
        if (f & F_XM) { a[1] = -g.cx; dg += g.cx; }
        if (f & F_XP) { a[2] = -g.cx; dg += g.cx; }
        if (f & F_YM) { a[3] = -g.cy; dg += g.cy; }
        if (f & F_YP) { a[4] = -g.cy; dg += g.cy; }
        if (f & F_ZM) { a[5] = -g.cz; dg += g.cz; }
        if (f & F_ZP) { a[6] = -g.cz; dg += g.cz; }
        a[0] = dg;
    } else {
        a[0] = 1.0;
        if (f & F_DIR) rhs = (d == 0) ? vlo : vhi;
    }
}

__global__ void __launch_bounds__(BT)
export_rows_kernel(Grid g, const uint8_t* __restrict__ flags, int dir, int n_dir, double vlo,
                   double vhi, double* __restrict__ a7, double* __restrict__ rhs) {
    const long long n = (long long)g.nz * g.plane;
    const long long stride = (long long)gridDim.x * BT;
    for (long long idx = (long long)blockIdx.x * BT + threadIdx.x; idx < n; idx += stride) {
        const int i = (int)(idx % g.nx);
        const int j = (int)((idx / g.nx) % g.ny);
        const int k = (int)(idx / g.plane);
        double a[7], r;
        make_row(g, flags[idx], dir_index(dir, i, j, g.z0 + k), n_dir, vlo, vhi, a, r, dir);
        if (a7) {
#pragma unroll
            for (int s = 0; s < 7; ++s) a7[idx * 7 + s] = a[s];
        }
        if (rhs) rhs[idx] = r;
    }
}

__global__ void __launch_bounds__(BT)
check_rows_kernel(Grid g, const uint8_t* __restrict__ flags, const uint8_t* __restrict__ active,
                  int dir, int n_dir, unsigned long long* bad) {
    const long long n = (long long)g.nz * g.plane;
    const long long stride = (long long)gridDim.x * BT;
    const double tol = 1e-14;                                         // TortuosityHypre.cpp:903
    long long nbad = 0;
    for (long long idx = (long long)blockIdx.x * BT + threadIdx.x; idx < n; idx += stride) {
        const int i = (int)(idx % g.nx);
        const int j = (int)((idx / g.nx) % g.ny);
        const int k = (int)(idx / g.plane);
        const int d = dir_index(dir, i, j, g.z0 + k);
        double a[7], r;
        make_row(g, flags[idx], d, n_dir, 0.25, 0.75, a, r, dir);
        bool ok = true;
        for (int s = 0; s < 7; ++s) ok = ok && isfinite(a[s]);
        const bool act = active[idx] != 0;
        const bool cellp = g.diag_full > 0.0;
        const bool dirichlet = !cellp && act && (d == 0 || d == n_dir - 1);     // :950-956
        double off = 0.0;
        for (int s = 1; s < 7; ++s) off = fmax(off, fabs(a[s]));
        if (!act) {                                                   // :957-960
            ok = ok && fabs(a[0] - 1.0) <= tol && fabs(r) <= tol && off <= tol;
        } else if (dirichlet) {                                       // :961-965
            const double e = (d == 0) ? 0.25 : 0.75;
            ok = ok && fabs(a[0] - 1.0) <= tol && fabs(r - e) <= tol && off <= tol;
        } else {                                                      // :966-972
            double row = 0.0;
            for (int s = 0; s < 7; ++s) row += a[s];
            // tortuosity rows sum to zero; cell-problem rows are weakly diagonally dominant
            if (!cellp) ok = ok && (a[0] > tol) && fabs(r) <= tol && fabs(row) <= tol;
            else ok = ok && (a[0] > tol) && row >= -tol && isfinite(r);
            // and the stored couplings must mirror the active neighbours
            uint8_t e = F_UNK;
            const int im = wrap_lo(i, g.nx, g.periodic & PER_X), ip = wrap_hi(i, g.nx, g.periodic & PER_X);
            const int jm = wrap_lo(j, g.ny, g.periodic & PER_Y), jp = wrap_hi(j, g.ny, g.periodic & PER_Y);
            const long long rowi = idx - i, colk = idx - (long long)j * g.nx;
            if (im >= 0 && active[rowi + im]) e |= F_XM;
            if (ip >= 0 && active[rowi + ip]) e |= F_XP;
            if (jm >= 0 && active[colk + (long long)jm * g.nx]) e |= F_YM;
            if (jp >= 0 && active[colk + (long long)jp * g.nx]) e |= F_YP;
            if (active[idx - g.plane]) e |= F_ZM;
            if (active[idx + g.plane]) e |= F_ZP;
            ok = ok && (e == flags[idx]);
        }
        if (!ok) ++nbad;
    }
    nbad = warp_sum_ll(nbad);
    if ((threadIdx.x & 31) == 0 && nbad) atomicAdd(bad, (unsigned long long)nbad);
}

// ------------------------------------------------------------------ cell problem (homogenisation)
// r += sign * b with b the rhs of the chi_dir cell problem (r may be null), and
// out[0] = sum b^2 (HYPRE's stop rule is relative to ||b||_2).
__global__ void __launch_bounds__(BT)
cellp_rhs_kernel(Grid g, const uint8_t* __restrict__ flags, double* __restrict__ r, int dir, double sign,
                 double* partials, unsigned int* counter, double* out) {
    const long long n = (long long)g.nz * g.plane;
    const long long stride = (long long)gridDim.x * BT;
    double acc = 0.0;
    for (long long idx = (long long)blockIdx.x * BT + threadIdx.x; idx < n; idx += stride) {
        const double b = cellp_rhs_at(g, flags[idx], dir);
        if (b != 0.0) {
            if (r) r[idx] += sign * b;
            acc += b * b;
        }
    }
    double v[1] = {acc};
    grid_reduce<1>(v, partials, counter, out);
}

// out[a] = sum over active cells of the central difference of chi along axis a
// (calculate_Deff_tensor_homogenization, src/props/Diffusion.cpp:115-131); chi is zero
// in the solid and the box is periodic (z through the ghost planes).
__global__ void __launch_bounds__(BT)
cellp_grad_sums_kernel(Grid g, const uint8_t* __restrict__ flags, const double* __restrict__ x,
                       double* partials, unsigned int* counter, double* out) {
    const long long n = (long long)g.nz * g.plane;
    const long long stride = (long long)gridDim.x * BT;
    const double i2x = 1.0 / (2.0 * g.hx), i2y = 1.0 / (2.0 * g.hy), i2z = 1.0 / (2.0 * g.hz);
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (long long idx = (long long)blockIdx.x * BT + threadIdx.x; idx < n; idx += stride) {
        if (!(flags[idx] & F_UNK)) continue;
        const int i = (int)(idx % g.nx);
        const int j = (int)((idx / g.nx) % g.ny);
        const int im = wrap_lo(i, g.nx, true), ip = wrap_hi(i, g.nx, true);
        const int jm = wrap_lo(j, g.ny, true), jp = wrap_hi(j, g.ny, true);
        const long long row = idx - i, colk = idx - (long long)j * g.nx;
        sx += (x[row + ip] - x[row + im]) * i2x;
        sy += (x[colk + (long long)jp * g.nx] - x[colk + (long long)jm * g.nx]) * i2y;
        sz += (x[idx + g.plane] - x[idx - g.plane]) * i2z;
    }
    double v[3] = {sx, sy, sz};
    grid_reduce<3>(v, partials, counter, out);
}

// out[0] = unknowns, out[1] = aligned 2-cell groups holding an unknown, out[2] = aligned
// 4-cell groups holding one (the 16-byte granules the fp64 / fp32 kernels skip when empty)
__global__ void __launch_bounds__(BT)
flag_stats_kernel(const uint8_t* __restrict__ flags, long long n4, unsigned long long* out) {
    const long long stride = (long long)gridDim.x * BT;
    long long unk = 0, pairs = 0, quads = 0;
    for (long long t = (long long)blockIdx.x * BT + threadIdx.x; t < n4; t += stride) {
        const unsigned int f = *reinterpret_cast<const unsigned int*>(flags + 4 * t) & 0x40404040u;
        unk += __popc(f);
        pairs += ((f & 0x00004040u) ? 1 : 0) + ((f & 0x40400000u) ? 1 : 0);
        quads += f ? 1 : 0;
    }
    unk = warp_sum_ll(unk); pairs = warp_sum_ll(pairs); quads = warp_sum_ll(quads);
    if ((threadIdx.x & 31) == 0) {
        if (unk) atomicAdd(&out[0], (unsigned long long)unk);
        if (pairs) atomicAdd(&out[1], (unsigned long long)pairs);
        if (quads) atomicAdd(&out[2], (unsigned long long)quads);
    }
}

}  // namespace

void flag_stats(const uint8_t* flags, long long n, unsigned long long* out, cudaStream_t st) {
    flag_stats_kernel<<<nblocks(n / 4, 148, 4), BT, 0, st>>>(flags, n / 4, out);
}

void cellp_rhs(const Grid& g, const uint8_t* flags, double* r, int dir, double sign, double* partials,
               unsigned int* counter, double* out, cudaStream_t st) {
    const long long n = (long long)g.nz * g.plane;
    cellp_rhs_kernel<<<nblocks(n, 148, 4), BT, 0, st>>>(g, flags, r, dir, sign, partials, counter, out);
}
void cellp_grad_sums(const Grid& g, const uint8_t* flags, const double* x, double* partials,
                     unsigned int* counter, double* out, cudaStream_t st) {
    const long long n = (long long)g.nz * g.plane;
    cellp_grad_sums_kernel<<<nblocks(n, 148, 4), BT, 0, st>>>(g, flags, x, partials, counter, out);
}

// ------------------------------------------------------------------ launchers
void count_phase_u8(const uint8_t* f, long long n, int phase, unsigned long long* out, int n_sm, cudaStream_t st) {
    if ((reinterpret_cast<uintptr_t>(f) & 15) == 0)
        count_u8_vec_kernel<<<nblocks((n + 15) / 16, n_sm), BT, 0, st>>>(f, n, phase, out);
    else
        count_kernel<uint8_t><<<nblocks(n, n_sm), BT, 0, st>>>(f, n, phase, out);
}
void count_phase_i32(const int32_t* f, long long n, int phase, unsigned long long* out, int n_sm, cudaStream_t st) {
    count_kernel<int32_t><<<nblocks(n, n_sm), BT, 0, st>>>(f, n, phase, out);
}
void phase_i32_to_u8(const int32_t* in, uint8_t* o, long long n, int phase, int n_sm, cudaStream_t st) {
    to_isphase_kernel<int32_t><<<nblocks(n, n_sm), BT, 0, st>>>(in, o, n, phase);
}
void phase_u8_to_isphase(const uint8_t* in, uint8_t* o, long long n, int phase, int n_sm, cudaStream_t st) {
    to_isphase_kernel<uint8_t><<<nblocks(n, n_sm), BT, 0, st>>>(in, o, n, phase);
}

int ccl_label(const uint8_t* ph, int* L, int nx, int ny, int nz, int n_sm, cudaStream_t st) {
    const long long n = (long long)nx * ny * nz;
    const long long nrows = (long long)ny * nz;
    ccl_rows_kernel<<<nblocks(nrows * 32, n_sm), BT, 0, st>>>(ph, L, nx, nrows);
    // OI_CCL=0: every union (across y and across z) in one pass over the box.  At 1024^3 that pass costs
    // ~80 ms (ncu: profiles/r2_launches_1024.csv) although it moves little data: every plane hooks onto
    // every other at once, and finds walk parent chains that cross many planes before path halving has
    // shortened them (the same happens inside a plane: one pass over all rows of all planes took 52 ms).
    // Default: divide and conquer -- first the boundaries between single rows (rows 1, 3, 5, ... of every plane
    // onto the row before), then between pairs of rows (2, 6, 10, ...), then groups of four, ... and the same along
    // z -- so that both sides of a boundary are already resolved when it is merged and a find is two or three
    // hops.  log2(ny) + log2(nz) launches, one thread per cell of the merged slices.  (Round 2 first tried one
    // launch per slice in stream order: 46 ms at 1024^3, launch bound at 20 us per slice.)
    // Small boxes keep the single pass (its chains are short there, and a few hundred launches would cost more
    // than they save); OI_CCL=1 forces the slices, OI_CCL=0 the single pass.
    const char* e = getenv("OI_CCL");
    const bool slices = e ? (e[0] != '0') : (n >= (1LL << 25));
    if (!slices) {
        ccl_merge_kernel<3><<<nblocks(n, n_sm), BT, 0, st>>>(ph, L, nx, ny, nz);
        ccl_flatten_kernel<<<nblocks(n, n_sm), BT, 0, st>>>(ph, L, n);
        return 3;
    }
    const long long plane = (long long)nx * ny;
    int launches = 2;
    // rows first (trees stay inside one plane), then planes; span = 1, 2, 4, ...: at every launch the groups of
    // `span` slices on either side of a merged boundary are already resolved, so parent chains grow by one hop per
    // launch at most and path halving keeps a find at two or three hops
    const long long rowcells = (long long)nx * nz;
    for (int span = 1; span < ny; span *= 2, ++launches) {
        const long long nb = (ny - span + 2 * span - 1) / (2 * span);           // slices span, 3 span, ... < ny
        const long long total = nb * rowcells;
        ccl_merge_slices_kernel<1><<<(unsigned)((total + BT - 1) / BT), BT, 0, st>>>(ph, L, nx, ny, plane, rowcells, total, span);
    }
    for (int span = 1; span < nz; span *= 2, ++launches) {
        const long long nb = (nz - span + 2 * span - 1) / (2 * span);
        const long long total = nb * plane;
        ccl_merge_slices_kernel<2><<<(unsigned)((total + BT - 1) / BT), BT, 0, st>>>(ph, L, nx, ny, plane, plane, total, span);
    }
    ccl_flatten_kernel<<<nblocks(n, n_sm), BT, 0, st>>>(ph, L, n);
    return launches;
}
void ccl_mark_planes(const uint8_t* ph, const int* L, unsigned int* reach, int nx, int ny, int nz,
                     int dir, int lo_local, int hi_local, int n_sm, cudaStream_t st) {
    const long long np = (long long)((dir == 0) ? ny : nx) * ((dir == 2) ? ny : nz);
    ccl_mark_kernel<<<nblocks(np, n_sm), BT, 0, st>>>(ph, L, reach, nx, ny, nz, dir, lo_local, hi_local);
}
void ccl_export_plane(const uint8_t* ph, const int* L, const unsigned int* reach, uint8_t* bits,
                      int nx, int ny, int k, int n_sm, cudaStream_t st) {
    const long long np = (long long)nx * ny;
    ccl_export_kernel<<<nblocks(np, n_sm), BT, 0, st>>>(ph, L, reach, bits, np, (long long)k * np);
}
void ccl_import_plane(const uint8_t* ph, const int* L, unsigned int* reach, const uint8_t* nbr,
                      int nx, int ny, int k, int* changed, int n_sm, cudaStream_t st) {
    const long long np = (long long)nx * ny;
    ccl_import_kernel<<<nblocks(np, n_sm), BT, 0, st>>>(ph, L, reach, nbr, np, (long long)k * np, changed);
}
void build_active(const uint8_t* ph, const int* L, const unsigned int* reach, uint8_t* active,
                  int nx, int ny, int nz, unsigned long long* n_active, int n_sm, cudaStream_t st) {
    const long long n = (long long)nx * ny * nz;
    build_active_kernel<<<nblocks(n, n_sm), BT, 0, st>>>(ph, L, reach, active, n, n_active);
}
void build_flags(const Grid& g, const uint8_t* active, uint8_t* flags, int dir,
                 unsigned long long* counts, cudaStream_t st) {
    const long long n = (long long)g.nz * g.plane;
    const int n_dir = (dir == 0) ? g.nx : (dir == 1 ? g.ny : g.nzg);
    // (plane 0 of every field is 256-byte aligned, so with nx % 4 == 0 every quad is 4-byte aligned)
    if ((g.nx & 3) == 0 && (reinterpret_cast<uintptr_t>(active) & 3) == 0 && (reinterpret_cast<uintptr_t>(flags) & 3) == 0)
        build_flags_vec4_kernel<<<nblocks(n >> 2, 148), BT, 0, st>>>(g, active, flags, dir, n_dir, counts);
    else
        build_flags_kernel<<<nblocks(n, 148), BT, 0, st>>>(g, active, flags, dir, n_dir, counts);
}
void fill_initial_guess(const Grid& g, const uint8_t* flags, double* x, int dir, int n_dir,
                        double vlo, double vhi, int mirror_quirk, cudaStream_t st) {
    const long long n = (long long)g.nz * g.plane;
    initial_guess_kernel<<<nblocks(n, 148), BT, 0, st>>>(g, flags, x, dir, n_dir, vlo, vhi, mirror_quirk);
}
void flux_planes(const Grid& g, const uint8_t* flags, const double* x, int dir, int n_dir,
                 double* partials, unsigned int* counter, double* out, cudaStream_t st) {
    const long long np = (long long)((dir == 0) ? g.ny : g.nx) * ((dir == 2) ? g.ny : g.nz);
    const double dxd = 1.0 / sqrt(dir == 0 ? g.cx : (dir == 1 ? g.cy : g.cz));
    flux_kernel<<<nblocks(np, 148, 4), BT, 0, st>>>(g, flags, x, dir, n_dir, 1.0 / dxd, partials, counter, out);
}
void export_rows(const Grid& g, const uint8_t* flags, const uint8_t*, int dir, int n_dir,
                 double vlo, double vhi, double* a7, double* rhs, cudaStream_t st) {
    const long long n = (long long)g.nz * g.plane;
    export_rows_kernel<<<nblocks(n, 148), BT, 0, st>>>(g, flags, dir, n_dir, vlo, vhi, a7, rhs);
}
void check_rows(const Grid& g, const uint8_t* flags, const uint8_t* active, int dir, int n_dir,
                unsigned long long* bad, cudaStream_t st) {
    const long long n = (long long)g.nz * g.plane;
    check_rows_kernel<<<nblocks(n, 148), BT, 0, st>>>(g, flags, active, dir, n_dir, bad);
}
namespace {
// tortuosity_remspot (src/props/Tortuosity_filcc.F90:88-177) is an IN-PLACE sweep in
// i-fastest order: a cell sees the already-updated values of its -x,-y,-z neighbours
// and the old values of its +x,+y,+z neighbours.  That is a triangular dependency,
// solved here by fixed-point iteration on the flip flags (exact after as many rounds
// as the longest chain of mutually dependent isolated voxels, usually 2-3):
//   flip_c = no in-domain neighbour equals v_c, with earlier neighbours read as
//            v ^ flip (current estimate) and later neighbours as v.
__global__ void __launch_bounds__(BT)
remspot_round_kernel(const uint8_t* __restrict__ v, const uint8_t* __restrict__ fcur,
                     uint8_t* __restrict__ fnext, int nx, int ny, int nz, int* changed) {
    const long long n = (long long)nx * ny * nz;
    const long long plane = (long long)nx * ny;
    const long long stride = (long long)gridDim.x * BT;
    bool any = false;
    for (long long idx = (long long)blockIdx.x * BT + threadIdx.x; idx < n; idx += stride) {
        const int i = (int)(idx % nx);
        const int j = (int)((idx / nx) % ny);
        const int k = (int)(idx / plane);
        const uint8_t c = v[idx];
        bool connected = false;                                  // outside the domain never matches
        if (i + 1 < nx) connected |= (v[idx + 1] == c);
        if (j + 1 < ny) connected |= (v[idx + nx] == c);
        if (k + 1 < nz) connected |= (v[idx + plane] == c);
        if (i > 0) connected |= ((uint8_t)(v[idx - 1] ^ fcur[idx - 1]) == c);
        if (j > 0) connected |= ((uint8_t)(v[idx - nx] ^ fcur[idx - nx]) == c);
        if (k > 0) connected |= ((uint8_t)(v[idx - plane] ^ fcur[idx - plane]) == c);
        const uint8_t f = connected ? 0 : 1;
        if (f != fcur[idx]) any = true;
        fnext[idx] = f;
    }
    if (__any_sync(0xffffffffu, any) && (threadIdx.x & 31) == 0) *changed = 1;
}

__global__ void __launch_bounds__(BT)
remspot_apply_kernel(uint8_t* __restrict__ v, const uint8_t* __restrict__ f, long long n,
                     unsigned long long* flips) {
    const long long stride = (long long)gridDim.x * BT;
    long long acc = 0;
    for (long long idx = (long long)blockIdx.x * BT + threadIdx.x; idx < n; idx += stride) {
        const uint8_t fl = f[idx];
        v[idx] ^= fl;
        acc += fl;
    }
    acc = warp_sum_ll(acc);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(flips, (unsigned long long)acc);
}

template <typename T>
__global__ void __launch_bounds__(BT)
count_nonbinary_kernel(const T* __restrict__ f, long long n, unsigned long long* out) {
    const long long stride = (long long)gridDim.x * BT;
    long long acc = 0;
    for (long long i = (long long)blockIdx.x * BT + threadIdx.x; i < n; i += stride) {
        const int v = (int)f[i];
        acc += (v != 0 && v != 1) ? 1 : 0;
    }
    acc = warp_sum_ll(acc);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, (unsigned long long)acc);
}
}  // namespace

void remspot_round(const uint8_t* v, const uint8_t* fcur, uint8_t* fnext, int nx, int ny, int nz,
                   int* changed, int n_sm, cudaStream_t st) {
    const long long n = (long long)nx * ny * nz;
    remspot_round_kernel<<<nblocks(n, n_sm), BT, 0, st>>>(v, fcur, fnext, nx, ny, nz, changed);
}
void remspot_apply(uint8_t* v, const uint8_t* f, long long n, unsigned long long* flips, int n_sm,
                   cudaStream_t st) {
    remspot_apply_kernel<<<nblocks(n, n_sm), BT, 0, st>>>(v, f, n, flips);
}
void count_nonbinary_u8(const uint8_t* f, long long n, unsigned long long* out, int n_sm, cudaStream_t st) {
    count_nonbinary_kernel<uint8_t><<<nblocks(n, n_sm), BT, 0, st>>>(f, n, out);
}
void count_nonbinary_i32(const int32_t* f, long long n, unsigned long long* out, int n_sm, cudaStream_t st) {
    count_nonbinary_kernel<int32_t><<<nblocks(n, n_sm), BT, 0, st>>>(f, n, out);
}

}  // namespace oi
