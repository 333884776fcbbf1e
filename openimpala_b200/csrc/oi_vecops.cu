// Krylov vector kernels (K4): fused AXPY pairs + dot products with warp-shuffle
// and block reductions, deterministic last-block finish.  These replace HYPRE's
// struct axpy / inner-product calls inside the Krylov loop (reference call site
// src/props/TortuosityHypre.cpp:681-683).  Scalars (alpha, beta) are read from
// device memory so the loop needs no host round trip to form them.
#include <cstdlib>

#include "oi_kernels.h"

namespace oi {

namespace {
constexpr int VT = 256;

__global__ void __launch_bounds__(VT)
axpy2_dot_kernel(long long n, double* __restrict__ x, double* __restrict__ r,
                 const double* __restrict__ p, const double* __restrict__ q,
                 const double* __restrict__ num, const double* __restrict__ den,
                 double* partials, unsigned int* counter, double* out) {
    const double a = num[0] / den[0];
    const long long stride = (long long)gridDim.x * VT * 2;
    double acc = 0.0;
    // two doubles per thread per step (16-byte vector access when aligned)
    const long long n2 = n & ~1LL;
    for (long long i = ((long long)blockIdx.x * VT + threadIdx.x) * 2; i < n2; i += stride) {
        double2 xv = *reinterpret_cast<double2*>(x + i);
        double2 rv = *reinterpret_cast<double2*>(r + i);
        const double2 pv = *reinterpret_cast<const double2*>(p + i);
        const double2 qv = *reinterpret_cast<const double2*>(q + i);
        xv.x += a * pv.x; xv.y += a * pv.y;
        rv.x -= a * qv.x; rv.y -= a * qv.y;
        *reinterpret_cast<double2*>(x + i) = xv;
        *reinterpret_cast<double2*>(r + i) = rv;
        acc += rv.x * rv.x + rv.y * rv.y;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && n2 < n) {
        const long long i = n2;
        x[i] += a * p[i];
        const double rv = r[i] - a * q[i];
        r[i] = rv;
        acc += rv * rv;
    }
    double v[1] = {acc};
    grid_reduce<1>(v, partials, counter, out);
}

// Same update, plus what the next preconditioner application needs: the residual in
// multigrid precision (r32) and its first smoothing sweep from a zero guess,
// z1 = w0 * r_new / diag.  Saves re-reading r and one launch per Krylov iteration.
// UPDATE_X = false: x += alpha p is left to the next xpby (which has p in hand anyway), so this
// kernel touches neither x nor p.
// Every vector here is zero on cells that are not unknowns (solid, non-percolating,
// Dirichlet planes: p = q = r = 0 there and x keeps its value), so a 16-byte pair
// without an unknown is skipped: its sectors are never fetched or written.  NC pairs
// per trip, all their loads issued before the first store (the compiler cannot
// reorder a load of x[i2] over a store to x[i]), and the flags of the next trip are
// fetched a trip ahead, so the flag -> data dependency is off the critical path.
// Works on the pairs [lo, hi) with `nthr` threads striding together, this one being thread `t`;
// PUSH: the range is a boundary plane and z1 also goes to dst (a neighbour's ghost plane).
// Returns this thread's share of r.r.
template <bool UPDATE_X, bool PUSH, int NC>
__device__ __forceinline__ double axpy2_first_range(long long lo, long long hi, long long t, long long nthr,
                                                    const uint8_t* __restrict__ flags, double* __restrict__ x,
                                                    double* __restrict__ r, const double* __restrict__ p,
                                                    const double* __restrict__ q, mg_t* __restrict__ r32,
                                                    mg_t* __restrict__ z1, double a, const double* winv,
                                                    mg_t* __restrict__ dst) {
    const long long stride = nthr * 2;
    double acc = 0.0;
    typedef typename Vec2<mg_t>::type mg2;
    constexpr unsigned int UNK2 = (unsigned int)F_UNK | ((unsigned int)F_UNK << 8);
    const long long i0 = lo + t * 2;
    unsigned int fnext[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const long long ii = i0 + c * stride;
        fnext[c] = (ii < hi) ? *reinterpret_cast<const unsigned short*>(flags + ii) : 0u;
    }
    for (long long i = i0; i < hi; i += NC * stride) {
        unsigned int f[NC];
        double2 xv[NC], rv[NC], pv[NC], qv[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            f[c] = fnext[c];
            const long long in = i + (NC + c) * stride;
            fnext[c] = (in < hi) ? *reinterpret_cast<const unsigned short*>(flags + in) : 0u;
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            if (f[c] & UNK2) {
                const long long ii = i + c * stride;
                if (UPDATE_X) {
                    xv[c] = *reinterpret_cast<const double2*>(x + ii);
                    pv[c] = *reinterpret_cast<const double2*>(p + ii);
                }
                rv[c] = *reinterpret_cast<const double2*>(r + ii);
                qv[c] = *reinterpret_cast<const double2*>(q + ii);
            }
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            if (f[c] & UNK2) {
                const long long ii = i + c * stride;
                if (UPDATE_X) {
                    xv[c].x += a * pv[c].x; xv[c].y += a * pv[c].y;
                    *reinterpret_cast<double2*>(x + ii) = xv[c];
                }
                rv[c].x -= a * qv[c].x; rv[c].y -= a * qv[c].y;
                *reinterpret_cast<double2*>(r + ii) = rv[c];
                acc += rv[c].x * rv[c].x + rv[c].y * rv[c].y;
                const unsigned int f0 = f[c] & 0xffu, f1 = f[c] >> 8;
                mg2 rr, zv;
                rr.x = (mg_t)rv[c].x; rr.y = (mg_t)rv[c].y;
                zv.x = (f0 & F_UNK) ? (mg_t)(rv[c].x * winv[f0 & 63u]) : (mg_t)0;
                zv.y = (f1 & F_UNK) ? (mg_t)(rv[c].y * winv[f1 & 63u]) : (mg_t)0;
                *reinterpret_cast<mg2*>(r32 + ii) = rr;
                *reinterpret_cast<mg2*>(z1 + ii) = zv;
                if (PUSH) *reinterpret_cast<mg2*>(dst + (ii - lo)) = zv;
            }
        }
    }
    return acc;
}

// HALO (z-slabs): the first `halo_blocks` blocks own the two boundary planes and store z1 into the
// neighbours' ghost planes as well; the last block to finish publishes the exchange (needs an even
// plane size and at least two planes).
template <bool UPDATE_X, bool HALO, int NCI>
__global__ void __launch_bounds__(VT, NCI == 4 ? 0 : 4)
axpy2_dot_first_kernel(Grid g, const uint8_t* __restrict__ flags, long long n, double* __restrict__ x,
                       double* __restrict__ r, const double* __restrict__ p, const double* __restrict__ q,
                       mg_t* __restrict__ r32, mg_t* __restrict__ z1, const double* __restrict__ num,
                       const double* __restrict__ den, double w0, double* partials, unsigned int* counter,
                       double* out, int halo_blocks, HaloOut ho) {
    __shared__ double winv[64];
    if (threadIdx.x < 64) {
        const int t = threadIdx.x;
        const double d = row_diag<double>((unsigned int)t, g);
        winv[t] = d > 0.0 ? w0 / d : 0.0;
    }
    __syncthreads();
    const double a = num[0] / den[0];
    const long long n2 = n & ~1LL;
    double acc = 0.0;
    if (!HALO) {
        acc = axpy2_first_range<UPDATE_X, false, NCI>(0, n2, (long long)blockIdx.x * VT + threadIdx.x,
                                                    (long long)gridDim.x * VT, flags, x, r, p, q, r32, z1, a, winv, nullptr);
        if (blockIdx.x == 0 && threadIdx.x == 0 && n2 < n) {
            const long long i = n2;
            if (UPDATE_X) x[i] += a * p[i];
            const double rv = r[i] - a * q[i];
            r[i] = rv;
            acc += rv * rv;
            const unsigned int f = flags[i];
            r32[i] = (mg_t)rv;
            z1[i] = (f & F_UNK) ? (mg_t)(rv * winv[f & 63u]) : (mg_t)0;
        }
    } else {
        const long long plane = g.plane, top = n - plane;
        if ((int)blockIdx.x < halo_blocks) {
            const long long t = (long long)blockIdx.x * VT + threadIdx.x, nthr = (long long)halo_blocks * VT;
            if (ho.dst_lo) acc += axpy2_first_range<UPDATE_X, true, 2>(0, plane, t, nthr, flags, x, r, p, q, r32, z1, a, winv, static_cast<mg_t*>(ho.dst_lo));
            else acc += axpy2_first_range<UPDATE_X, false, 2>(0, plane, t, nthr, flags, x, r, p, q, r32, z1, a, winv, nullptr);
            if (ho.dst_hi) acc += axpy2_first_range<UPDATE_X, true, 2>(top, n, t, nthr, flags, x, r, p, q, r32, z1, a, winv, static_cast<mg_t*>(ho.dst_hi));
            else acc += axpy2_first_range<UPDATE_X, false, 2>(top, n, t, nthr, flags, x, r, p, q, r32, z1, a, winv, nullptr);
            // only the blocks that stored into the neighbours fence at system scope and count in
            halo_publish(ho.counter, (unsigned int)halo_blocks, ho.flag_lo, ho.flag_hi, ho.seq);
        } else if (top > plane) {
            acc = axpy2_first_range<UPDATE_X, false, NCI>(plane, top, (long long)(blockIdx.x - halo_blocks) * VT + threadIdx.x,
                                                        (long long)(gridDim.x - halo_blocks) * VT, flags, x, r, p, q, r32, z1, a,
                                                        winv, nullptr);
        }
    }
    double v[1] = {acc};
    grid_reduce<1>(v, partials, counter, out);
}

// p = z + beta p, skipping 16-byte pairs without an unknown (z = p = 0 there).
// UPDATE_X: first x += alpha p with the OLD p (the solution update deferred from the
// residual kernel: p is read once for both).
// Works on the pairs [lo, hi) with `nthr` threads striding together, this one being thread `t`.
// PUSH: the range is a boundary plane; the new values also go to dst (a neighbour's ghost plane, whose
// element 0 corresponds to element lo).  NC pairs per trip: loads first, then stores.
template <bool UPDATE_X, bool PUSH, int NC>
__device__ __forceinline__ void xpby_range(long long lo, long long hi, long long t, long long nthr,
                                           const uint8_t* __restrict__ flags, double* __restrict__ p,
                                           const mg_t* __restrict__ z, double* __restrict__ x, double bta,
                                           double alpha, double* __restrict__ dst) {
    const long long stride = nthr * 2;
    typedef typename Vec2<mg_t>::type mg2;
    constexpr unsigned int UNK2 = (unsigned int)F_UNK | ((unsigned int)F_UNK << 8);
    const long long i0 = lo + t * 2;
    unsigned int fnext[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const long long ii = i0 + c * stride;
        fnext[c] = (ii < hi) ? *reinterpret_cast<const unsigned short*>(flags + ii) : 0u;
    }
    for (long long i = i0; i < hi; i += NC * stride) {
        unsigned int f[NC];
        double2 pv[NC], xv[NC];
        mg2 zv[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            f[c] = fnext[c];
            const long long in = i + (NC + c) * stride;
            fnext[c] = (in < hi) ? *reinterpret_cast<const unsigned short*>(flags + in) : 0u;
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            if (f[c] & UNK2) {
                pv[c] = *reinterpret_cast<const double2*>(p + i + c * stride);
                zv[c] = *reinterpret_cast<const mg2*>(z + i + c * stride);
                if (UPDATE_X) xv[c] = *reinterpret_cast<const double2*>(x + i + c * stride);
            }
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            if (f[c] & UNK2) {
                if (UPDATE_X) {
                    xv[c].x += alpha * pv[c].x; xv[c].y += alpha * pv[c].y;
                    *reinterpret_cast<double2*>(x + i + c * stride) = xv[c];
                }
                pv[c].x = (double)zv[c].x + bta * pv[c].x; pv[c].y = (double)zv[c].y + bta * pv[c].y;
                *reinterpret_cast<double2*>(p + i + c * stride) = pv[c];
                if (PUSH) *reinterpret_cast<double2*>(dst + (i + c * stride - lo)) = pv[c];
            }
        }
    }
}

// HALO (z-slabs): the first blocks of the grid own the two boundary planes and store the new p into the
// neighbours' ghost planes as well, the rest of the grid runs the plain loop over the interior planes;
// the last block to finish publishes the exchange.
template <bool UPDATE_X, bool HALO, int NCI>
__global__ void __launch_bounds__(VT, NCI == 4 ? 0 : 4)
xpby_kernel(long long n, const uint8_t* __restrict__ flags, double* __restrict__ p, const mg_t* __restrict__ z,
            const double* __restrict__ num, const double* __restrict__ den, double* __restrict__ x,
            const double* __restrict__ anum, const double* __restrict__ aden, long long plane, int halo_blocks,
            HaloOut ho) {
    const double bta = num[0] / den[0];
    const double alpha = UPDATE_X ? anum[0] / aden[0] : 0.0;
    const long long n2 = n & ~1LL;
    if (!HALO) {
        xpby_range<UPDATE_X, false, NCI>(0, n2, (long long)blockIdx.x * VT + threadIdx.x, (long long)gridDim.x * VT,
                                       flags, p, z, x, bta, alpha, nullptr);
        if (blockIdx.x == 0 && threadIdx.x == 0 && n2 < n) {
            if (UPDATE_X) x[n2] += alpha * p[n2];
            p[n2] = (double)z[n2] + bta * p[n2];
        }
    } else {
        // plane and n are even here (checked by the launcher)
        const long long top = n - plane;
        if ((int)blockIdx.x < halo_blocks) {
            const long long t = (long long)blockIdx.x * VT + threadIdx.x, nthr = (long long)halo_blocks * VT;
            if (ho.dst_lo) xpby_range<UPDATE_X, true, 2>(0, plane, t, nthr, flags, p, z, x, bta, alpha, static_cast<double*>(ho.dst_lo));
            else xpby_range<UPDATE_X, false, 2>(0, plane, t, nthr, flags, p, z, x, bta, alpha, nullptr);
            // (at least two planes, checked by the launcher: the two boundary planes are distinct)
            if (ho.dst_hi) xpby_range<UPDATE_X, true, 2>(top, n, t, nthr, flags, p, z, x, bta, alpha, static_cast<double*>(ho.dst_hi));
            else xpby_range<UPDATE_X, false, 2>(top, n, t, nthr, flags, p, z, x, bta, alpha, nullptr);
            // only the blocks that stored into the neighbours fence at system scope and count in
            halo_publish(ho.counter, (unsigned int)halo_blocks, ho.flag_lo, ho.flag_hi, ho.seq);
        } else if (top > plane) {
            xpby_range<UPDATE_X, false, NCI>(plane, top, (long long)(blockIdx.x - halo_blocks) * VT + threadIdx.x,
                                             (long long)(gridDim.x - halo_blocks) * VT, flags, p, z, x, bta, alpha, nullptr);
        }
    }
}

// x += (num/den) p : the deferred solution update when the loop ends between two xpby
__global__ void __launch_bounds__(VT)
axpy_kernel(long long n, const uint8_t* __restrict__ flags, double* __restrict__ x, const double* __restrict__ p,
            const double* __restrict__ num, const double* __restrict__ den) {
    const double a = num[0] / den[0];
    const long long stride = (long long)gridDim.x * VT;
    for (long long i = (long long)blockIdx.x * VT + threadIdx.x; i < n; i += stride)
        if (flags[i] & F_UNK) x[i] += a * p[i];
}

template <typename D, typename S>
__global__ void __launch_bounds__(VT)
convert_kernel(long long n, D* __restrict__ d, const S* __restrict__ s) {
    const long long stride = (long long)gridDim.x * VT;
    for (long long i = (long long)blockIdx.x * VT + threadIdx.x; i < n; i += stride) d[i] = (D)s[i];
}

__global__ void __launch_bounds__(VT)
dot_kernel(long long n, const double* __restrict__ a, const double* __restrict__ b,
           double* partials, unsigned int* counter, double* out) {
    const long long stride = (long long)gridDim.x * VT;
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * VT + threadIdx.x; i < n; i += stride) acc += a[i] * b[i];
    double v[1] = {acc};
    grid_reduce<1>(v, partials, counter, out);
}

__global__ void __launch_bounds__(VT)
copy_kernel(long long n, double* __restrict__ d, const double* __restrict__ s) {
    const long long stride = (long long)gridDim.x * VT;
    for (long long i = (long long)blockIdx.x * VT + threadIdx.x; i < n; i += stride) d[i] = s[i];
}

__global__ void __launch_bounds__(VT)
jacobi_precond_dot_kernel(Grid g, const uint8_t* __restrict__ flags, const double* __restrict__ r,
                          mg_t* __restrict__ z, long long n, double* partials,
                          unsigned int* counter, double* out) {
    const long long stride = (long long)gridDim.x * VT;
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * VT + threadIdx.x; i < n; i += stride) {
        const uint8_t f = flags[i];
        double zv = 0.0;
        if (f & F_UNK) {
            const double d = row_diag<double>(f, g);
            const double rv = r[i];
            zv = rv / d;
            acc += rv * zv;
        }
        z[i] = (mg_t)zv;
    }
    double v[1] = {acc};
    grid_reduce<1>(v, partials, counter, out);
}

inline int nblocks(long long work_items, int n_sm) {
    long long b = (work_items + VT - 1) / VT;
    const long long cap = (long long)n_sm * 8;     // 8 x 256 threads resident per SM
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}
}  // namespace

int vec_max_blocks(int n_sm) { return n_sm * 8; }

// OI_VEC_NC=2|4: pairs per trip of the residual-update and xpby kernels.  Measured at 1024^3
// (profiles/r2_variants_ab.md): xpby 4.77 ms with 2 pairs and four blocks per SM against 4.91 ms with 4,
// the residual update 6.19 against 6.13 -- so xpby defaults to 2, the residual update to 4.
static int vec_nc(int dflt) {
    const char* e = getenv("OI_VEC_NC");
    return (e && (e[0] == '2' || e[0] == '4')) ? e[0] - '0' : dflt;
}

void vec_axpy2_dot(long long n, double* x, double* r, const double* p, const double* q,
                   const double* num, const double* den, double* partials, unsigned int* counter,
                   double* out, int n_sm, cudaStream_t st) {
    axpy2_dot_kernel<<<nblocks((n + 1) / 2, n_sm), VT, 0, st>>>(n, x, r, p, q, num, den, partials, counter, out);
}
void vec_axpy2_dot_first(const Grid& g, const uint8_t* flags, long long n, double* x, double* r,
                         const double* p, const double* q, mg_t* r32, mg_t* z1, const double* num,
                         const double* den, double w0, double* partials, unsigned int* counter,
                         double* out, int n_sm, cudaStream_t st, const HaloOut* ho) {
    const int nb = nblocks((n + 1) / 2, n_sm);
    const HaloOut none{};
    if (ho && x == nullptr && vec_halo_supported(g.plane, n)) {
        int hb = (int)((8 * g.plane * (long long)nb + n - 1) / n);        // four times the fair share (see vec_xpby)
        hb = hb < 8 ? 8 : hb;
        // (the grid reduction's scratch holds vec_max_blocks + a margin of blocks: stay within nb)
        const int nbt = nb;
        hb = hb >= nbt ? nbt - 1 : hb;
        if (hb < 1) hb = 1;
        axpy2_dot_first_kernel<false, true, 4><<<nbt, VT, 0, st>>>(g, flags, n, x, r, p, q, r32, z1, num, den, w0, partials,
                                                                   counter, out, hb, *ho);
    } else if (x)
        axpy2_dot_first_kernel<true, false, 4><<<nb, VT, 0, st>>>(g, flags, n, x, r, p, q, r32, z1, num, den, w0, partials,
                                                                  counter, out, 0, none);
    else if (vec_nc(4) == 2)       // OI_VEC_NC=2: two pairs per trip, four blocks per SM (A/B knob)
        axpy2_dot_first_kernel<false, false, 2><<<nb, VT, 0, st>>>(g, flags, n, x, r, p, q, r32, z1, num, den, w0, partials,
                                                                   counter, out, 0, none);
    else
        axpy2_dot_first_kernel<false, false, 4><<<nb, VT, 0, st>>>(g, flags, n, x, r, p, q, r32, z1, num, den, w0, partials,
                                                                   counter, out, 0, none);
}
void vec_xpby(long long n, const uint8_t* flags, double* p, const mg_t* z, const double* num,
              const double* den, double* x, const double* anum, const double* aden, int n_sm, cudaStream_t st,
              long long plane, const HaloOut* ho) {
    const int nb = nblocks((n + 1) / 2, n_sm);
    const HaloOut none{};
    if (ho && vec_halo_supported(plane, n)) {
        // blocks for the two boundary planes: four times their fair share (their loop carries fewer loads in
        // flight and stores across NVLink), so that they are never the tail of the kernel
        int hb = (int)((8 * plane * (long long)nb + n - 1) / n);
        hb = hb < 8 ? 8 : (hb > nb / 4 ? nb / 4 : hb);
        hb = hb < 1 ? 1 : hb;
        const int nbt = nb + hb;
        if (x) xpby_kernel<true, true, 2><<<nbt, VT, 0, st>>>(n, flags, p, z, num, den, x, anum, aden, plane, hb, *ho);
        else xpby_kernel<false, true, 4><<<nbt, VT, 0, st>>>(n, flags, p, z, num, den, nullptr, nullptr, nullptr, plane, hb, *ho);
    } else {
        if (x && vec_nc(2) == 2) xpby_kernel<true, false, 2><<<nb, VT, 0, st>>>(n, flags, p, z, num, den, x, anum, aden, 0, 0, none);
        else if (x) xpby_kernel<true, false, 4><<<nb, VT, 0, st>>>(n, flags, p, z, num, den, x, anum, aden, 0, 0, none);
        else xpby_kernel<false, false, 4><<<nb, VT, 0, st>>>(n, flags, p, z, num, den, nullptr, nullptr, nullptr, 0, 0, none);
    }
}
// whether the two kernels above can carry the boundary-plane push for this plane size
// (even plane: a pair never straddles two planes; at least two planes: the boundary planes are distinct)
bool vec_halo_supported(long long plane, long long n) { return plane > 0 && (plane & 1) == 0 && n >= 2 * plane; }
void vec_axpy(long long n, const uint8_t* flags, double* x, const double* p, const double* num, const double* den,
              int n_sm, cudaStream_t st) {
    axpy_kernel<<<nblocks(n, n_sm), VT, 0, st>>>(n, flags, x, p, num, den);
}
void vec_to_mg(long long n, mg_t* dst, const double* src, int n_sm, cudaStream_t st) {
    convert_kernel<mg_t, double><<<nblocks(n, n_sm), VT, 0, st>>>(n, dst, src);
}
void vec_from_mg(long long n, double* dst, const mg_t* src, int n_sm, cudaStream_t st) {
    convert_kernel<double, mg_t><<<nblocks(n, n_sm), VT, 0, st>>>(n, dst, src);
}
void vec_dot(long long n, const double* a, const double* b, double* partials, unsigned int* counter,
             double* out, int n_sm, cudaStream_t st) {
    dot_kernel<<<nblocks(n, n_sm), VT, 0, st>>>(n, a, b, partials, counter, out);
}
void vec_copy(long long n, double* dst, const double* src, int n_sm, cudaStream_t st) {
    copy_kernel<<<nblocks(n, n_sm), VT, 0, st>>>(n, dst, src);
}
void l0_jacobi_precond_dot(const Grid& g, const uint8_t* flags, const double* r, mg_t* z,
                           double* partials, unsigned int* counter, double* out, int n_sm,
                           cudaStream_t st) {
    const long long n = (long long)g.nz * g.plane;
    jacobi_precond_dot_kernel<<<nblocks(n, n_sm), VT, 0, st>>>(g, flags, r, z, n, partials, counter, out);
}

}  // namespace oi
