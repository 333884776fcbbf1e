// Level-0 stencil, shared-memory ring variant (default when nx % 4 == 0).
//
// ncu on the register z-march kernel showed it issue-bound, not DRAM-bound
// (sm__inst_issued 69 %, dram 47 %, ~160 thread-instructions per cell): selects on
// six face bits, 64-bit index products, an fp64 divide and int->double converts
// per cell, and a CTA barrier per plane.  This kernel puts the instruction
// stream on a diet and lets the copy side of the LSU do the staging:
//
//   * every thread owns 16 bytes of x-adjacent cells (2 fp64 or 4 fp32 cells):
//     16-byte cp.async / LDS.128 / STG.128;
//   * plane k+RING_P+1 of the input, of the rhs and of the connectivity bytes is
//     already in flight into a RING_R-stage shared-memory ring while plane k is
//     computed (cp.async groups, one per plane), so a CTA keeps ~5 planes of loads
//     outstanding without spending registers (a register queue for the
//     connectivity bytes was the top stall of the previous version);
//   * the plane loop is unrolled by the ring length, so stage numbers and the
//     register window over the column are compile-time constants;
//   * no face selects: every vector A is applied to is zero on inactive cells
//     (x0, p, z are built that way and the updates preserve it; out-of-box halo
//     cells are zero-filled by cp.async, or wrapped when the box is periodic), so
//         (A u)_c = d_c u_c - cx (u_w + u_e) - cy (u_s + u_n) - cz (u_d + u_u)
//     with d_c and 1/d_c looked up by the 6 face bits in a 64-entry table;
//   * the z neighbours travel in registers (one centre LDS per plane, not three).
//
// Element type: double for the Krylov operator apply, mg_t for multigrid sweeps.
// Algorithmic traffic per cell: APPLY 17 B (fp64); SMOOTH 2*4+1+4 = 13 B and
// RESTRICT 9.5 B with mg_t = float (25 / 17.1 B with an fp64 V-cycle).
#include "oi_kernels.h"

namespace oi {

namespace {

constexpr int RING_P = 4;                  // planes in flight beyond k+1
constexpr int RING_R = RING_P + 2;         // stages: planes k-1 .. k+RING_P

template <typename T>
struct Cfg {
    static constexpr int CPT = 16 / (int)sizeof(T);   // cells per thread (2 fp64 / 4 fp32)
    static constexpr int TX = 64;                     // CTA tile width in cells
    static constexpr int XT = TX / CPT;               // threads along x (32 / 16)
    static constexpr int NT = 256;
    static constexpr int TY = NT / XT;                // CTA tile height (8 / 16 rows)
    static constexpr int PITCH = TX + 2 * CPT;        // [pad.., west halo | 64 centres | east halo, pad..]
    static constexpr int U_STAGE = (TY + 2) * PITCH;  // elements per field stage (south halo, rows, north halo)
    static constexpr int B_STAGE = TY * TX;           // elements per rhs stage
    static constexpr int F_STAGE = TY * TX;           // bytes per connectivity stage
    static constexpr int FCOPY = 4 / CPT;             // threads sharing one 4-byte flag copy (2 / 1)
};

template <typename T>
struct alignas(16) Pack { T v[16 / sizeof(T)]; };

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 16 : 0;          // src-size 0 -> destination zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gsrc), "r"(sz) : "memory");
}
template <int BYTES>
__device__ __forceinline__ void cp_async_small(void* smem_dst, const void* gsrc, bool valid) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? BYTES : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;\n" ::"r"(sa), "l"(gsrc), "n"(BYTES), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

struct RingCoarse { int cnx, cny, z0h; };   // RESTRICT target dims, (z0 >> 1)

// MODE 0 APPLY: out = w * A u (+ dot u.out) ; 1 SMOOTH: out = u + w (b - A u)/d (+ dot b.out)
// MODE 2 RESTRICT (2x2x2 or 2x2x1): out[coarse] = sum_children (b - A u)
// HALO (z-slabs): the CTAs of the first / last z-chunk wait for the neighbours' boundary planes on the
// flag words of `hin` before they touch a ghost plane, store plane 0 / nz-1 of `out` into the
// neighbours' ghost planes as they produce them (MODE 1) and publish `hout.seq`; those two chunks are
// dispatched last (blockIdx.z is remapped), so transfer and rank skew hide behind the interior planes.
template <typename T, int MODE, bool DOT, bool HALO>
__global__ void __launch_bounds__(256)
l0_ring_kernel(Grid g, const uint8_t* __restrict__ flags, const T* __restrict__ u,
               const T* __restrict__ b, T* __restrict__ out, T w, RingCoarse rc, int fz, int zchunk,
               double* red_partials, unsigned int* red_counter, double* red_out, HaloIn hin, HaloOut hout, int bnd_only) {
    typedef Cfg<T> C;
    constexpr int CPT = C::CPT, TX = C::TX, TY = C::TY, PITCH = C::PITCH;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* us = reinterpret_cast<T*>(smem_raw);                     // [RING_R][U_STAGE]
    T* bs = us + RING_R * C::U_STAGE;                           // [RING_R][B_STAGE]   (MODE != 0)
    T* dtab = bs + (MODE != 0 ? RING_R * C::B_STAGE : 0);       // [64] diagonal, [64] its inverse
    unsigned char* fs = reinterpret_cast<unsigned char*>(dtab + 128);   // [RING_R][F_STAGE]

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int WX = C::XT / 16;                              // warps side by side in x (2 / 1)
    const int tx = (warp % WX) * 16 + (lane & 15);              // index of this thread's cell group
    const int ty = (warp / WX) * 2 + (lane >> 4);
    const int i = blockIdx.x * TX + CPT * tx;                   // first of the CPT cells (multiple of CPT)
    const int j = blockIdx.y * TY + ty;
    const bool inb = (i < g.nx) && (j < g.ny);                  // nx % 4 == 0: all CPT cells in or out
    int zc_idx = blockIdx.z;
    if (HALO) {                                                 // interior chunks first, chunk 0 and the last one at the end
        const int nzc = gridDim.z;
        if (nzc > 2) zc_idx = ((int)blockIdx.z < nzc - 2) ? (int)blockIdx.z + 1 : ((int)blockIdx.z == nzc - 2 ? 0 : nzc - 1);
    }
    // bnd_only (HALO, MODE 1): two one-plane chunks, plane 0 and plane nz-1; nothing is stored locally
    const bool bnd = HALO && MODE == 1 && bnd_only != 0;
    const int k0 = bnd ? (blockIdx.z == 0 ? 0 : g.nz - 1) : zc_idx * zchunk;
    const int k1 = bnd ? k0 + 1 : min(k0 + zchunk, g.nz);
    const T cx = (T)g.cx, cy = (T)g.cy, cz = (T)g.cz;

    if (HALO) {
        const bool wlo = hin.flag_lo && k0 == 0, whi = hin.flag_hi && k1 == g.nz;
        if (wlo || whi) {
            if (tid == 0) {
                if (wlo) halo_spin(hin.flag_lo, hin.seq);
                if (whi) halo_spin(hin.flag_hi, hin.seq);
            }
            __syncthreads();
        }
    }
    const bool push_lo = HALO && MODE == 1 && hout.dst_lo && k0 == 0;
    const bool push_hi = HALO && MODE == 1 && hout.dst_hi && k1 == g.nz;

    if (tid < 64) {
        const T d = row_diag<T>((unsigned int)tid, g);
        dtab[tid] = d;
        dtab[64 + tid] = d > (T)0 ? (T)1 / d : (T)0;
    }

    // halo duties
    const bool hw = (tx == 0), he = (tx == C::XT - 1);
    const bool hs = (ty == 0), hn = (ty == TY - 1);
    // (periodic box: the halo of an edge tile comes from the opposite side)
    const int iw = (i < g.nx) ? wrap_lo(i, g.nx, g.periodic & PER_X) : -1;
    const int ie = (i + CPT < g.nx) ? i + CPT : ((i < g.nx && (g.periodic & PER_X)) ? 0 : -1);
    const int js = (j < g.ny) ? wrap_lo(j, g.ny, g.periodic & PER_Y) : -1;
    const int jn = (j < g.ny) ? wrap_hi(j, g.ny, g.periodic & PER_Y) : -1;
    const bool hw_ok = hw && (j < g.ny) && iw >= 0;
    const bool he_ok = he && (j < g.ny) && ie >= 0;
    const bool hs_ok = hs && (i < g.nx) && js >= 0;
    const bool hn_ok = hn && (i < g.nx) && jn >= 0;

    const long long col = inb ? (long long)j * g.nx + i : 0;
    // periodic box, partial edge tile: the first cell group / row past the edge stands in
    // for the east / north halo and loads the wrapped column 0 / row 0
    const bool wrap_col = (g.periodic & PER_X) && i == g.nx;
    const bool wrap_row = (g.periodic & PER_Y) && j == g.ny;
    const int il = wrap_col ? 0 : i, jl = wrap_row ? 0 : j;
    const bool ld_ok = (il < g.nx) && (jl < g.ny);
    const T* u_own = u + (ld_ok ? (long long)jl * g.nx + il : 0);
    const T* u_w = u + (hw_ok ? (long long)j * g.nx + iw : 0);
    const T* u_e = u + (he_ok ? (long long)j * g.nx + ie : 0);
    const T* u_s = u + (hs_ok ? (long long)js * g.nx + i : 0);
    const T* u_n = u + (hn_ok ? (long long)jn * g.nx + i : 0);
    const T* b_own = (MODE != 0) ? b + col : nullptr;
    const uint8_t* f_own = flags + col;

    const int c_off = (ty + 1) * PITCH + CPT + CPT * tx;        // own centre group inside a field stage
    const int b_off = ty * TX + CPT * tx;                       // inside a rhs stage / flag stage
    const bool f_copy = (tx % C::FCOPY) == 0;                   // this thread copies 4 connectivity bytes

    // issue all copies of plane kk into stage st (poff = kk * plane, carried by the caller);
    // rhs and connectivity bytes only exist for planes k0 .. k1-1
    auto issue = [&](long long poff, int st, bool interior) {
        T* S = us + st * C::U_STAGE;
        cp_async16(S + c_off, u_own + poff, ld_ok);
        if (hw) cp_async_small<(int)sizeof(T)>(S + (ty + 1) * PITCH + CPT - 1, u_w + poff, hw_ok);
        if (he) cp_async_small<(int)sizeof(T)>(S + (ty + 1) * PITCH + CPT + TX, u_e + poff, he_ok);
        if (hs) cp_async16(S + CPT + CPT * tx, u_s + poff, hs_ok);
        if (hn) cp_async16(S + (TY + 1) * PITCH + CPT + CPT * tx, u_n + poff, hn_ok);
        if (interior) {
            if (MODE != 0) cp_async16(bs + st * C::B_STAGE + b_off, b_own + poff, inb);
            if (f_copy) cp_async_small<4>(fs + st * C::F_STAGE + b_off, f_own + poff, inb);
        }
    };

    // prologue: planes k0-1 .. k0+RING_P, one commit group per plane (empty groups
    // keep the count uniform at the end of the chunk)
    long long poff_issue = (long long)(k0 - 1) * g.plane;
    int kk_issue = k0 - 1;
#pragma unroll
    for (int s = 0; s < RING_R; ++s) {
        if (kk_issue <= k1) issue(poff_issue, s, kk_issue >= k0 && kk_issue < k1);
        cp_async_commit();
        poff_issue += g.plane;
        ++kk_issue;
    }

    cp_async_wait<RING_P>();                                    // planes k0-1 and k0 have landed
    __syncthreads();
    // register window over the column: vv[s % 3] = plane k-1, vv[(s+1) % 3] = plane k
    Pack<T> vv[3];
    vv[0] = *reinterpret_cast<const Pack<T>*>(us + 0 * C::U_STAGE + c_off);
    vv[1] = *reinterpret_cast<const Pack<T>*>(us + 1 * C::U_STAGE + c_off);

    double dot_acc = 0.0;
    T zpair[CPT / 2];
#pragma unroll
    for (int q = 0; q < CPT / 2; ++q) zpair[q] = (T)0;
    T* out_own = out + col + (long long)k0 * g.plane;           // MODE 0/1

    // The plane loop is unrolled by the ring length: plane kb+s always lives in
    // stage (s+1) % RING_R.
    for (int kb = k0; kb < k1; kb += RING_R) {
#pragma unroll
        for (int s = 0; s < RING_R; ++s) {
            const int k = kb + s;
            if (k >= k1) break;                                 // CTA-uniform
            const int sc = (s + 1) % RING_R, sp = (s + 2) % RING_R;   // stages of planes k, k+1
            cp_async_wait<RING_P - 1>();                        // own copies of planes <= k+1 landed
            __syncthreads();                                    // ... and everybody else's

            const T* Sc = us + sc * C::U_STAGE;
            const Pack<T> v_m = vv[s % 3], v_c = vv[(s + 1) % 3];
            const Pack<T> v_p = *reinterpret_cast<const Pack<T>*>(us + sp * C::U_STAGE + c_off);
            vv[(s + 2) % 3] = v_p;
            const T xw = Sc[c_off - 1], xe = Sc[c_off + CPT];
            const Pack<T> ys = *reinterpret_cast<const Pack<T>*>(Sc + c_off - PITCH);
            const Pack<T> yn = *reinterpret_cast<const Pack<T>*>(Sc + c_off + PITCH);
            Pack<T> bb;
            if (MODE != 0) bb = *reinterpret_cast<const Pack<T>*>(bs + sc * C::B_STAGE + b_off);

            unsigned int fword;
            if (CPT == 2) fword = *reinterpret_cast<const unsigned short*>(fs + sc * C::F_STAGE + b_off);
            else          fword = *reinterpret_cast<const unsigned int*>(fs + sc * C::F_STAGE + b_off);

            Pack<T> o;
            T res[CPT / 2];
#pragma unroll
            for (int q = 0; q < CPT / 2; ++q) res[q] = (T)0;
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const unsigned int f = (fword >> (8 * c)) & 0xffu;
                const T west = (c == 0) ? xw : v_c.v[c > 0 ? c - 1 : 0];
                const T east = (c == CPT - 1) ? xe : v_c.v[c < CPT - 1 ? c + 1 : CPT - 1];
                const T au = dtab[f & 63u] * v_c.v[c] -
                             (cx * (west + east) + cy * (ys.v[c] + yn.v[c]) + cz * (v_m.v[c] + v_p.v[c]));
                const bool unk = (f & F_UNK) != 0;
                if (MODE == 0) {
                    o.v[c] = unk ? w * au : (T)0;
                    if (DOT) dot_acc += (double)v_c.v[c] * (double)o.v[c];
                } else if (MODE == 1) {
                    o.v[c] = unk ? v_c.v[c] + w * (bb.v[c] - au) * dtab[64 + (f & 63u)] : (T)0;
                    if (DOT) dot_acc += (double)bb.v[c] * (double)o.v[c];
                } else {
                    if (unk) res[c >> 1] += bb.v[c] - au;       // x pair summed in-thread
                }
            }
            if (MODE != 2) {
                // a 16-byte group without an unknown stays zero (every output field is
                // zero off the unknowns from allocation on): its sector is never written
                constexpr unsigned int UNKS = (CPT == 2) ? 0x4040u : 0x40404040u;
                if (inb && (fword & UNKS)) {
                    if (!bnd) *reinterpret_cast<Pack<T>*>(out_own) = o;
                    if (HALO) {
                        if (push_lo && k == 0) *reinterpret_cast<Pack<T>*>(static_cast<T*>(hout.dst_lo) + col) = o;
                        if (push_hi && k == g.nz - 1) *reinterpret_cast<Pack<T>*>(static_cast<T*>(hout.dst_hi) + col) = o;
                    }
                }
                out_own += g.plane;
            } else {
                // y pair = lane ^ 16; z pair carried across two planes
                const int kg = g.z0 + k;
                const bool even = (fz == 2) && ((kg & 1) == 0);
                const bool flush = !even || (k + 1 == k1);
#pragma unroll
                for (int q = 0; q < CPT / 2; ++q) {
                    T sum = res[q] + __shfl_xor_sync(0xffffffffu, res[q], 16);
                    if (fz == 2) {
                        if (even) zpair[q] = sum;
                        else { sum += zpair[q]; zpair[q] = (T)0; }
                    }
                    if (flush && inb && ((j & 1) == 0)) {
                        const int ck = (fz == 2) ? ((kg >> 1) - rc.z0h) : k;
                        out[((long long)ck * rc.cny + (j >> 1)) * rc.cnx + (i >> 1) + q] = sum;
                    }
                }
            }

            // refill the stage that held plane k-1 (stage s) with plane k+RING_P+1.  Its
            // halo cells are dead and its centre cells are only ever read by their own
            // thread (as v_p two iterations ago), so no second barrier is needed.
            if (kk_issue <= k1) issue(poff_issue, s, kk_issue < k1);
            cp_async_commit();
            poff_issue += g.plane;
            ++kk_issue;
        }
    }
    cp_async_wait<0>();

    if (HALO && MODE == 1) {
        const unsigned int tiles = gridDim.x * gridDim.y;
        if (hout.flag_lo && k0 == 0) halo_publish(hout.counter + 0, tiles, hout.flag_lo, nullptr, hout.seq);
        if (hout.flag_hi && k1 == g.nz) halo_publish(hout.counter + 1, tiles, nullptr, hout.flag_hi, hout.seq);
    }

    if (DOT) {
        double v[1] = {dot_acc};
        grid_reduce<1>(v, red_partials, red_counter, red_out);
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
prolong_add_kernel(Grid g, const uint8_t* __restrict__ flags, T* __restrict__ z,
                   const T* __restrict__ ec, int cnx, int cny, int fx, int fy, int fz) {
    const int i = blockIdx.x * 64 + (threadIdx.x & 63);
    const int j = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (i >= g.nx || j >= g.ny) return;
    const int ci = (fx == 2) ? (i >> 1) : i, cj = (fy == 2) ? (j >> 1) : j;
    const long long col = (long long)j * g.nx + i;
    const long long ccol = (long long)cj * cnx + ci;
    const long long cplane = (long long)cnx * cny;
    for (int k = blockIdx.z; k < g.nz; k += gridDim.z) {
        const long long idx = (long long)k * g.plane + col;
        if (flags[idx] & F_UNK) {
            const int ck = (fz == 2) ? (k >> 1) : k;
            z[idx] += ec[(long long)ck * cplane + ccol];
        }
    }
}

// 4 cells per thread (fp32, nx % 4 == 0): float4 z, 4 connectivity bytes, float2 coarse values.
// HALO (z-slabs): the corrected boundary planes also go into the neighbours' ghost planes.
template <bool HALO>
__global__ void __launch_bounds__(256)
prolong_add_vec4_kernel(Grid g, const uint8_t* __restrict__ flags, float* __restrict__ z,
                        const float* __restrict__ ec, int cnx, int cny, int fy, int fz, HaloOut ho) {
    const int i = blockIdx.x * 64 + (threadIdx.x & 15) * 4;
    const int j = blockIdx.y * 16 + (threadIdx.x >> 4);
    if (i < g.nx && j < g.ny) {
        const int cj = (fy == 2) ? (j >> 1) : j;
        const long long col = (long long)j * g.nx + i;
        const long long ccol = (long long)cj * cnx + (i >> 1);      // fx == 2
        const long long cplane = (long long)cnx * cny;
        for (int k = blockIdx.z; k < g.nz; k += gridDim.z) {
            const long long idx = (long long)k * g.plane + col;
            const unsigned int f = *reinterpret_cast<const unsigned int*>(flags + idx);
            if ((f & 0x40404040u) == 0u) continue;
            const int ck = (fz == 2) ? (k >> 1) : k;
            const float2 e = *reinterpret_cast<const float2*>(ec + (long long)ck * cplane + ccol);
            float4 v = *reinterpret_cast<float4*>(z + idx);
            if (f & 0x00000040u) v.x += e.x;
            if (f & 0x00004000u) v.y += e.x;
            if (f & 0x00400000u) v.z += e.y;
            if (f & 0x40000000u) v.w += e.y;
            *reinterpret_cast<float4*>(z + idx) = v;
            if (HALO) {
                if (ho.dst_lo && k == 0) *reinterpret_cast<float4*>(static_cast<float*>(ho.dst_lo) + col) = v;
                if (ho.dst_hi && k == g.nz - 1) *reinterpret_cast<float4*>(static_cast<float*>(ho.dst_hi) + col) = v;
            }
        }
    }
    if (HALO) {
        // only the blocks that own a boundary plane stored into a neighbour: they alone fence and count in
        const unsigned int tiles = gridDim.x * gridDim.y;
        if (ho.flag_lo && blockIdx.z == 0) halo_publish(ho.counter + 0, tiles, ho.flag_lo, nullptr, ho.seq);
        if (ho.flag_hi && (int)blockIdx.z == (g.nz - 1) % (int)gridDim.z) halo_publish(ho.counter + 1, tiles, nullptr, ho.flag_hi, ho.seq);
    }
}

template <typename T, int MODE>
size_t ring_smem_bytes() {
    typedef Cfg<T> C;
    return sizeof(T) * (size_t)(RING_R * C::U_STAGE + (MODE != 0 ? RING_R * C::B_STAGE : 0) + 128) +
           (size_t)RING_R * C::F_STAGE;
}

template <typename T, int MODE, bool DOT, bool HALO>
void launch_h(const L0Args& a, cudaStream_t st) {
    typedef Cfg<T> C;
    static unsigned long long configured = 0;
    const size_t smem = ring_smem_bytes<T, MODE>();
    if (first_use_on_this_device(configured))
        cudaFuncSetAttribute(l0_ring_kernel<T, MODE, DOT, HALO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int zc = pick_zchunk(a.g, a.n_sm);
    dim3 grid((a.g.nx + C::TX - 1) / C::TX, (a.g.ny + C::TY - 1) / C::TY, (a.g.nz + zc - 1) / zc);
    const int bnd = (HALO && MODE == 1 && a.bnd_only && a.g.nz >= 2) ? 1 : 0;
    if (bnd) grid.z = 2;
    RingCoarse rc{a.cnx, a.cny, a.g.z0 >> 1};
    l0_ring_kernel<T, MODE, DOT, HALO><<<grid, C::NT, smem, st>>>(
        a.g, a.flags, static_cast<const T*>(a.u), static_cast<const T*>(a.b), static_cast<T*>(a.out), (T)a.w, rc,
        a.fz, zc, a.red_partials, a.red_counter, a.red_out, a.hin, a.hout, bnd);
}

template <typename T, int MODE, bool DOT>
void launch(const L0Args& a, cudaStream_t st) {
    const bool halo = a.hin.flag_lo || a.hin.flag_hi || a.hout.flag_lo || a.hout.flag_hi;
    if (halo) launch_h<T, MODE, DOT, true>(a, st);
    else launch_h<T, MODE, DOT, false>(a, st);
}

}  // namespace

// whether l0_prolong_add runs the kernel that can carry the boundary-plane push
bool prolong_halo_supported(const L0Args& a) { return sizeof(mg_t) == 4 && (a.g.nx & 3) == 0 && a.fx == 2; }

bool ring_supported(const L0Args& a, int mode) {
    // 16-byte vector accesses and the 4-byte copies of the connectivity bytes need
    // nx % 4 == 0 (then every row and plane start is suitably aligned: plane 0 of
    // every Field is 256-byte aligned)
    if (a.g.nx & 3) return false;
    if (mode == 2 && !(a.fx == 2 && a.fy == 2)) return false;
    return true;
}

void ring_launch(const L0Args& a, int mode, bool dot, cudaStream_t st) {
    if (mode == 0) { if (dot) launch<double, 0, true>(a, st); else launch<double, 0, false>(a, st); }
    else if (mode == 1) { if (dot) launch<mg_t, 1, true>(a, st); else launch<mg_t, 1, false>(a, st); }
    else launch<mg_t, 2, false>(a, st);
}

void l0_prolong_add(const L0Args& a, cudaStream_t st) {
    if (sizeof(mg_t) == 4 && (a.g.nx & 3) == 0 && a.fx == 2) {
        int gzv = a.g.nz < 128 ? a.g.nz : 128;
        dim3 gridv((a.g.nx + 63) / 64, (a.g.ny + 15) / 16, gzv);
        if (a.hout.flag_lo || a.hout.flag_hi)
            prolong_add_vec4_kernel<true><<<gridv, 256, 0, st>>>(a.g, a.flags, reinterpret_cast<float*>(a.out),
                                                                 reinterpret_cast<const float*>(a.ec), a.cnx, a.cny, a.fy, a.fz, a.hout);
        else
            prolong_add_vec4_kernel<false><<<gridv, 256, 0, st>>>(a.g, a.flags, reinterpret_cast<float*>(a.out),
                                                                  reinterpret_cast<const float*>(a.ec), a.cnx, a.cny, a.fy, a.fz, a.hout);
        return;
    }
    int gz = a.g.nz < 64 ? a.g.nz : 64;
    dim3 grid((a.g.nx + 63) / 64, (a.g.ny + 3) / 4, gz);
    prolong_add_kernel<mg_t><<<grid, 256, 0, st>>>(a.g, a.flags, static_cast<mg_t*>(a.out),
                                                   static_cast<const mg_t*>(a.ec), a.cnx, a.cny, a.fx, a.fy, a.fz);
}

}  // namespace oi
