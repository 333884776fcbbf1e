// Level-0 stencil, shared-memory ring variant (default on even nx).
//
// ncu on the register z-march kernel showed it issue-bound, not DRAM-bound
// (sm__inst_issued 69 %, dram 47 %, ~160 thread-instructions per cell): selects on
// six face bits, 64-bit index products, an fp64 divide and int->double converts
// per cell, and a CTA barrier per plane.  This kernel puts the instruction
// stream on a diet and lets the copy engine side of the LSU do the staging:
//
//   * every thread owns TWO x-adjacent cells: 16-byte cp.async / LDS.128 / STG.128;
//   * plane k+RING_P+1 of the input (and of the rhs) is already in flight into a
//     RING_R-stage shared-memory ring while plane k is computed (cp.async groups,
//     one per plane), so a CTA keeps ~5 planes x 9.5 KB of loads outstanding;
//   * no face selects: every vector A is applied to is zero on inactive cells
//     (x0, p, z are built that way and the updates preserve it; out-of-box halo
//     cells are zero-filled by cp.async), so
//         (A u)_c = d_c u_c - cx (u_w + u_e) - cy (u_s + u_n) - cz (u_d + u_u)
//     with d_c and 1/d_c looked up by the 6 face bits in a 64-entry table;
//   * the z neighbours travel in registers (one centre LDS per plane, not three).
//
// Algorithmic traffic is unchanged: 17 B/cell (APPLY), 25 (SMOOTH), 17.1 (RESTRICT).
#include "oi_kernels.h"

namespace oi {

namespace {

constexpr int TX = 64, TY = 8;             // CTA tile in cells
constexpr int NT = (TX / 2) * TY;          // 256 threads, two cells each
constexpr int RING_P = 4;                  // planes in flight beyond k+1
constexpr int RING_R = RING_P + 2;         // stages: planes k-1 .. k+RING_P
constexpr int PITCH = TX + 4;              // [pad, west halo, 64 centres, east halo, pad]
constexpr int U_STAGE = (TY + 2) * PITCH;  // doubles per u stage (rows: south halo, 8, north halo)
constexpr int B_STAGE = TY * TX;           // doubles per rhs stage

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 16 : 0;          // src-size 0 -> destination zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc, bool valid) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sa), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc, bool valid) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(sa), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

struct RingCoarse { int cnx, cny, z0h; };   // RESTRICT target dims, (z0 >> 1)

// MODE 0 APPLY: out = w * A u (+ dot u.out) ; 1 SMOOTH: out = u + w (b - A u)/d (+ dot b.out)
// MODE 2 RESTRICT (2x2x2 or 2x2x1): out[coarse] = sum_children (b - A u)
template <int MODE, bool DOT>
__global__ void __launch_bounds__(NT)
l0_ring_kernel(Grid g, const uint8_t* __restrict__ flags, const double* __restrict__ u,
               const double* __restrict__ b, double* __restrict__ out, double w, RingCoarse rc,
               int fz, int zchunk, double* red_partials, unsigned int* red_counter, double* red_out) {
    extern __shared__ __align__(16) double smem[];
    double* us = smem;                                          // [RING_R][U_STAGE]
    double* bs = us + RING_R * U_STAGE;                         // [RING_R][B_STAGE]   (MODE != 0)
    double* dtab = bs + (MODE != 0 ? RING_R * B_STAGE : 0);     // [64] diagonal, [64] its inverse
    unsigned char* fs = reinterpret_cast<unsigned char*>(dtab + 128);   // [RING_R][TY][TX] connectivity bytes

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int tx2 = (warp & 1) * 16 + (lane & 15);              // x pair index inside the tile
    const int ty = (warp >> 1) * 2 + (lane >> 4);
    const int i = blockIdx.x * TX + 2 * tx2;                    // first of the two cells (even)
    const int j = blockIdx.y * TY + ty;
    const bool inb = (i < g.nx) && (j < g.ny);                  // nx even: both cells in or out
    const int k0 = blockIdx.z * zchunk;
    const int k1 = min(k0 + zchunk, g.nz);

    if (tid < 64) {
        const double d = g.cx * (double)__popc(tid & 0x03) + g.cy * (double)__popc(tid & 0x0c) +
                         g.cz * (double)__popc(tid & 0x30);
        dtab[tid] = d;
        dtab[64 + tid] = d > 0.0 ? 1.0 / d : 0.0;
    }

    // halo duties
    const bool hw = (tx2 == 0), he = (tx2 == TX / 2 - 1);
    const bool hs = (ty == 0), hn = (ty == TY - 1);
    const int iw = i - 1, ie = i + 2, js = j - 1, jn = j + 1;
    const bool hw_ok = hw && (j < g.ny) && iw >= 0;
    const bool he_ok = he && (j < g.ny) && ie < g.nx;
    const bool hs_ok = hs && (i < g.nx) && js >= 0;
    const bool hn_ok = hn && (i < g.nx) && jn < g.ny;

    const long long col = inb ? (long long)j * g.nx + i : 0;
    const double* u_own = u + col;
    const double* u_w = u + (hw_ok ? (long long)j * g.nx + iw : 0);
    const double* u_e = u + (he_ok ? (long long)j * g.nx + ie : 0);
    const double* u_s = u + (hs_ok ? (long long)js * g.nx + i : 0);
    const double* u_n = u + (hn_ok ? (long long)jn * g.nx + i : 0);
    const double* b_own = (MODE != 0) ? b + col : nullptr;

    const int c_off = (ty + 1) * PITCH + 2 + 2 * tx2;           // own centre pair inside a u stage
    const int b_off = ty * TX + 2 * tx2;
    const int f_off = ty * TX + 2 * tx2;                        // byte offset inside a flag stage

    // issue all copies of plane kk into stage st (offset kk*plane is carried by the caller)
    auto issue = [&](long long poff, int st, bool with_b) {
        double* S = us + st * U_STAGE;
        cp_async16(S + c_off, u_own + poff, inb);
        if (hw) cp_async8(S + (ty + 1) * PITCH + 1, u_w + poff, hw_ok);
        if (he) cp_async8(S + (ty + 1) * PITCH + 2 + TX, u_e + poff, he_ok);
        if (hs) cp_async16(S + 2 + 2 * tx2, u_s + poff, hs_ok);
        if (hn) cp_async16(S + (TY + 1) * PITCH + 2 + 2 * tx2, u_n + poff, hn_ok);
        if (MODE != 0) {
            if (with_b) cp_async16(bs + st * B_STAGE + b_off, b_own + poff, inb);
        }
        // connectivity bytes of 4 cells (this thread's pair and its odd neighbour's)
        if (with_b && (tx2 & 1) == 0) cp_async4(fs + st * (TX * TY) + f_off, flags + col + poff, inb);
    };

    // prologue: planes k0-1 .. k0+RING_P, one commit group per plane
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
    double2 vv[3];
    vv[0] = *reinterpret_cast<const double2*>(us + 0 * U_STAGE + c_off);
    vv[1] = *reinterpret_cast<const double2*>(us + 1 * U_STAGE + c_off);

    double dot_acc = 0.0, zpair = 0.0;
    double* out_own = out + col + (long long)k0 * g.plane;      // MODE 0/1

    // The plane loop is unrolled by the ring length, so that stage numbers, the
    // register window and the flag queue slot are compile-time constants: plane
    // kb+s always lives in stage (s+1) % RING_R.
    for (int kb = k0; kb < k1; kb += RING_R) {
#pragma unroll
        for (int s = 0; s < RING_R; ++s) {
            const int k = kb + s;
            if (k >= k1) break;                                 // CTA-uniform
            constexpr int R = RING_R;
            const int sc = (s + 1) % R, sp = (s + 2) % R;       // stages of planes k, k+1
            cp_async_wait<RING_P - 1>();                        // own copies of planes <= k+1 landed
            __syncthreads();                                    // ... and everybody else's

            const double* Sc = us + sc * U_STAGE;
            const double2 v_m = vv[s % 3], v_c = vv[(s + 1) % 3];
            const double2 v_p = *reinterpret_cast<const double2*>(us + sp * U_STAGE + c_off);
            vv[(s + 2) % 3] = v_p;
            const double xw = Sc[c_off - 1], xe = Sc[c_off + 2];
            const double2 ys = *reinterpret_cast<const double2*>(Sc + c_off - PITCH);
            const double2 yn = *reinterpret_cast<const double2*>(Sc + c_off + PITCH);
            double2 bb = make_double2(0.0, 0.0);
            if (MODE != 0) bb = *reinterpret_cast<const double2*>(bs + sc * B_STAGE + b_off);

            const unsigned int f2 = *reinterpret_cast<const unsigned short*>(fs + sc * (TX * TY) + f_off);

            const unsigned int f0 = f2 & 0xffu, f1 = f2 >> 8;
            // (A u) for the two cells
            const double au0 = dtab[f0 & 63u] * v_c.x - (g.cx * (xw + v_c.y) + g.cy * (ys.x + yn.x) + g.cz * (v_m.x + v_p.x));
            const double au1 = dtab[f1 & 63u] * v_c.y - (g.cx * (v_c.x + xe) + g.cy * (ys.y + yn.y) + g.cz * (v_m.y + v_p.y));
            double2 o = make_double2(0.0, 0.0);
            double res = 0.0;
            if (MODE == 0) {
                if (f0 & F_UNK) o.x = w * au0;
                if (f1 & F_UNK) o.y = w * au1;
                if (DOT) dot_acc += v_c.x * o.x + v_c.y * o.y;
            } else if (MODE == 1) {
                if (f0 & F_UNK) o.x = v_c.x + w * (bb.x - au0) * dtab[64 + (f0 & 63u)];
                if (f1 & F_UNK) o.y = v_c.y + w * (bb.y - au1) * dtab[64 + (f1 & 63u)];
                if (DOT) dot_acc += bb.x * o.x + bb.y * o.y;
            } else {
                if (f0 & F_UNK) res = bb.x - au0;
                if (f1 & F_UNK) res += bb.y - au1;
            }
            if (MODE != 2) {
                if (inb) *reinterpret_cast<double2*>(out_own) = o;
                out_own += g.plane;
            } else {
                // x pair already summed in-thread; y pair = lane ^ 16; z pair carried
                double sum = res + __shfl_xor_sync(0xffffffffu, res, 16);
                const int kg = g.z0 + k;
                bool flush = true;
                if (fz == 2) {
                    if ((kg & 1) == 0) { zpair = sum; flush = (k + 1 == k1); }
                    else { sum += zpair; zpair = 0.0; }
                }
                if (flush && inb && ((j & 1) == 0)) {
                    const int ck = (fz == 2) ? ((kg >> 1) - rc.z0h) : k;
                    out[((long long)ck * rc.cny + (j >> 1)) * rc.cnx + (i >> 1)] = sum;
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

    if (DOT) {
        double v[1] = {dot_acc};
        grid_reduce<1>(v, red_partials, red_counter, red_out);
    }
}

__global__ void __launch_bounds__(256)
prolong_add_kernel(Grid g, const uint8_t* __restrict__ flags, double* __restrict__ z,
                   const double* __restrict__ ec, int cnx, int cny, int fx, int fy, int fz) {
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

template <int MODE>
size_t ring_smem_bytes() {
    return sizeof(double) * (size_t)(RING_R * U_STAGE + (MODE != 0 ? RING_R * B_STAGE : 0) + 128) +
           (size_t)RING_R * TX * TY;
}

template <int MODE, bool DOT>
void launch(const L0Args& a, cudaStream_t st) {
    static bool configured = false;
    const size_t smem = ring_smem_bytes<MODE>();
    if (!configured) {
        cudaFuncSetAttribute(l0_ring_kernel<MODE, DOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    const int zc = pick_zchunk(a.g, a.n_sm);
    dim3 grid((a.g.nx + TX - 1) / TX, (a.g.ny + TY - 1) / TY, (a.g.nz + zc - 1) / zc);
    RingCoarse rc{a.cnx, a.cny, a.g.z0 >> 1};
    l0_ring_kernel<MODE, DOT><<<grid, NT, smem, st>>>(a.g, a.flags, a.u, a.b, a.out, a.w, rc, a.fz, zc,
                                                       a.red_partials, a.red_counter, a.red_out);
}

}  // namespace

bool ring_supported(const L0Args& a, int mode) {
    // 16-byte vector accesses need even nx (then every row and plane start is
    // 16-byte aligned: plane 0 of every Field is 256-byte aligned)
    // ... and the 4-byte copies of the connectivity bytes need nx % 4 == 0
    if (a.g.nx & 3) return false;
    if (mode == 2 && !(a.fx == 2 && a.fy == 2)) return false;
    return true;
}

void ring_launch(const L0Args& a, int mode, bool dot, cudaStream_t st) {
    if (mode == 0) { if (dot) launch<0, true>(a, st); else launch<0, false>(a, st); }
    else if (mode == 1) { if (dot) launch<1, true>(a, st); else launch<1, false>(a, st); }
    else launch<2, false>(a, st);
}

void l0_prolong_add(const L0Args& a, cudaStream_t st) {
    int gz = a.g.nz < 64 ? a.g.nz : 64;
    dim3 grid((a.g.nx + 63) / 64, (a.g.ny + 3) / 4, gz);
    prolong_add_kernel<<<grid, 256, 0, st>>>(a.g, a.flags, a.out, a.ec, a.cnx, a.cny, a.fx, a.fy, a.fz);
}

}  // namespace oi
