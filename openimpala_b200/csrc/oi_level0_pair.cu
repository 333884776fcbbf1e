// Level-0 smoother, two sweeps per pass (temporal blocking of the weighted-Jacobi pair).
//
// The single-sweep ring kernel (oi_level0_ring.cu) runs at ~90 % of the HBM roofline, and
// level-0 smoothing is 43 % of a PCG iteration: the only way left to make it cheaper is to
// move fewer bytes.  This kernel applies
//     v = u + w1 (b - A u)/d ,   out = v + w2 (b - A v)/d
// in ONE pass over u, b and the connectivity bytes: 13 B per cell for two sweeps instead
// of 26.  A CTA owns a 64 x 16 tile and marches along z.  Per plane it
//   A) forms the intermediate v on the tile plus a one-cell rim (66 x 18 cells) from three
//      input planes staged with a two-cell rim (68 x 20), and keeps the last three planes
//      of v in shared memory;
//   B) forms the final value of the plane two below from the last three planes of v (its own
//      column travels in registers, only the x / y neighbours come from shared memory), so one
//      CTA barrier per plane suffices.
// Input, rhs and flag planes arrive through cp.async rings exactly as in the single-sweep
// kernel.  Rim values are recomputed by the neighbouring CTAs (16 % more arithmetic), their
// loads mostly hit L2.  Every field is zero off the unknowns and outside the box, so no
// face selects are needed (see oi_level0_ring.cu).
//
// Restrictions (the caller falls back to two single sweeps otherwise): one z-slab and a
// non-periodic box (ghost planes are one deep), nx % 4 == 0, fp32 multigrid vectors.
//
// STATUS (round 1), time per pair at 1024^3 against 2 x 2.43 ms for two single sweeps:
//   v1  256 threads, two barriers per plane, every operand from shared memory      6.3 ms
//   v2  256 threads, one barrier per plane, own columns in registers (OI_PAIR=1)   5.7 ms
//   v3  512 threads x 2 cells, 32 warps per SM                                      5.0 ms
//   v4  v3 + packed fp32 arithmetic (FFMA2 / FADD2 / FMUL2)  (default, OI_PAIR=2)  4.45 ms
// The kernel halves the DRAM bytes but is instruction bound (two stencil evaluations plus the
// 16 % rim per cell per pass), hence the packed arithmetic.  Next steps: rim work spread over
// all warps, packed arithmetic for the rim, a shallower rhs / flag ring.
#include "oi_kernels.h"

namespace oi {

namespace {

constexpr int PTX = 64, PTY = 16;              // tile (cells)
constexpr int PW = PTX + 8;                    // smem row pitch: [4 left | 64 | 4 right] cells
constexpr int UH = PTY + 4;                    // rows of an input stage   (two-cell rim)
constexpr int VH = PTY + 2;                    // rows of a v / rhs / flag stage (one-cell rim)
constexpr int PP = 2;                          // planes in flight beyond the newest one needed
constexpr int PR = PP + 4;                     // ring stages (one spare: see the refill comment)
constexpr int U_ST = UH * PW, V_ST = VH * PW;  // elements per stage

__device__ __forceinline__ void cpa16(void* s, const void* g, bool ok) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(s);
    const int sz = ok ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(g), "r"(sz) : "memory");
}
__device__ __forceinline__ void cpa4(void* s, const void* g, bool ok) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(s);
    const int sz = ok ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(sa), "l"(g), "r"(sz) : "memory");
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cpa_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// one weighted-Jacobi update of a cell from its seven values
__device__ __forceinline__ float relax(unsigned int f, float c, float w_, float e_, float s_, float n_, float d_,
                                       float u_, float bb, float w, float cx, float cy, float cz,
                                       const float* dtab) {
    const float au = dtab[f & 63u] * c - (cx * (w_ + e_) + cy * (s_ + n_) + cz * (d_ + u_));
    return (f & F_UNK) ? c + w * (bb - au) * dtab[64 + (f & 63u)] : 0.f;
}

template <bool DOT>
__global__ void __launch_bounds__(256)
l0_pair_kernel(Grid g, const uint8_t* __restrict__ flags, const float* __restrict__ u,
               const float* __restrict__ b, float* __restrict__ out, float w1, float w2, int zchunk,
               double* red_partials, unsigned int* red_counter, double* red_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* us = reinterpret_cast<float*>(smem_raw);            // [PR][UH][PW]
    float* bs = us + PR * U_ST;                                // [PR][VH][PW]
    float* vs = bs + PR * V_ST;                                // [3][VH][PW]
    float* dtab = vs + 3 * V_ST;                               // [64] diagonal, [64] inverse
    unsigned char* fs = reinterpret_cast<unsigned char*>(dtab + 128);   // [PR][VH][PW] bytes

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int i0 = blockIdx.x * PTX, j0 = blockIdx.y * PTY;
    const int k0 = blockIdx.z * zchunk, k1 = min(k0 + zchunk, g.nz);
    const float cx = (float)g.cx, cy = (float)g.cy, cz = (float)g.cz;
    if (tid < 64) {
        const float d = row_diag<float>((unsigned int)tid, g);
        dtab[tid] = d;
        dtab[64 + tid] = d > 0.f ? 1.f / d : 0.f;
    }

    // ---- copy duties.  A stage row is 18 groups of 4 cells: group 0 = cells i0-4..i0-1,
    // groups 1..16 the tile, group 17 = cells i0+64..i0+67.
    // input plane: 20 rows (j0-2 .. j0+17): own group (rows 0..15 of the tile) + one of the 104
    // remaining groups for tid < 104 (4 extra rows x 16 groups, then 20 rows x 2 side groups)
    int ur2 = -1, ug2 = 0;                 // second duty: stage row (0..19), group (0..17)
    if (tid < 64) { const int e = tid >> 4; ur2 = (e < 2) ? e : UH - 4 + e; ug2 = 1 + (tid & 15); }
    else if (tid < 104) { const int t = tid - 64; ur2 = t % UH; ug2 = (t / UH) ? 17 : 0; }
    // rhs / flag plane: 18 rows (j0-1 .. j0+16): own group + one of 68 remaining for tid < 68
    int vr2 = -1, vg2 = 0;
    if (tid < 32) { vr2 = (tid >> 4) ? VH - 1 : 0; vg2 = 1 + (tid & 15); }
    else if (tid < 68) { const int t = tid - 32; vr2 = t % VH; vg2 = (t / VH) ? 17 : 0; }

    auto in_box = [&](int gi, int gj) { return gi >= 0 && gi < g.nx && gj >= 0 && gj < g.ny; };
    // global (i, j) of a stage position
    const int u_i1 = i0 + 4 * tx, u_j1 = j0 + ty;                           // own group
    const bool u_ok1 = in_box(u_i1, u_j1);
    const long long u_col1 = u_ok1 ? (long long)u_j1 * g.nx + u_i1 : 0;
    const int u_i2 = i0 + 4 * (ug2 - 1), u_j2 = j0 - 2 + ur2;
    const bool u_ok2 = (ur2 >= 0) && in_box(u_i2, u_j2);
    const long long u_col2 = u_ok2 ? (long long)u_j2 * g.nx + u_i2 : 0;
    const int v_i2 = i0 + 4 * (vg2 - 1), v_j2 = j0 - 1 + vr2;
    const bool v_ok2 = (vr2 >= 0) && in_box(v_i2, v_j2);
    const long long v_col2 = v_ok2 ? (long long)v_j2 * g.nx + v_i2 : 0;
    const int u_off1 = (ty + 2) * PW + 4 + 4 * tx, u_off2 = (ur2 >= 0 ? ur2 : 0) * PW + 4 * ug2;
    const int v_off1 = (ty + 1) * PW + 4 + 4 * tx, v_off2 = (vr2 >= 0 ? vr2 : 0) * PW + 4 * vg2;

    // all copies of plane q (input, rhs, flags) into the stage of q; planes outside [-1, nz] are skipped
    auto issue = [&](int q) {
        if (q < -1 || q > g.nz) return;
        const int st = ((q - (k0 - 2)) % PR + PR) % PR;
        const long long poff = (long long)q * g.plane;
        float* U = us + st * U_ST;
        cpa16(U + u_off1, u + poff + u_col1, u_ok1);
        if (ur2 >= 0) cpa16(U + u_off2, u + poff + u_col2, u_ok2);
        float* B = bs + st * V_ST;
        unsigned char* F = fs + st * V_ST;
        cpa16(B + v_off1, b + poff + u_col1, u_ok1);
        cpa4(F + v_off1, flags + poff + u_col1, u_ok1);
        if (vr2 >= 0) {
            cpa16(B + v_off2, b + poff + v_col2, v_ok2);
            cpa4(F + v_off2, flags + poff + v_col2, v_ok2);
        }
    };
    auto stage_of = [&](int q) { return ((q - (k0 - 2)) % PR + PR) % PR; };

    // prologue: planes k0-2 .. k0+PP, one commit group per plane
#pragma unroll 1
    for (int q = k0 - 2; q <= k0 + PP; ++q) { issue(q); cpa_commit(); }

    const bool inb = u_ok1;
    float* out_own = out + u_col1;
    double dot_acc = 0.0;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    // own-group values that travel in registers instead of through shared memory:
    //   input column: planes kk-1, kk (the plane kk+1 is read once per trip)
    //   intermediate: v(kk-3), v(kk-2), v(kk-1) as produced by stage A of the last three trips
    //   rhs / flags of planes kk-2, kk-1 (read once, in stage A)
    float4 u_m = zero4, u_c = zero4;
    float4 v3 = zero4, v2 = zero4, v1 = zero4;
    float4 b2 = zero4, b1 = zero4;
    unsigned int f2 = 0u, f1 = 0u;
    cpa_wait<PP + 1>();                    // planes k0-2, k0-1 have landed
    __syncthreads();
    if (k0 - 2 >= -1) u_m = *reinterpret_cast<const float4*>(us + stage_of(k0 - 2) * U_ST + u_off1);
    u_c = *reinterpret_cast<const float4*>(us + stage_of(k0 - 1) * U_ST + u_off1);

    // Trip kk: stage A forms v(kk) (tile + rim, written to the v ring), stage B the final value of
    // plane kk-2 from v(kk-3 .. kk-1).  B only reads rim values that were written at least one trip
    // (= one barrier) earlier, so one barrier per plane is enough and A and B overlap freely.
#pragma unroll 1
    for (int kk = k0 - 1; kk <= k1 + 1; ++kk) {
        cpa_wait<PP>();                    // planes <= kk+1 have landed (this thread's copies)
        __syncthreads();                   // ... everybody's; v(kk-1) is complete; last trip's reads are done

        // ---- A: v(kk) on the tile + rim
        float* V = vs + (((kk % 3) + 3) % 3) * V_ST;
        const bool plane_in = (kk >= 0 && kk < g.nz);
        float4 oA = zero4, bA = zero4;
        unsigned int fA = 0u;
        float4 u_p = zero4;
        if (kk <= k1) {
            const float* Um = us + stage_of(kk - 1) * U_ST;
            const float* Uc = us + stage_of(kk) * U_ST;
            const float* Up = us + stage_of(kk + 1) * U_ST;
            const float* B = bs + stage_of(kk) * V_ST;
            const unsigned char* F = fs + stage_of(kk) * V_ST;
            if (kk + 1 <= g.nz) u_p = *reinterpret_cast<const float4*>(Up + u_off1);
            if (plane_in) {
                // own group: z neighbours from registers
                const int vo = v_off1, uo = u_off1;
                fA = *reinterpret_cast<const unsigned int*>(F + vo);
                bA = *reinterpret_cast<const float4*>(B + vo);
                const float4 sS = *reinterpret_cast<const float4*>(Uc + uo - PW);
                const float4 nN = *reinterpret_cast<const float4*>(Uc + uo + PW);
                const float xw = Uc[uo - 1], xe = Uc[uo + 4];
                oA.x = relax(fA & 0xffu, u_c.x, xw, u_c.y, sS.x, nN.x, u_m.x, u_p.x, bA.x, w1, cx, cy, cz, dtab);
                oA.y = relax((fA >> 8) & 0xffu, u_c.y, u_c.x, u_c.z, sS.y, nN.y, u_m.y, u_p.y, bA.y, w1, cx, cy, cz, dtab);
                oA.z = relax((fA >> 16) & 0xffu, u_c.z, u_c.y, u_c.w, sS.z, nN.z, u_m.z, u_p.z, bA.z, w1, cx, cy, cz, dtab);
                oA.w = relax(fA >> 24, u_c.w, u_c.z, xe, sS.w, nN.w, u_m.w, u_p.w, bA.w, w1, cx, cy, cz, dtab);
            }
            *reinterpret_cast<float4*>(V + v_off1) = oA;
            if (vr2 >= 0) {                // rim group (everything from shared memory)
                const int only = vg2 == 0 ? 3 : (vg2 == 17 ? 0 : -1);
                const int vo = vr2 * PW + 4 * vg2, uo = (vr2 + 1) * PW + 4 * vg2;
                if (!plane_in) {
                    if (only < 0) *reinterpret_cast<float4*>(V + vo) = zero4;
                    else V[vo + only] = 0.f;
                } else {
                    const unsigned int fw = *reinterpret_cast<const unsigned int*>(F + vo);
                    if (only < 0) {
                        const float4 c = *reinterpret_cast<const float4*>(Uc + uo);
                        const float4 sS = *reinterpret_cast<const float4*>(Uc + uo - PW);
                        const float4 nN = *reinterpret_cast<const float4*>(Uc + uo + PW);
                        const float4 d = *reinterpret_cast<const float4*>(Um + uo);
                        const float4 pp = *reinterpret_cast<const float4*>(Up + uo);
                        const float4 bb = *reinterpret_cast<const float4*>(B + vo);
                        const float xw = Uc[uo - 1], xe = Uc[uo + 4];
                        float4 o;
                        o.x = relax(fw & 0xffu, c.x, xw, c.y, sS.x, nN.x, d.x, pp.x, bb.x, w1, cx, cy, cz, dtab);
                        o.y = relax((fw >> 8) & 0xffu, c.y, c.x, c.z, sS.y, nN.y, d.y, pp.y, bb.y, w1, cx, cy, cz, dtab);
                        o.z = relax((fw >> 16) & 0xffu, c.z, c.y, c.w, sS.z, nN.z, d.z, pp.z, bb.z, w1, cx, cy, cz, dtab);
                        o.w = relax(fw >> 24, c.w, c.z, xe, sS.w, nN.w, d.w, pp.w, bb.w, w1, cx, cy, cz, dtab);
                        *reinterpret_cast<float4*>(V + vo) = o;
                    } else {
                        const int a = uo + only;
                        V[vo + only] = relax((fw >> (8 * only)) & 0xffu, Uc[a], Uc[a - 1], Uc[a + 1], Uc[a - PW],
                                             Uc[a + PW], Um[a], Up[a], B[vo + only], w1, cx, cy, cz, dtab);
                    }
                }
            }
        }

        // ---- B: final value of plane kk-2 on the tile: z neighbours and centre from registers,
        // x / y neighbours from the v stage of plane kk-2 (written two trips ago)
        const int k = kk - 2;
        if (k >= k0 && k < k1) {
            const float* Vc = vs + (((k % 3) + 3) % 3) * V_ST;
            const int vo = v_off1;
            const float4 sS = *reinterpret_cast<const float4*>(Vc + vo - PW);
            const float4 nN = *reinterpret_cast<const float4*>(Vc + vo + PW);
            const float xw = Vc[vo - 1], xe = Vc[vo + 4];
            float4 o;
            o.x = relax(f2 & 0xffu, v2.x, xw, v2.y, sS.x, nN.x, v3.x, v1.x, b2.x, w2, cx, cy, cz, dtab);
            o.y = relax((f2 >> 8) & 0xffu, v2.y, v2.x, v2.z, sS.y, nN.y, v3.y, v1.y, b2.y, w2, cx, cy, cz, dtab);
            o.z = relax((f2 >> 16) & 0xffu, v2.z, v2.y, v2.w, sS.z, nN.z, v3.z, v1.z, b2.z, w2, cx, cy, cz, dtab);
            o.w = relax(f2 >> 24, v2.w, v2.z, xe, sS.w, nN.w, v3.w, v1.w, b2.w, w2, cx, cy, cz, dtab);
            if (DOT) dot_acc += (double)b2.x * (double)o.x + (double)b2.y * (double)o.y + (double)b2.z * (double)o.z +
                                (double)b2.w * (double)o.w;
            // a group without an unknown stays zero: its sector is never written
            if (inb && (f2 & 0x40404040u)) *reinterpret_cast<float4*>(out_own + (long long)k * g.plane) = o;
        }
        // rotate the register queues
        u_m = u_c; u_c = u_p;
        v3 = v2; v2 = v1; v1 = oA;
        b2 = b1; b1 = bA;
        f2 = f1; f1 = fA;

        // refill: plane kk+2+PP takes the stage of plane kk-2 (PR = PP + 4 stages).  Nobody reads
        // that plane any more: its input was last used by stage A of the previous trip, which every
        // thread left before this trip's barrier; slower threads still in this trip need kk-1 .. kk+1.
        issue(kk + 2 + PP);
        cpa_commit();
    }
    cpa_wait<0>();

    if (DOT) {
        double v[1] = {dot_acc};
        grid_reduce<1>(v, red_partials, red_counter, red_out);
    }
}

// ---------------------------------------------------------------------------------------
// Variant with 512 threads per CTA (2 cells per thread): same tile, rings and trip structure,
// twice the warps per SM (2 CTAs x 16 warps) to hide the shared-memory and barrier latency
// that bounds the 256-thread version.  Copy duties are one 16-byte group per thread.
__device__ __forceinline__ float relax2(unsigned int f, float c, float w_, float e_, float s_, float n_, float d_,
                                        float u_, float bb, float w, float cx, float cy, float cz,
                                        const float2* dtab2) {
    const float2 dd = dtab2[f & 63u];
    const float au = dd.x * c - (cx * (w_ + e_) + cy * (s_ + n_) + cz * (d_ + u_));
    return (f & F_UNK) ? c + w * (bb - au) * dd.y : 0.f;
}

// Packed fp32 arithmetic (sm_100: one FFMA2 / FADD2 / FMUL2 per two lanes): the two cells of a
// thread's pair go through the stencil update together.
__device__ __forceinline__ unsigned long long pk(float2 v) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(v.x), "f"(v.y));
    return r;
}
__device__ __forceinline__ float2 upk(unsigned long long r) {
    float2 v;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(v.x), "=f"(v.y) : "l"(r));
    return v;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk(a)), "l"(pk(b)), "l"(pk(c)));
    return upk(r);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a)), "l"(pk(b)));
    return upk(r);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a)), "l"(pk(b)));
    return upk(r);
}
// weighted-Jacobi update of a pair of x-adjacent cells; nc* = (-c, -c) coefficient pairs
__device__ __forceinline__ float2 relax_pair(unsigned int f0, unsigned int f1, float2 c, float xw, float xe, float2 s_,
                                             float2 n_, float2 d_, float2 u_, float2 bb, float w, float2 ncx,
                                             float2 ncy, float2 ncz, const float2* dtab2) {
    const float2 dd0 = dtab2[f0 & 63u], dd1 = dtab2[f1 & 63u];
    float2 t = mul2(ncz, add2(d_, u_));
    t = fma2(ncy, add2(s_, n_), t);
    t = fma2(ncx, add2(make_float2(xw, c.x), make_float2(c.y, xe)), t);       // -(off-diagonal part)
    const float2 au = fma2(make_float2(dd0.x, dd1.x), c, t);
    const float2 res = fma2(au, make_float2(-1.f, -1.f), bb);
    float2 o = fma2(res, make_float2(w * dd0.y, w * dd1.y), c);
    o.x = (f0 & F_UNK) ? o.x : 0.f;
    o.y = (f1 & F_UNK) ? o.y : 0.f;
    return o;
}

template <bool DOT>
__global__ void __launch_bounds__(512, 2)
l0_pair512_kernel(Grid g, const uint8_t* __restrict__ flags, const float* __restrict__ u,
                  const float* __restrict__ b, float* __restrict__ out, float w1, float w2, int zchunk,
                  double* red_partials, unsigned int* red_counter, double* red_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* us = reinterpret_cast<float*>(smem_raw);            // [PR][UH][PW]
    float* bs = us + PR * U_ST;                                // [PR][VH][PW]
    float* vs = bs + PR * V_ST;                                // [3][VH][PW]
    float2* dtab2 = reinterpret_cast<float2*>(vs + 3 * V_ST);  // [64] {diagonal, inverse}
    unsigned char* fs = reinterpret_cast<unsigned char*>(dtab2 + 64);   // [PR][VH][PW] bytes

    const int tid = threadIdx.x;
    const int tx = tid & 31, ty = tid >> 5;                    // 32 pairs x 16 rows
    const int i0 = blockIdx.x * PTX, j0 = blockIdx.y * PTY;
    const int k0 = blockIdx.z * zchunk, k1 = min(k0 + zchunk, g.nz);
    const float cx = (float)g.cx, cy = (float)g.cy, cz = (float)g.cz;
    const float2 ncx = make_float2(-cx, -cx), ncy = make_float2(-cy, -cy), ncz = make_float2(-cz, -cz);
    if (tid < 64) {
        const float d = row_diag<float>((unsigned int)tid, g);
        dtab2[tid] = make_float2(d, d > 0.f ? 1.f / d : 0.f);
    }
    auto in_box = [&](int gi, int gj) { return gi >= 0 && gi < g.nx && gj >= 0 && gj < g.ny; };

    // copy duties: input plane 20 rows x 18 groups (tid < 360), rhs / flag plane 18 x 18 (tid < 324)
    const bool u_duty = tid < UH * 18, v_duty = tid < VH * 18;
    const int ur = tid / 18, ug = tid % 18;
    const int u_i = i0 + 4 * (ug - 1), u_j = j0 - 2 + ur;
    const bool u_ok = u_duty && in_box(u_i, u_j);
    const long long u_col = u_ok ? (long long)u_j * g.nx + u_i : 0;
    const int u_dst = ur * PW + 4 * ug;
    const int v_j = j0 - 1 + ur;                               // same (row, group) split, one row less rim
    const bool v_ok = v_duty && in_box(u_i, v_j);
    const long long v_col = v_ok ? (long long)v_j * g.nx + u_i : 0;
    const int v_dst = ur * PW + 4 * ug;

    // own pair
    const int oi = i0 + 2 * tx, oj = j0 + ty;
    const bool inb = in_box(oi, oj);
    const long long o_col = inb ? (long long)oj * g.nx + oi : 0;
    const int u_off = (ty + 2) * PW + 4 + 2 * tx;
    const int v_off = (ty + 1) * PW + 4 + 2 * tx;
    // rim duty: rows 0 and 17 as 64 pairs (tid < 64), the two side columns as 36 single cells (tid 64..99)
    int rim_kind = 0, rim_vo = 0;                              // 1 = pair, 2 = single cell
    if (tid < 64) { rim_kind = 1; rim_vo = ((tid >> 5) ? VH - 1 : 0) * PW + 4 + 2 * (tid & 31); }
    else if (tid < 100) { const int t = tid - 64; rim_kind = 2; rim_vo = (t % VH) * PW + ((t / VH) ? 4 + PTX : 3); }

    auto stage_of = [&](int q) { return ((q - (k0 - 2)) % PR + PR) % PR; };
    auto issue = [&](int q) {
        if (q < -1 || q > g.nz) return;
        const int st = stage_of(q);
        const long long poff = (long long)q * g.plane;
        if (u_duty) cpa16(us + st * U_ST + u_dst, u + poff + u_col, u_ok);
        if (v_duty) {
            cpa16(bs + st * V_ST + v_dst, b + poff + v_col, v_ok);
            cpa4(fs + st * V_ST + v_dst, flags + poff + v_col, v_ok);
        }
    };
#pragma unroll 1
    for (int q = k0 - 2; q <= k0 + PP; ++q) { issue(q); cpa_commit(); }

    float* out_own = out + o_col;
    double dot_acc = 0.0;
    const float2 zero2 = make_float2(0.f, 0.f);
    float2 u_m = zero2, u_c = zero2, v3 = zero2, v2 = zero2, v1 = zero2, b2 = zero2, b1 = zero2;
    unsigned int f2 = 0u, f1 = 0u;
    cpa_wait<PP + 1>();
    __syncthreads();
    if (k0 - 2 >= -1) u_m = *reinterpret_cast<const float2*>(us + stage_of(k0 - 2) * U_ST + u_off);
    u_c = *reinterpret_cast<const float2*>(us + stage_of(k0 - 1) * U_ST + u_off);

#pragma unroll 1
    for (int kk = k0 - 1; kk <= k1 + 1; ++kk) {
        cpa_wait<PP>();
        __syncthreads();

        float* V = vs + (((kk % 3) + 3) % 3) * V_ST;
        const bool plane_in = (kk >= 0 && kk < g.nz);
        float2 oA = zero2, bA = zero2, u_p = zero2;
        unsigned int fA = 0u;
        if (kk <= k1) {
            const float* Um = us + stage_of(kk - 1) * U_ST;
            const float* Uc = us + stage_of(kk) * U_ST;
            const float* Up = us + stage_of(kk + 1) * U_ST;
            const float* B = bs + stage_of(kk) * V_ST;
            const unsigned char* F = fs + stage_of(kk) * V_ST;
            if (kk + 1 <= g.nz) u_p = *reinterpret_cast<const float2*>(Up + u_off);
            if (plane_in) {
                fA = *reinterpret_cast<const unsigned short*>(F + v_off);
                bA = *reinterpret_cast<const float2*>(B + v_off);
                const float2 sS = *reinterpret_cast<const float2*>(Uc + u_off - PW);
                const float2 nN = *reinterpret_cast<const float2*>(Uc + u_off + PW);
                const float xw = Uc[u_off - 1], xe = Uc[u_off + 2];
                oA = relax_pair(fA & 0xffu, fA >> 8, u_c, xw, xe, sS, nN, u_m, u_p, bA, w1, ncx, ncy, ncz, dtab2);
            }
            *reinterpret_cast<float2*>(V + v_off) = oA;
            if (rim_kind) {
                const int vo = rim_vo, uo = rim_vo + PW;       // same cell one row further down in an input stage
                if (!plane_in) {
                    V[vo] = 0.f;
                    if (rim_kind == 1) V[vo + 1] = 0.f;
                } else {
                    V[vo] = relax2(F[vo], Uc[uo], Uc[uo - 1], Uc[uo + 1], Uc[uo - PW], Uc[uo + PW], Um[uo], Up[uo],
                                   B[vo], w1, cx, cy, cz, dtab2);
                    if (rim_kind == 1)
                        V[vo + 1] = relax2(F[vo + 1], Uc[uo + 1], Uc[uo], Uc[uo + 2], Uc[uo + 1 - PW], Uc[uo + 1 + PW],
                                           Um[uo + 1], Up[uo + 1], B[vo + 1], w1, cx, cy, cz, dtab2);
                }
            }
        }

        const int k = kk - 2;
        if (k >= k0 && k < k1) {
            const float* Vc = vs + (((k % 3) + 3) % 3) * V_ST;
            const float2 sS = *reinterpret_cast<const float2*>(Vc + v_off - PW);
            const float2 nN = *reinterpret_cast<const float2*>(Vc + v_off + PW);
            const float xw = Vc[v_off - 1], xe = Vc[v_off + 2];
            const float2 o = relax_pair(f2 & 0xffu, f2 >> 8, v2, xw, xe, sS, nN, v3, v1, b2, w2, ncx, ncy, ncz, dtab2);
            if (DOT) dot_acc += (double)b2.x * (double)o.x + (double)b2.y * (double)o.y;
            if (inb && (f2 & 0x4040u)) *reinterpret_cast<float2*>(out_own + (long long)k * g.plane) = o;
        }
        u_m = u_c; u_c = u_p;
        v3 = v2; v2 = v1; v1 = oA;
        b2 = b1; b1 = bA;
        f2 = f1; f1 = fA;

        issue(kk + 2 + PP);
        cpa_commit();
    }
    cpa_wait<0>();

    if (DOT) {
        double v[1] = {dot_acc};
        grid_reduce<1>(v, red_partials, red_counter, red_out);
    }
}

// ---------------------------------------------------------------------------------------
// Variant 3: the 512-thread kernel with the trip loop unrolled by the ring length, so that every
// stage offset (input ring of PR = 6 stages, intermediate ring of 3) is a compile-time constant.
// The SASS of variant 2 spends more than a third of its 309 instructions per trip on stage
// arithmetic (IMAD / LEA / ISETP from the runtime modulo of stage_of()); the kernel is issue bound
// (ncu: sm__inst_issued 77-80 %, DRAM 33 %), so those instructions are what it pays for.
// Trip kk = kb + s with kb = k0 - 1 + 6 t:  plane kk+d of the input lives in stage (s + d + 1) % 6,
// v(kk) in intermediate stage s % 3, v(kk-2) in (s + 1) % 3, the refill plane kk+4 goes to (s + 5) % 6.
// HALO (one z-slab of several): u carries the neighbours' boundary planes in its ghost planes, and the
// intermediate iterate at k = -1 / k = nz -- the neighbours' own v(nz-1) / v(0), made by the boundary-plane
// run of the ring kernel and stored into vb_lo / vb_hi -- is read instead of taken as zero.  The first and
// last z-chunk are dispatched last, wait for those planes on the flag words of `hin`, store plane 0 / nz-1
// of the result into the neighbours' ghost planes and publish `hout.seq`.
template <bool DOT, bool HALO>
__global__ void __launch_bounds__(512, 2)
l0_pair512u_kernel(Grid g, const uint8_t* __restrict__ flags, const float* __restrict__ u,
                   const float* __restrict__ b, float* __restrict__ out, float w1, float w2, int zchunk,
                   double* red_partials, unsigned int* red_counter, double* red_out, HaloIn hin, HaloOut hout,
                   const float* __restrict__ vb_lo, const float* __restrict__ vb_hi) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* us = reinterpret_cast<float*>(smem_raw);            // [PR][UH][PW]
    float* bs = us + PR * U_ST;                                // [PR][VH][PW]
    float* vs = bs + PR * V_ST;                                // [3][VH][PW]
    float2* dtab2 = reinterpret_cast<float2*>(vs + 3 * V_ST);  // [64] {diagonal, inverse}
    unsigned char* fs = reinterpret_cast<unsigned char*>(dtab2 + 64);   // [PR][VH][PW] bytes

    const int tid = threadIdx.x;
    const int tx = tid & 31, ty = tid >> 5;                    // 32 pairs x 16 rows
    const int i0 = blockIdx.x * PTX, j0 = blockIdx.y * PTY;
    int zc_idx = blockIdx.z;
    if (HALO) {                                                 // interior chunks first, chunk 0 and the last one at the end
        const int nzc = gridDim.z;
        if (nzc > 2) zc_idx = ((int)blockIdx.z < nzc - 2) ? (int)blockIdx.z + 1 : ((int)blockIdx.z == nzc - 2 ? 0 : nzc - 1);
    }
    const int k0 = zc_idx * zchunk, k1 = min(k0 + zchunk, g.nz);
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
    const float cx = (float)g.cx, cy = (float)g.cy, cz = (float)g.cz;
    const float2 ncx = make_float2(-cx, -cx), ncy = make_float2(-cy, -cy), ncz = make_float2(-cz, -cz);
    if (tid < 64) {
        const float d = row_diag<float>((unsigned int)tid, g);
        dtab2[tid] = make_float2(d, d > 0.f ? 1.f / d : 0.f);
    }
    auto in_box = [&](int gi, int gj) { return gi >= 0 && gi < g.nx && gj >= 0 && gj < g.ny; };

    const bool u_duty = tid < UH * 18, v_duty = tid < VH * 18;
    const int ur = tid / 18, ug = tid % 18;
    const int u_i = i0 + 4 * (ug - 1), u_j = j0 - 2 + ur;
    const bool u_ok = u_duty && in_box(u_i, u_j);
    const float* u_src = u + (u_ok ? (long long)u_j * g.nx + u_i : 0);
    const int u_dst = ur * PW + 4 * ug;
    const int v_j = j0 - 1 + ur;
    const bool v_ok = v_duty && in_box(u_i, v_j);
    const long long v_col = v_ok ? (long long)v_j * g.nx + u_i : 0;
    const float* b_src = b + v_col;
    const uint8_t* f_src = flags + v_col;
    const int v_dst = ur * PW + 4 * ug;

    const int oi = i0 + 2 * tx, oj = j0 + ty;
    const bool inb = in_box(oi, oj);
    const long long o_col = inb ? (long long)oj * g.nx + oi : 0;
    const int u_off = (ty + 2) * PW + 4 + 2 * tx;
    const int v_off = (ty + 1) * PW + 4 + 2 * tx;
    int rim_kind = 0, rim_vo = 0;                              // 1 = pair, 2 = single cell
    if (tid < 64) { rim_kind = 1; rim_vo = ((tid >> 5) ? VH - 1 : 0) * PW + 4 + 2 * (tid & 31); }
    else if (tid < 100) { const int t = tid - 64; rim_kind = 2; rim_vo = (t % VH) * PW + ((t / VH) ? 4 + PTX : 3); }

    // copies of plane q into stage st (a constant at every call site of the unrolled loop)
    auto issue = [&](int q, int st) {
        if (q < -1 || q > g.nz) return;
        const long long poff = (long long)q * g.plane;
        if (u_duty) cpa16(us + st * U_ST + u_dst, u_src + poff, u_ok);
        if (v_duty) {
            cpa16(bs + st * V_ST + v_dst, b_src + poff, v_ok);
            cpa4(fs + st * V_ST + v_dst, f_src + poff, v_ok);
        }
    };
    // prologue: planes k0-2 .. k0+PP into stages 0 .. PP+2, one commit group per plane
#pragma unroll
    for (int d = 0; d <= PP + 2; ++d) { issue(k0 - 2 + d, d); cpa_commit(); }

    float* out_own = out + o_col;
    double dot_acc = 0.0;
    const float2 zero2 = make_float2(0.f, 0.f);
    float2 u_m = zero2, u_c = zero2, v3 = zero2, v2 = zero2, v1 = zero2, b2 = zero2, b1 = zero2;
    unsigned int f2 = 0u, f1 = 0u;
    cpa_wait<PP + 1>();
    __syncthreads();
    if (k0 - 2 >= -1) u_m = *reinterpret_cast<const float2*>(us + 0 * U_ST + u_off);
    u_c = *reinterpret_cast<const float2*>(us + 1 * U_ST + u_off);

    for (int kb = k0 - 1; kb <= k1 + 1; kb += PR) {
#pragma unroll
        for (int s = 0; s < PR; ++s) {
            const int kk = kb + s;
            if (kk > k1 + 1) break;                            // CTA-uniform
            cpa_wait<PP>();
            __syncthreads();

            float* V = vs + (s % 3) * V_ST;
            const bool plane_in = (kk >= 0 && kk < g.nz);
            float2 oA = zero2, bA = zero2, u_p = zero2;
            unsigned int fA = 0u;
            if (kk <= k1) {
                const float* Um = us + (s % PR) * U_ST;
                const float* Uc = us + ((s + 1) % PR) * U_ST;
                const float* Up = us + ((s + 2) % PR) * U_ST;
                const float* B = bs + ((s + 1) % PR) * V_ST;
                const unsigned char* F = fs + ((s + 1) % PR) * V_ST;
                if (kk + 1 <= g.nz) u_p = *reinterpret_cast<const float2*>(Up + u_off);
                if (plane_in) {
                    fA = *reinterpret_cast<const unsigned short*>(F + v_off);
                    bA = *reinterpret_cast<const float2*>(B + v_off);
                    const float2 sS = *reinterpret_cast<const float2*>(Uc + u_off - PW);
                    const float2 nN = *reinterpret_cast<const float2*>(Uc + u_off + PW);
                    const float xw = Uc[u_off - 1], xe = Uc[u_off + 2];
                    oA = relax_pair(fA & 0xffu, fA >> 8, u_c, xw, xe, sS, nN, u_m, u_p, bA, w1, ncx, ncy, ncz, dtab2);
                }
                // the neighbours' intermediate plane stands in at k = -1 / k = nz (HALO); own values are loaded
                // here, rim values below
                const float* vbp = nullptr;
                if (HALO && !plane_in) vbp = (kk < 0) ? vb_lo : vb_hi;
                if (vbp && inb) oA = *reinterpret_cast<const float2*>(vbp + o_col);
                *reinterpret_cast<float2*>(V + v_off) = oA;
                if (rim_kind) {
                    const int vo = rim_vo, uo = rim_vo + PW;
                    if (!plane_in) {
                        float r0 = 0.f, r1 = 0.f;
                        if (vbp) {
                            // stage position vo -> cell (i0 - 4 + column, j0 - 1 + row) of the plane
                            const int rr = vo / PW, cc = vo - rr * PW;
                            const int gi = i0 - 4 + cc, gj = j0 - 1 + rr;
                            if (in_box(gi, gj)) r0 = vbp[(long long)gj * g.nx + gi];
                            if (rim_kind == 1 && in_box(gi + 1, gj)) r1 = vbp[(long long)gj * g.nx + gi + 1];
                        }
                        V[vo] = r0;
                        if (rim_kind == 1) V[vo + 1] = r1;
                    } else {
                        V[vo] = relax2(F[vo], Uc[uo], Uc[uo - 1], Uc[uo + 1], Uc[uo - PW], Uc[uo + PW], Um[uo], Up[uo],
                                       B[vo], w1, cx, cy, cz, dtab2);
                        if (rim_kind == 1)
                            V[vo + 1] = relax2(F[vo + 1], Uc[uo + 1], Uc[uo], Uc[uo + 2], Uc[uo + 1 - PW], Uc[uo + 1 + PW],
                                               Um[uo + 1], Up[uo + 1], B[vo + 1], w1, cx, cy, cz, dtab2);
                    }
                }
            }

            const int k = kk - 2;
            if (k >= k0 && k < k1) {
                const float* Vc = vs + ((s + 1) % 3) * V_ST;
                const float2 sS = *reinterpret_cast<const float2*>(Vc + v_off - PW);
                const float2 nN = *reinterpret_cast<const float2*>(Vc + v_off + PW);
                const float xw = Vc[v_off - 1], xe = Vc[v_off + 2];
                const float2 o = relax_pair(f2 & 0xffu, f2 >> 8, v2, xw, xe, sS, nN, v3, v1, b2, w2, ncx, ncy, ncz, dtab2);
                if (DOT) dot_acc += (double)b2.x * (double)o.x + (double)b2.y * (double)o.y;
                if (inb && (f2 & 0x4040u)) {
                    *reinterpret_cast<float2*>(out_own + (long long)k * g.plane) = o;
                    if (HALO) {
                        if (hout.dst_lo && k == 0) *reinterpret_cast<float2*>(static_cast<float*>(hout.dst_lo) + o_col) = o;
                        if (hout.dst_hi && k == g.nz - 1) *reinterpret_cast<float2*>(static_cast<float*>(hout.dst_hi) + o_col) = o;
                    }
                }
            }
            u_m = u_c; u_c = u_p;
            v3 = v2; v2 = v1; v1 = oA;
            b2 = b1; b1 = bA;
            f2 = f1; f1 = fA;

            issue(kk + 2 + PP, (s + 5) % PR);
            cpa_commit();
        }
    }
    cpa_wait<0>();

    if (HALO) {
        const unsigned int tiles = gridDim.x * gridDim.y;
        if (hout.flag_lo && k0 == 0) halo_publish(hout.counter + 0, tiles, hout.flag_lo, nullptr, hout.seq);
        if (hout.flag_hi && k1 == g.nz) halo_publish(hout.counter + 1, tiles, nullptr, hout.flag_hi, hout.seq);
    }

    if (DOT) {
        double v[1] = {dot_acc};
        grid_reduce<1>(v, red_partials, red_counter, red_out);
    }
}

size_t pair_smem_bytes() {
    return sizeof(float) * (size_t)(PR * U_ST + PR * V_ST + 3 * V_ST + 128) + (size_t)PR * V_ST;
}

}  // namespace

bool pair_supported(const L0Args& a) {
    return sizeof(mg_t) == 4 && (a.g.nx & 3) == 0 && a.g.periodic == 0 && a.g.nz == a.g.nzg;
}
bool pair_supported_slab(const L0Args& a) {
    return sizeof(mg_t) == 4 && (a.g.nx & 3) == 0 && a.g.periodic == 0 && a.g.nz >= 2;
}

// out = S_w2(S_w1(u)) ; dot: also red_out = b . out.  variant 1: 256 threads x 4 cells, 2: 512 x 2.
void l0_smooth_pair(const L0Args& a, double w1, double w2, bool dot, int variant, cudaStream_t st) {
    static unsigned long long configured = 0;
    const size_t smem = pair_smem_bytes();
    if (first_use_on_this_device(configured)) {
        cudaFuncSetAttribute(l0_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(l0_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(l0_pair512_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(l0_pair512_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(l0_pair512u_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(l0_pair512u_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(l0_pair512u_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(l0_pair512u_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    const int zc = pick_zchunk(a.g, a.n_sm);
    dim3 grid((a.g.nx + PTX - 1) / PTX, (a.g.ny + PTY - 1) / PTY, (a.g.nz + zc - 1) / zc);
    const float* u = static_cast<const float*>(a.u);
    const float* b = static_cast<const float*>(a.b);
    float* out = static_cast<float*>(a.out);
    const bool halo = a.hin.flag_lo || a.hin.flag_hi || a.hout.flag_lo || a.hout.flag_hi || a.vb_lo || a.vb_hi;
    const float* vlo = static_cast<const float*>(a.vb_lo);
    const float* vhi = static_cast<const float*>(a.vb_hi);
    if (halo) {          // one z-slab of several: always the unrolled 512-thread kernel
        if (dot) l0_pair512u_kernel<true, true><<<grid, 512, smem, st>>>(a.g, a.flags, u, b, out, (float)w1, (float)w2, zc,
                                                                         a.red_partials, a.red_counter, a.red_out, a.hin, a.hout, vlo, vhi);
        else l0_pair512u_kernel<false, true><<<grid, 512, smem, st>>>(a.g, a.flags, u, b, out, (float)w1, (float)w2, zc,
                                                                      a.red_partials, a.red_counter, a.red_out, a.hin, a.hout, vlo, vhi);
    } else if (variant == 3) {
        if (dot) l0_pair512u_kernel<true, false><<<grid, 512, smem, st>>>(a.g, a.flags, u, b, out, (float)w1, (float)w2, zc,
                                                                          a.red_partials, a.red_counter, a.red_out, a.hin, a.hout, nullptr, nullptr);
        else l0_pair512u_kernel<false, false><<<grid, 512, smem, st>>>(a.g, a.flags, u, b, out, (float)w1, (float)w2, zc,
                                                                       a.red_partials, a.red_counter, a.red_out, a.hin, a.hout, nullptr, nullptr);
    } else if (variant == 2) {
        if (dot) l0_pair512_kernel<true><<<grid, 512, smem, st>>>(a.g, a.flags, u, b, out, (float)w1, (float)w2, zc,
                                                                  a.red_partials, a.red_counter, a.red_out);
        else l0_pair512_kernel<false><<<grid, 512, smem, st>>>(a.g, a.flags, u, b, out, (float)w1, (float)w2, zc,
                                                               a.red_partials, a.red_counter, a.red_out);
    } else {
        if (dot) l0_pair_kernel<true><<<grid, 256, smem, st>>>(a.g, a.flags, u, b, out, (float)w1, (float)w2, zc,
                                                               a.red_partials, a.red_counter, a.red_out);
        else l0_pair_kernel<false><<<grid, 256, smem, st>>>(a.g, a.flags, u, b, out, (float)w1, (float)w2, zc,
                                                            a.red_partials, a.red_counter, a.red_out);
    }
}

}  // namespace oi
