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
//   B) forms the final value of the plane below from those three planes of v.
// Input, rhs and flag planes arrive through cp.async rings exactly as in the single-sweep
// kernel.  Rim values are recomputed by the neighbouring CTAs (16 % more arithmetic), their
// loads mostly hit L2.  Every field is zero off the unknowns and outside the box, so no
// face selects are needed (see oi_level0_ring.cu).
//
// Restrictions (the caller falls back to two single sweeps otherwise): one z-slab and a
// non-periodic box (ghost planes are one deep), nx % 4 == 0, fp32 multigrid vectors.
//
// STATUS (round 1): correct (tests/test_gpu_parity.py::test_pair_kernel_matches_single_sweeps) but
// not yet faster -- 6.1 ms per pair at 1024^3 against 2 x 2.43 ms for two single sweeps.  It is
// latency-bound, not bandwidth-bound: 89 KB of shared memory per CTA leaves 2 CTAs = 16 warps per
// SM, there are two CTA barriers per plane, and all seven operands of both stages come from shared
// memory.  Opt-in with OI_PAIR=1; next steps are register-carried z columns (as in the ring
// kernel), 512-thread CTAs and a shallower rhs ring.
#include "oi_kernels.h"

namespace oi {

namespace {

constexpr int PTX = 64, PTY = 16;              // tile (cells)
constexpr int PW = PTX + 8;                    // smem row pitch: [4 left | 64 | 4 right] cells
constexpr int UH = PTY + 4;                    // rows of an input stage   (two-cell rim)
constexpr int VH = PTY + 2;                    // rows of a v / rhs / flag stage (one-cell rim)
constexpr int PP = 3;                          // planes in flight beyond the newest one needed
constexpr int PR = PP + 3;                     // ring stages
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

    // kk = plane whose intermediate is formed this trip; the final value of plane kk-1 follows
#pragma unroll 1
    for (int kk = k0 - 1; kk <= k1; ++kk) {
        cpa_wait<PP>();                    // planes <= kk+1 have landed (this thread's copies)
        __syncthreads();                   // ... everybody's; and the previous trip's reads are done

        // ---- A: v(kk) on the tile + rim
        float* V = vs + (((kk % 3) + 3) % 3) * V_ST;
        const bool plane_in = (kk >= 0 && kk < g.nz);
        {
            const float* Um = us + stage_of(kk - 1) * U_ST;
            const float* Uc = us + stage_of(kk) * U_ST;
            const float* Up = us + stage_of(kk + 1) * U_ST;
            const float* B = bs + stage_of(kk) * V_ST;
            const unsigned char* F = fs + stage_of(kk) * V_ST;
            // a group at stage row r (0..17 in v coordinates), group gq (0..17); `only` < 0: all four
            // cells, else just that cell of the group (the rim needs one column of a side group)
            auto do_group = [&](int r, int gq, int only) {
                const int vo = r * PW + 4 * gq;              // v / rhs / flag offset
                const int uo = (r + 1) * PW + 4 * gq;        // same cell in an input stage
                if (!plane_in) {
                    if (only < 0) *reinterpret_cast<float4*>(V + vo) = make_float4(0.f, 0.f, 0.f, 0.f);
                    else V[vo + only] = 0.f;
                    return;
                }
                const unsigned int fw = *reinterpret_cast<const unsigned int*>(F + vo);
                if (only < 0) {
                    const float4 c = *reinterpret_cast<const float4*>(Uc + uo);
                    const float4 s = *reinterpret_cast<const float4*>(Uc + uo - PW);
                    const float4 n = *reinterpret_cast<const float4*>(Uc + uo + PW);
                    const float4 d = *reinterpret_cast<const float4*>(Um + uo);
                    const float4 p = *reinterpret_cast<const float4*>(Up + uo);
                    const float4 bb = *reinterpret_cast<const float4*>(B + vo);
                    const float xw = Uc[uo - 1], xe = Uc[uo + 4];
                    float4 o;
                    o.x = relax(fw & 0xffu, c.x, xw, c.y, s.x, n.x, d.x, p.x, bb.x, w1, cx, cy, cz, dtab);
                    o.y = relax((fw >> 8) & 0xffu, c.y, c.x, c.z, s.y, n.y, d.y, p.y, bb.y, w1, cx, cy, cz, dtab);
                    o.z = relax((fw >> 16) & 0xffu, c.z, c.y, c.w, s.z, n.z, d.z, p.z, bb.z, w1, cx, cy, cz, dtab);
                    o.w = relax(fw >> 24, c.w, c.z, xe, s.w, n.w, d.w, p.w, bb.w, w1, cx, cy, cz, dtab);
                    *reinterpret_cast<float4*>(V + vo) = o;
                } else {
                    const int a = uo + only;
                    V[vo + only] = relax((fw >> (8 * only)) & 0xffu, Uc[a], Uc[a - 1], Uc[a + 1], Uc[a - PW], Uc[a + PW],
                                         Um[a], Up[a], B[vo + only], w1, cx, cy, cz, dtab);
                }
            };
            do_group(ty + 1, 1 + tx, -1);
            if (vr2 >= 0) do_group(vr2, vg2, vg2 == 0 ? 3 : (vg2 == 17 ? 0 : -1));
        }
        __syncthreads();

        // ---- B: final value of plane kk-1 on the tile
        const int k = kk - 1;
        if (k >= k0 && k < k1) {
            const float* Vm = vs + ((((k - 1) % 3) + 3) % 3) * V_ST;
            const float* Vc = vs + (((k % 3) + 3) % 3) * V_ST;
            const float* Vp = V;
            const float* B = bs + stage_of(k) * V_ST;
            const unsigned char* F = fs + stage_of(k) * V_ST;
            const int vo = v_off1;
            const unsigned int fw = *reinterpret_cast<const unsigned int*>(F + vo);
            const float4 c = *reinterpret_cast<const float4*>(Vc + vo);
            const float4 s = *reinterpret_cast<const float4*>(Vc + vo - PW);
            const float4 n = *reinterpret_cast<const float4*>(Vc + vo + PW);
            const float4 d = *reinterpret_cast<const float4*>(Vm + vo);
            const float4 p = *reinterpret_cast<const float4*>(Vp + vo);
            const float4 bb = *reinterpret_cast<const float4*>(B + vo);
            const float xw = Vc[vo - 1], xe = Vc[vo + 4];
            float4 o;
            o.x = relax(fw & 0xffu, c.x, xw, c.y, s.x, n.x, d.x, p.x, bb.x, w2, cx, cy, cz, dtab);
            o.y = relax((fw >> 8) & 0xffu, c.y, c.x, c.z, s.y, n.y, d.y, p.y, bb.y, w2, cx, cy, cz, dtab);
            o.z = relax((fw >> 16) & 0xffu, c.z, c.y, c.w, s.z, n.z, d.z, p.z, bb.z, w2, cx, cy, cz, dtab);
            o.w = relax(fw >> 24, c.w, c.z, xe, s.w, n.w, d.w, p.w, bb.w, w2, cx, cy, cz, dtab);
            if (DOT) dot_acc += (double)bb.x * (double)o.x + (double)bb.y * (double)o.y + (double)bb.z * (double)o.z +
                                (double)bb.w * (double)o.w;
            // a group without an unknown stays zero: its sector is never written
            if (inb && (fw & 0x40404040u)) *reinterpret_cast<float4*>(out_own + (long long)k * g.plane) = o;
        }

        // refill: plane kk+2+PP takes the stage of plane kk-1, the oldest one.  Its input was last read
        // in A above (before the barrier); of its rhs / flags B read only this thread's own group,
        // and the other groups copied here are rim groups, which B never reads.
        issue(kk + 2 + PP);
        cpa_commit();
    }
    cpa_wait<0>();

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

// out = S_w2(S_w1(u)) ; dot: also red_out = b . out
void l0_smooth_pair(const L0Args& a, double w1, double w2, bool dot, cudaStream_t st) {
    static bool configured = false;
    const size_t smem = pair_smem_bytes();
    if (!configured) {
        cudaFuncSetAttribute(l0_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(l0_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    const int zc = pick_zchunk(a.g, a.n_sm);
    dim3 grid((a.g.nx + PTX - 1) / PTX, (a.g.ny + PTY - 1) / PTY, (a.g.nz + zc - 1) / zc);
    if (dot)
        l0_pair_kernel<true><<<grid, 256, smem, st>>>(a.g, a.flags, static_cast<const float*>(a.u),
            static_cast<const float*>(a.b), static_cast<float*>(a.out), (float)w1, (float)w2, zc, a.red_partials,
            a.red_counter, a.red_out);
    else
        l0_pair_kernel<false><<<grid, 256, smem, st>>>(a.g, a.flags, static_cast<const float*>(a.u),
            static_cast<const float*>(a.b), static_cast<float*>(a.out), (float)w1, (float)w2, zc, a.red_partials,
            a.red_counter, a.red_out);
}

}  // namespace oi
