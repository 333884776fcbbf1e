// Level-0 stencil, TMA variant of the shared-memory ring (oi_level0_ring.cu).
//
// Same tile, ring and arithmetic as the cp.async ring kernel, but every z-plane of a stage arrives
// as ONE bulk tensor copy per array (cp.async.bulk.tensor.3d, SASS UTMALDG) issued by a single
// elected thread and tracked by an mbarrier per stage:
//   * the input box is (TX + 2 CPT) x (TY + 2) x 1 cells starting at (i0 - CPT, j0 - 1, k): the tile, its
//     one-cell rim and the alignment pad in one descriptor-driven copy.  Cells outside the box of the
//     field are zero-filled by the TMA unit, which is exactly the boundary condition the operator needs
//     (every vector is zero off the unknowns and outside the domain), so the five per-thread halo
//     copies of the cp.async kernel, their predicates and their address arithmetic disappear;
//   * rhs (SMOOTH) and connectivity bytes ride in the same stage as two more boxes of TX x TY x 1;
//   * consumers wait on the stage's mbarrier (try_wait.parity) instead of cp.async.wait_group; the
//     one CTA barrier per plane stays (it is what makes the stage of plane k-1 free for the refill).
// The tensor maps are encoded once per field by the host (cuTensorMapEncodeTiled through the runtime's
// driver entry point) with the ghost plane below plane 0 as the origin, so plane k is coordinate k + 1
// and both ghost planes are ordinary coordinates.
//
// Scope: APPLY (fp64) and SMOOTH (fp32), non-periodic boxes, nx % 16 == 0 (the connectivity bytes'
// row pitch must be a multiple of 16 bytes for a tensor map), single z-slab or explicit halo exchange.
// Selected with OI_TMA=1 (A/B against the cp.async ring: profiles/r2_tma_ab.md).
#include <cuda.h>

#include "oi_kernels.h"

namespace oi {

namespace {

constexpr int TRING_P = 4;
constexpr int TRING_R = TRING_P + 2;

template <typename T>
struct TCfg {
    static constexpr int CPT = 16 / (int)sizeof(T);
    static constexpr int TX = 64;
    static constexpr int XT = TX / CPT;
    static constexpr int NT = 256;
    static constexpr int TY = NT / XT;
    static constexpr int PITCH = TX + 2 * CPT;
    static constexpr int U_ELEMS = (TY + 2) * PITCH;
    static constexpr int U_BYTES = U_ELEMS * (int)sizeof(T);
    static constexpr int U_STAGE_BYTES = (U_BYTES + 127) / 128 * 128;      // TMA destinations are 128-byte aligned
    static constexpr int B_BYTES = TY * TX * (int)sizeof(T);
    static constexpr int F_BYTES = TY * TX;
};

template <typename T>
struct alignas(16) TPack { T v[16 / sizeof(T)]; };

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(void* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
    const unsigned a = smem_u32(bar);
    unsigned done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(a), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, void* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// MODE 0 APPLY: out = w * A u (+ dot u.out) ; MODE 1 SMOOTH: out = u + w (b - A u)/d (+ dot b.out)
template <typename T, int MODE, bool DOT>
__global__ void __launch_bounds__(256)
l0_tma_kernel(Grid g, const __grid_constant__ CUtensorMap tm_u, const __grid_constant__ CUtensorMap tm_b,
              const __grid_constant__ CUtensorMap tm_f, T* __restrict__ out, T w, int zchunk,
              double* red_partials, unsigned int* red_counter, double* red_out) {
    typedef TCfg<T> C;
    constexpr int CPT = C::CPT, TX = C::TX, TY = C::TY, PITCH = C::PITCH;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* us_b = smem_raw;                                             // [R][U_STAGE_BYTES]
    unsigned char* bs_b = us_b + TRING_R * C::U_STAGE_BYTES;                    // [R][B_BYTES]   (MODE 1)
    unsigned char* fs = bs_b + (MODE != 0 ? TRING_R * C::B_BYTES : 0);          // [R][F_BYTES]
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(fs + TRING_R * C::F_BYTES);   // [R]
    T* dtab = reinterpret_cast<T*>(bars + 8);                                   // [64] diagonal, [64] its inverse

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int WX = C::XT / 16;
    const int tx = (warp % WX) * 16 + (lane & 15);
    const int ty = (warp / WX) * 2 + (lane >> 4);
    const int i0 = blockIdx.x * TX, j0 = blockIdx.y * TY;
    const int i = i0 + CPT * tx;
    const int j = j0 + ty;
    const bool inb = (i < g.nx) && (j < g.ny);
    const int k0 = blockIdx.z * zchunk;
    const int k1 = min(k0 + zchunk, g.nz);
    const T cx = (T)g.cx, cy = (T)g.cy, cz = (T)g.cz;

    if (tid < 64) {
        const T d = row_diag<T>((unsigned int)tid, g);
        dtab[tid] = d;
        dtab[64 + tid] = d > (T)0 ? (T)1 / d : (T)0;
    }
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < TRING_R; ++s) mbar_init(bars + s, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    const long long col = inb ? (long long)j * g.nx + i : 0;
    const int c_off = (ty + 1) * PITCH + CPT + CPT * tx;        // own centre group inside a field stage
    const int b_off = ty * TX + CPT * tx;                       // inside a rhs stage / flag stage

    // plane kk -> stage st: one thread arms the stage's mbarrier with the bytes to come and issues the boxes
    auto issue = [&](int kk, int st, bool interior) {
        constexpr unsigned full = (unsigned)C::U_BYTES + (MODE != 0 ? (unsigned)C::B_BYTES : 0u) + (unsigned)C::F_BYTES;
        mbar_expect_tx(bars + st, interior ? full : (unsigned)C::U_BYTES);
        tma_load_3d(us_b + st * C::U_STAGE_BYTES, &tm_u, bars + st, i0 - CPT, j0 - 1, kk + 1);
        if (interior) {
            if (MODE != 0) tma_load_3d(bs_b + st * C::B_BYTES, &tm_b, bars + st, i0, j0, kk + 1);
            tma_load_3d(fs + st * C::F_BYTES, &tm_f, bars + st, i0, j0, kk + 1);
        }
    };

    // prologue: planes k0-1 .. k0+TRING_P
    int kk_issue = k0 - 1;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < TRING_R; ++s)
            if (kk_issue + s <= k1) issue(kk_issue + s, s, kk_issue + s >= k0 && kk_issue + s < k1);
    }
    kk_issue += TRING_R;

    mbar_wait(bars + 0, 0u);                                    // plane k0-1
    mbar_wait(bars + 1, 0u);                                    // plane k0
    TPack<T> vv[3];
    vv[0] = *reinterpret_cast<const TPack<T>*>(reinterpret_cast<const T*>(us_b + 0 * C::U_STAGE_BYTES) + c_off);
    vv[1] = *reinterpret_cast<const TPack<T>*>(reinterpret_cast<const T*>(us_b + 1 * C::U_STAGE_BYTES) + c_off);

    double dot_acc = 0.0;
    T* out_own = out + col + (long long)k0 * g.plane;

    // plane kb+s lives in stage (s+1) % R; its n-th tenant (n = trips so far, +1 once the stage index wrapped)
    // completes phase n of the stage's mbarrier, i.e. parity n & 1
    unsigned trip = 0;
    for (int kb = k0; kb < k1; kb += TRING_R, ++trip) {
#pragma unroll
        for (int s = 0; s < TRING_R; ++s) {
            const int k = kb + s;
            if (k >= k1) break;                                 // CTA-uniform
            const int sc = (s + 1) % TRING_R, sp = (s + 2) % TRING_R;
            mbar_wait(bars + sp, (trip + (s + 2 >= TRING_R ? 1u : 0u)) & 1u);   // plane k+1 has landed
            __syncthreads();                                    // everybody is done with plane k-1's stage

            // refill the stage that held plane k-1 (stage s) with plane k+TRING_P+1
            if (tid == 0 && kk_issue <= k1) issue(kk_issue, s, kk_issue < k1);
            ++kk_issue;

            const T* Sc = reinterpret_cast<const T*>(us_b + sc * C::U_STAGE_BYTES);
            const TPack<T> v_m = vv[s % 3], v_c = vv[(s + 1) % 3];
            const TPack<T> v_p = *reinterpret_cast<const TPack<T>*>(reinterpret_cast<const T*>(us_b + sp * C::U_STAGE_BYTES) + c_off);
            vv[(s + 2) % 3] = v_p;
            const T xw = Sc[c_off - 1], xe = Sc[c_off + CPT];
            const TPack<T> ys = *reinterpret_cast<const TPack<T>*>(Sc + c_off - PITCH);
            const TPack<T> yn = *reinterpret_cast<const TPack<T>*>(Sc + c_off + PITCH);
            TPack<T> bb;
            if (MODE != 0) bb = *reinterpret_cast<const TPack<T>*>(reinterpret_cast<const T*>(bs_b + sc * C::B_BYTES) + b_off);
            unsigned int fword;
            if (CPT == 2) fword = *reinterpret_cast<const unsigned short*>(fs + sc * C::F_BYTES + b_off);
            else          fword = *reinterpret_cast<const unsigned int*>(fs + sc * C::F_BYTES + b_off);

            TPack<T> o;
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
                } else {
                    o.v[c] = unk ? v_c.v[c] + w * (bb.v[c] - au) * dtab[64 + (f & 63u)] : (T)0;
                    if (DOT) dot_acc += (double)bb.v[c] * (double)o.v[c];
                }
            }
            constexpr unsigned int UNKS = (CPT == 2) ? 0x4040u : 0x40404040u;
            if (inb && (fword & UNKS)) *reinterpret_cast<TPack<T>*>(out_own) = o;
            out_own += g.plane;
        }
    }

    if (DOT) {
        double v[1] = {dot_acc};
        grid_reduce<1>(v, red_partials, red_counter, red_out);
    }
}

template <typename T, int MODE>
size_t tma_smem_bytes() {
    typedef TCfg<T> C;
    return (size_t)TRING_R * (C::U_STAGE_BYTES + (MODE != 0 ? C::B_BYTES : 0) + C::F_BYTES) + 64 + 128 * sizeof(T) + 128;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled encode_fn() {
    static bool tried = false;
    static PFN_encodeTiled fn = nullptr;
    if (!tried) {
        tried = true;
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(f);
        else
            cudaGetLastError();
    }
    return fn;
}

// tensor map over a field with ghost planes: origin = ghost plane below plane 0, dims (nx, ny, nz + 2)
bool encode_field(CUtensorMap* m, const void* plane0, int elem_bytes, const Grid& g, int box_x, int box_y) {
    PFN_encodeTiled enc = encode_fn();
    if (!enc) return false;
    const CUtensorMapDataType dt = elem_bytes == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64
                                 : (elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8);
    char* base = const_cast<char*>(static_cast<const char*>(plane0)) - (size_t)g.plane * elem_bytes;
    const cuuint64_t dims[3] = {(cuuint64_t)g.nx, (cuuint64_t)g.ny, (cuuint64_t)g.nz + 2};
    const cuuint64_t strides[2] = {(cuuint64_t)g.nx * elem_bytes, (cuuint64_t)g.plane * elem_bytes};
    const cuuint32_t box[3] = {(cuuint32_t)box_x, (cuuint32_t)box_y, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    return enc(m, dt, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename T, int MODE, bool DOT>
bool launch_tma(const L0Args& a, cudaStream_t st) {
    typedef TCfg<T> C;
    CUtensorMap tu, tb, tf;
    if (!encode_field(&tu, a.u, (int)sizeof(T), a.g, C::PITCH, C::TY + 2)) return false;
    if (MODE != 0) { if (!encode_field(&tb, a.b, (int)sizeof(T), a.g, C::TX, C::TY)) return false; }
    else tb = tu;
    if (!encode_field(&tf, a.flags, 1, a.g, C::TX, C::TY)) return false;
    static unsigned long long configured = 0;
    const size_t smem = tma_smem_bytes<T, MODE>();
    if (first_use_on_this_device(configured))
        cudaFuncSetAttribute(l0_tma_kernel<T, MODE, DOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int zc = pick_zchunk(a.g, a.n_sm);
    dim3 grid((a.g.nx + C::TX - 1) / C::TX, (a.g.ny + C::TY - 1) / C::TY, (a.g.nz + zc - 1) / zc);
    l0_tma_kernel<T, MODE, DOT><<<grid, C::NT, smem, st>>>(a.g, tu, tb, tf, static_cast<T*>(a.out), (T)a.w, zc,
                                                           a.red_partials, a.red_counter, a.red_out);
    return true;
}

}  // namespace

// non-periodic box, rows of the connectivity bytes 16-byte aligned, fp32 multigrid vectors, no in-kernel halo work
bool tma_supported(const L0Args& a, int mode) {
    if (mode != 0 && mode != 1) return false;
    if (mode == 1 && sizeof(mg_t) != 4) return false;
    if ((a.g.nx & 15) || a.g.periodic) return false;
    if (a.hin.flag_lo || a.hin.flag_hi || a.hout.flag_lo || a.hout.flag_hi) return false;
    return encode_fn() != nullptr;
}

bool tma_launch(const L0Args& a, int mode, bool dot, cudaStream_t st) {
    if (mode == 0) return dot ? launch_tma<double, 0, true>(a, st) : launch_tma<double, 0, false>(a, st);
    if (sizeof(mg_t) == 4)
        return dot ? launch_tma<float, 1, true>(a, st) : launch_tma<float, 1, false>(a, st);
    return false;
}

}  // namespace oi
