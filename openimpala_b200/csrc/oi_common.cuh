// Shared device helpers for the openimpala_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace oi {

// Element type of every vector inside the multigrid preconditioner (level-0 z /
// scratch / residual copy, all coarse-level vectors).  The Krylov vectors x, r, p,
// q = A p and the operator apply stay fp64; the V-cycle only has to be a good
// SPD approximation of A^-1, and a study on the sample image and on sphere packs
// (tools/mg_fp32_study.py) shows identical PCG iteration counts (+-1) down to
// 1e-12 with an fp32 V-cycle.  Build with -DOI_MG_FP64 for an all-fp64 V-cycle.
#ifdef OI_MG_FP64
typedef double mg_t;
#else
typedef float mg_t;
#endif

template <typename T> struct Vec2;
template <> struct Vec2<double> { typedef double2 type; };
template <> struct Vec2<float> { typedef float2 type; };

// ---- connectivity byte (one per cell; replaces the 7 stored fp64 coefficients
// of tortuosity_fillmtx, reference src/props/TortuosityHypreFill.F90:96-228) ----
// bits 0..5: this cell is an unknown AND the -x,+x,-y,+y,-z,+z neighbour is an
//            active cell of the same phase (coefficient -c_d, diag += c_d;
//            F90:126-166).  Neighbours may be Dirichlet cells.
// bit 6    : unknown  = active, not on a Dirichlet plane (eliminated system row)
// bit 7    : active cell on the inlet/outlet plane (identity row, b = vlo/vhi;
//            F90:192-228)
enum : uint8_t {
    F_XM = 1u << 0, F_XP = 1u << 1, F_YM = 1u << 2, F_YP = 1u << 3,
    F_ZM = 1u << 4, F_ZP = 1u << 5, F_UNK = 1u << 6, F_DIR = 1u << 7,
    F_FACES = 0x3f
};

// Problem geometry of one level-0 slab, passed by value to kernels.
struct Grid {
    int nx, ny, nz;        // local box (nz = local planes)
    int nzg;               // global nz
    int z0;                // global index of local plane 0
    long long plane;       // nx*ny
    double cx, cy, cz;     // 1/dx^2, 1/dy^2, 1/dz^2  (TortuosityHypre.cpp:580-582)
    // Cell problem of the homogenisation path (EffectiveDiffusivityHypre,
    // src/props/EffDiffFillMtx.F90:109-258): the box is periodic and every active row
    // has the full diagonal 2(cx+cy+cz) because a face towards the solid adds 1/dx^2
    // to the diagonal without a coupling (F90:156-165).  Zero / 0.0 for tortuosity.
    int periodic;          // bit 0: x, bit 1: y, bit 2: z
    double diag_full;      // > 0: diagonal of every unknown row; 0: sum of its couplings
    double hx, hy, hz;     // cell sizes (Geometry::CellSize)
};

enum : int { PER_X = 1, PER_Y = 2, PER_Z = 4 };

// diagonal of an unknown row from its connectivity bits
template <typename T>
__device__ __forceinline__ T row_diag(unsigned int f, const Grid& g) {
    if (g.diag_full > 0.0) return (T)g.diag_full;
    return (T)g.cx * (T)__popc(f & 0x03u) + (T)g.cy * (T)__popc(f & 0x0cu) + (T)g.cz * (T)__popc(f & 0x30u);
}
// x / y index of the -1 / +1 neighbour with periodic wrap; -1 = outside the box
__device__ __forceinline__ int wrap_lo(int i, int n, bool periodic) { return i > 0 ? i - 1 : (periodic ? n - 1 : -1); }
__device__ __forceinline__ int wrap_hi(int i, int n, bool periodic) { return i + 1 < n ? i + 1 : (periodic ? 0 : -1); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic grid reduction of NV doubles per thread.
//   partials : [gridDim_total][NV] scratch,  counter : zeroed uint (self-resetting)
//   out      : [NV] result, written by the last block in fixed block order.
// Every thread of the block must call this (contains __syncthreads).
template <int NV>
__device__ __forceinline__ void grid_reduce(double (&v)[NV], double* partials,
                                            unsigned int* counter, double* out) {
    __shared__ double s_part[32][NV];
    __shared__ bool s_last;
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int nthreads = blockDim.x * blockDim.y * blockDim.z;
    const int lane = tid & 31, warp = tid >> 5, nwarps = (nthreads + 31) >> 5;
    const unsigned int nblocks = gridDim.x * gridDim.y * gridDim.z;
    const unsigned int bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        double s = warp_sum(v[q]);
        if (lane == 0) s_part[warp][q] = s;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            double s = (lane < nwarps) ? s_part[lane][q] : 0.0;
            s = warp_sum(s);
            if (lane == 0) partials[(size_t)bid * NV + q] = s;
        }
    }
    if (tid == 0) {
        __threadfence();
        unsigned int ticket = atomicAdd(counter, 1u);
        s_last = (ticket == nblocks - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        if (warp == 0) {
#pragma unroll
            for (int q = 0; q < NV; ++q) {
                double s = 0.0;
                for (unsigned int b = lane; b < nblocks; b += 32)
                    s += __ldcg(&partials[(size_t)b * NV + q]);
                s = warp_sum(s);
                if (lane == 0) out[q] = s;
            }
            if (lane == 0) *counter = 0u;
        }
    }
}

// ---- peer halo fused into the producing / consuming kernels (z-slabs, one process per GPU) ----
// Replaces AMReX FillBoundary on the solve path (reference src/props/TortuosityHypre.cpp:339,
// 584-585, 1033).  A kernel that WRITES a field with ghost planes stores its two boundary planes
// straight into the z-neighbours' ghost planes (their arenas are mapped through CUDA IPC over
// NVLink) and the last of its boundary CTAs publishes the exchange's sequence number in the
// neighbour's flag word; a kernel that READS ghost planes lets only the CTAs that touch them spin
// on the local flag word, and those CTAs are scheduled last, so the transfer and any skew between
// the ranks hide behind the interior planes.  No push kernel, no stream wait, no host round trip.
struct HaloOut {
    void* dst_lo;                   // lower neighbour's ghost plane above its top plane   (nullptr: none)
    void* dst_hi;                   // upper neighbour's ghost plane below its plane 0     (nullptr: none)
    unsigned int* flag_lo;          // their flag words
    unsigned int* flag_hi;
    unsigned int* counter;          // local: [0] boundary CTAs done on the low side, [1] high side / all blocks
    unsigned int seq;
};
struct HaloIn {
    const unsigned int* flag_lo;    // local flag word written by the lower neighbour (nullptr: no wait)
    const unsigned int* flag_hi;    // ... by the upper neighbour
    unsigned int seq;
};

// one thread: spin until the flag word has reached seq (acquire at system scope)
__device__ __forceinline__ void halo_spin(const unsigned int* flag, unsigned int seq) {
    unsigned int v;
    for (;;) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if ((int)(v - seq) >= 0) break;
        __nanosleep(64);
    }
}
// Every thread of the CTA, after its peer stores: the last of `expected` CTAs to arrive publishes seq
// in the neighbours' flag words (either may be null).  Contains __syncthreads.
__device__ __forceinline__ void halo_publish(unsigned int* counter, unsigned int expected, unsigned int* flag_a,
                                             unsigned int* flag_b, unsigned int seq) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) {
        const unsigned int ticket = atomicAdd(counter, 1u);
        if (ticket == expected - 1u) {
            *counter = 0u;
            __threadfence_system();
            if (flag_a) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag_a), "r"(seq) : "memory");
            if (flag_b) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag_b), "r"(seq) : "memory");
        }
    }
}

// Function attributes (cudaFuncSetAttribute) belong to a device's context, so the one-time
// configuration of a kernel has to happen once per device, not once per process: true the
// first time a call site is reached with the current device.
inline bool first_use_on_this_device(unsigned long long& seen_devices) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    // (host threads may drive different devices concurrently: atomic read-modify-write)
    const unsigned long long before = __atomic_fetch_or(&seen_devices, bit, __ATOMIC_ACQ_REL);
    return (before & bit) == 0;
}

}  // namespace oi
