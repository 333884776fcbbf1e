// Peer-memory halo exchange between z-slabs (one process per GPU).
//
// Replaces AMReX FillBoundary of the reference (src/props/TortuosityHypre.cpp:339,
// 584-585, 1033) on the solve path: each rank stores its two boundary xy-planes
// straight into the ghost planes of its z-neighbours (whose arenas are mapped with
// CUDA IPC over NVLink) and then publishes a sequence number in the neighbour's
// flag word.  The consumer's stream waits on its own flag word before the next
// stencil kernel.  One launch per exchange, both directions in flight at once,
// no host round trip and no NCCL proxy on the path.
#include "oi_kernels.h"

namespace oi {

namespace {

template <typename V>
__global__ void __launch_bounds__(256)
halo_push_kernel(const V* __restrict__ src_lo, V* __restrict__ dst_lo, const V* __restrict__ src_hi,
                 V* __restrict__ dst_hi, long long nv, unsigned int* flag_lo, unsigned int* flag_hi,
                 unsigned int seq, unsigned int* counter) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (dst_lo)
        for (long long i = t0; i < nv; i += stride) dst_lo[i] = src_lo[i];
    if (dst_hi)
        for (long long i = t0; i < nv; i += stride) dst_hi[i] = src_hi[i];
    // publish: every thread's peer stores are fenced system-wide, the last block
    // to arrive writes the sequence number behind them
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int ticket = atomicAdd(counter, 1u);
        if (ticket == gridDim.x - 1) {
            *counter = 0u;
            __threadfence_system();
            if (flag_lo) *reinterpret_cast<volatile unsigned int*>(flag_lo) = seq;
            if (flag_hi) *reinterpret_cast<volatile unsigned int*>(flag_hi) = seq;
        }
    }
}

// Fallback consumer wait when stream memory operations are unavailable.
__global__ void halo_wait_kernel(const unsigned int* flag_a, const unsigned int* flag_b, unsigned int seq) {
    if (flag_a)
        while ((int)(*reinterpret_cast<const volatile unsigned int*>(flag_a) - seq) < 0) __nanosleep(200);
    if (flag_b)
        while ((int)(*reinterpret_cast<const volatile unsigned int*>(flag_b) - seq) < 0) __nanosleep(200);
    __threadfence_system();
}

}  // namespace

void halo_push(const void* src_lo, void* dst_lo, const void* src_hi, void* dst_hi, size_t plane_bytes,
               unsigned int* flag_lo, unsigned int* flag_hi, unsigned int seq, unsigned int* counter,
               int n_sm, cudaStream_t st) {
    const uintptr_t all = (uintptr_t)src_lo | (uintptr_t)dst_lo | (uintptr_t)src_hi | (uintptr_t)dst_hi |
                          (uintptr_t)plane_bytes;
    if ((all & 15u) == 0) {
        const long long nv = (long long)(plane_bytes / 16);
        long long nb = (nv + 1023) / 1024;
        nb = nb < 1 ? 1 : (nb > 2LL * n_sm ? 2LL * n_sm : nb);
        halo_push_kernel<uint4><<<(unsigned)nb, 256, 0, st>>>(
            static_cast<const uint4*>(src_lo), static_cast<uint4*>(dst_lo), static_cast<const uint4*>(src_hi),
            static_cast<uint4*>(dst_hi), nv, flag_lo, flag_hi, seq, counter);
    } else {
        // fields are 4- or 8-byte elements: every plane is a whole number of words
        const long long nv = (long long)(plane_bytes / 4);
        long long nb = (nv + 1023) / 1024;
        nb = nb < 1 ? 1 : (nb > 2LL * n_sm ? 2LL * n_sm : nb);
        halo_push_kernel<unsigned int><<<(unsigned)nb, 256, 0, st>>>(
            static_cast<const unsigned int*>(src_lo), static_cast<unsigned int*>(dst_lo),
            static_cast<const unsigned int*>(src_hi), static_cast<unsigned int*>(dst_hi), nv, flag_lo, flag_hi,
            seq, counter);
    }
}

void halo_wait_spin(const unsigned int* flag_a, const unsigned int* flag_b, unsigned int seq, cudaStream_t st) {
    halo_wait_kernel<<<1, 1, 0, st>>>(flag_a, flag_b, seq);
}

}  // namespace oi
