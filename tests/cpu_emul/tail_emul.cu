// Host emulation of the one-CTA coarse tail (TEST INFRASTRUCTURE, not product code).
// It compiles the very same tail_cycle() the CUDA kernel runs
// (openimpala_b200/csrc/oi_coarse_tail.cuh) and executes it on the host with a
// sequential `step`, so that the cycle's control flow and index arithmetic can be
// checked against an independent numpy V-cycle without a GPU
// (tests/test_host_cpu.py::test_coarse_tail_cycle_on_the_host).
#include <cstdlib>

#include "../../openimpala_b200/csrc/oi_coarse_tail.cuh"

namespace {
struct SeqStep {
    template <class F>
    __host__ __device__ void operator()(int n, F f) const {
        for (int i = 0; i < n; ++i) f(i);
    }
};
}  // namespace

// dims: 6 ints per level (nx, ny, nz, fx, fy, fz); fields: 7 pointers per level
// (cxp, cyp, czp, dg, x, b, t), each nx*ny*nz floats, x fastest, no ghost planes.
extern "C" int oi_tail_emulate(int n_levels, const int* dims, int periodic, float** fields, int deg, const double* w,
                               int deg_c, const double* wc, int staged) {
    if (n_levels < 1 || n_levels > oi::TAIL_MAX_LEVELS || deg < 1 || deg > 16 || deg_c < 1 || deg_c > 16) return 1;
    if (sizeof(oi::mg_t) != sizeof(float)) return 2;
    oi::TailArgs a{};
    a.n_levels = n_levels; a.deg = deg; a.deg_c = deg_c;
    for (int q = 0; q < deg; ++q) a.w[q] = (oi::mg_t)w[q];
    for (int q = 0; q < deg_c; ++q) a.wc[q] = (oi::mg_t)wc[q];
    for (int l = 0; l < n_levels; ++l) {
        oi::CoarseLevel& L = a.L[l];
        L.nx = dims[6 * l + 0]; L.ny = dims[6 * l + 1]; L.nz = dims[6 * l + 2];
        L.fx = dims[6 * l + 3]; L.fy = dims[6 * l + 4]; L.fz = dims[6 * l + 5];
        L.z0 = 0; L.nzg = L.nz; L.plane = (long long)L.nx * L.ny;
        L.periodic = periodic; L.replicated = 0;
        float** f = fields + 7 * l;
        L.cxp = f[0]; L.cyp = f[1]; L.czp = f[2]; L.dg = f[3];
        L.dgx = L.dgy = L.dgz = nullptr;
        L.x = reinterpret_cast<oi::mg_t*>(f[4]); L.b = reinterpret_cast<oi::mg_t*>(f[5]); L.t = reinterpret_cast<oi::mg_t*>(f[6]);
    }
    if (staged) {
        // the kernel's dynamic shared memory is a plain host buffer here
        const size_t bytes = oi::tail_staged_bytes(a);
        float* buf = static_cast<float*>(aligned_alloc(16, (bytes + 15) / 16 * 16));
        if (!buf) return 3;
        for (size_t q = 0; q < bytes / sizeof(float); ++q) buf[q] = -12345.f;   // garbage: nothing may be assumed zero
        oi::tail_cycle_staged(a, SeqStep(), buf);
        free(buf);
    } else {
        oi::tail_cycle(a, SeqStep());
    }
    return 0;
}
