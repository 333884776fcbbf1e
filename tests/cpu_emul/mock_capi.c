/* TEST INFRASTRUCTURE ONLY -- a stand-in for libopenimpala_b200.so that answers the
 * C-ABI of include/openimpala_b200.h on the CPU with the oracle (oracle/oi_oracle.c).
 *
 * Purpose: the host layer (readers -> Diffusion / tTortuosity apps -> TortuosityHypre /
 * EffectiveDiffusivityHypre classes -> results.txt / plotfiles / streamed upload) can be
 * run in the CPU-only build container by LD_PRELOADing this file in front of the real
 * library (tests/test_host_apps.py::test_host_layer_on_the_mock_*).  It is never built,
 * loaded or shipped by the product: the product library has no CPU fallback and the apps
 * abort without a CUDA device.  Only the entry points the host layer calls exist here. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/openimpala_b200.h"

/* oracle/oi_oracle.c */
int64_t oo_count_phase_i32(const int32_t* f, int64_t n, int32_t phase);
void oo_remspot(int32_t* q, int nx, int ny, int nz);
int64_t oo_activity_mask(const int32_t* phase, int32_t phase_id, int nx, int ny, int nz, int dir, uint8_t* mask, int capped);
void oo_fillmtx(double* a, double* rhs, double* xinit, const int32_t* p, const uint8_t* mask, int nx, int ny, int nz,
                const double* dxinv, double vlo, double vhi, int32_t phase, int dir);
int oo_solve_pcg(const double* a, const double* rhs, double* x, int nx, int ny, int nz, double eps, int maxiter, double* relres);
void oo_fluxes(const double* x, const uint8_t* mask, int nx, int ny, int nz, int dir, const double* dx, double* fin,
               double* fout, int64_t* n_in, int64_t* n_out);
void oo_effdiff_fillmtx(double* a, double* rhs, double* xinit, const int32_t* phase, int32_t phase_id, int nx, int ny,
                        int nz, const double* dx, int dir);
int oo_effdiff_solve(const double* a, const double* rhs, double* x, int nx, int ny, int nz, double eps, int maxiter,
                     double* relres);
void oo_effdiff_gradient_sums(const double* chi, const int32_t* phase, int32_t phase_id, int nx, int ny, int nz,
                              const double* dx, double* sums3, int64_t* n_active);

struct oi_solver {
    oi_params prm;
    int64_t n;
    int32_t* phase;
    uint8_t* mask;
    double* x;
    int64_t n_active;
    int mask_built, solved;
    uint8_t* stage[2];
    int stage_planes;
    int64_t planes_received;
    oi_solve_info info;
};

static char g_err[256] = "";
static int fail(const char* msg) { snprintf(g_err, sizeof g_err, "mock: %s", msg); return OI_ERR_INVALID; }

int oi_version(void) { return OI_B200_VERSION; }
const char* oi_last_error(void) { return g_err; }
int oi_device_count(int* count) { if (count) *count = 1; return OI_OK; }

void oi_default_params(oi_params* p) {
    memset(p, 0, sizeof *p);
    p->direction = OI_DIR_X; p->phase_id = 1; p->vlo = 0.0; p->vhi = 1.0;
    p->dx[0] = p->dx[1] = p->dx[2] = 1.0;
    p->eps = 1e-9; p->maxiter = 200; p->device = -1; p->flux_polish = 0;
}

int oi_count_phase_i32(const int32_t* f, int64_t n, int32_t phase, int64_t* pc, int64_t* tc) {
    if (!f && n) return fail("null field");
    if (pc) *pc = oo_count_phase_i32(f, n, phase);
    if (tc) *tc = n;
    return OI_OK;
}

int oi_create(oi_solver** out, const oi_params* p) {
    if (!out || !p) return fail("null argument");
    if (p->nx <= 0 || p->ny <= 0 || p->nz <= 0 || p->eps <= 0 || p->maxiter <= 0) return fail("bad parameters");
    if (p->comm) return fail("single slab only");
    oi_solver* S = (oi_solver*)calloc(1, sizeof *S);
    S->prm = *p;
    S->n = (int64_t)p->nx * p->ny * p->nz;
    S->phase = (int32_t*)calloc((size_t)S->n, sizeof(int32_t));
    S->mask = (uint8_t*)calloc((size_t)S->n, 1);
    S->x = (double*)calloc((size_t)S->n, sizeof(double));
    S->planes_received = -1;
    *out = S;
    return OI_OK;
}

int oi_destroy(oi_solver* S) {
    if (!S) return OI_OK;
    free(S->phase); free(S->mask); free(S->x); free(S->stage[0]); free(S->stage[1]);
    free(S);
    return OI_OK;
}

int oi_set_phase_i32(oi_solver* S, const int32_t* host) {
    if (!S || !host) return fail("null argument");
    memcpy(S->phase, host, (size_t)S->n * sizeof(int32_t));
    S->mask_built = S->solved = 0;
    return OI_OK;
}

int oi_phase_stream_begin(oi_solver* S, int32_t max_planes) {
    if (!S || max_planes <= 0) return fail("bad argument");
    const int planes = max_planes < S->prm.nz ? max_planes : S->prm.nz;
    for (int w = 0; w < 2; ++w) {
        free(S->stage[w]);
        S->stage[w] = (uint8_t*)malloc((size_t)planes * S->prm.nx * S->prm.ny);
    }
    S->stage_planes = planes;
    S->planes_received = 0;
    return OI_OK;
}
int oi_phase_stream_buffer(oi_solver* S, int32_t which, uint8_t** buf) {
    if (!S || !buf || which < 0 || which > 1 || S->planes_received < 0) return fail("bad argument");
    *buf = S->stage[which];
    return OI_OK;
}
int oi_phase_stream_submit(oi_solver* S, int32_t which, int32_t z0, int32_t nz) {
    if (!S || which < 0 || which > 1 || S->planes_received < 0) return fail("bad argument");
    if (nz <= 0 || nz > S->stage_planes || z0 < 0 || z0 + nz > S->prm.nz) return fail("chunk outside the slab");
    const int64_t plane = (int64_t)S->prm.nx * S->prm.ny;
    for (int64_t q = 0; q < plane * nz; ++q) S->phase[plane * z0 + q] = S->stage[which][q];
    S->planes_received += nz;
    return OI_OK;
}
int oi_phase_stream_end(oi_solver* S) {
    if (!S || S->planes_received != S->prm.nz) return fail("every plane of the slab must be submitted exactly once");
    S->planes_received = -1;
    S->mask_built = S->solved = 0;
    return OI_OK;
}

int oi_remspot(oi_solver* S, int32_t passes) {
    if (!S) return fail("null handle");
    for (int p = 0; p < passes; ++p) oo_remspot(S->phase, S->prm.nx, S->prm.ny, S->prm.nz);
    return OI_OK;
}

int oi_build_mask(oi_solver* S, int64_t* n_active) {
    if (!S) return fail("null handle");
    if (S->prm.problem == OI_PROBLEM_CELL) {
        S->n_active = 0;
        for (int64_t q = 0; q < S->n; ++q) { S->mask[q] = S->phase[q] == S->prm.phase_id; S->n_active += S->mask[q]; }
    } else {
        S->n_active = oo_activity_mask(S->phase, S->prm.phase_id, S->prm.nx, S->prm.ny, S->prm.nz, S->prm.direction, S->mask, 0);
    }
    S->mask_built = 1;
    if (n_active) *n_active = S->n_active;
    return OI_OK;
}

int oi_solve(oi_solver* S, oi_solve_info* info) {
    if (!S || !S->mask_built) return fail("call oi_build_mask first");
    const oi_params* p = &S->prm;
    memset(&S->info, 0, sizeof S->info);
    S->info.rel_residual = NAN;
    if (S->n_active > 0) {
        double* a = (double*)malloc(sizeof(double) * 7 * (size_t)S->n);
        double* rhs = (double*)malloc(sizeof(double) * (size_t)S->n);
        memset(S->x, 0, sizeof(double) * (size_t)S->n);
        double relres = NAN;
        int it;
        /* the oracle's Krylov method is Jacobi-PCG: give it the iterations it needs */
        const int maxit = p->maxiter > 1000000 / 200 ? 1000000 : p->maxiter * 200;
        if (p->problem == OI_PROBLEM_CELL) {
            oo_effdiff_fillmtx(a, rhs, S->x, S->phase, p->phase_id, p->nx, p->ny, p->nz, p->dx, p->direction);
            it = oo_effdiff_solve(a, rhs, S->x, p->nx, p->ny, p->nz, p->eps, maxit, &relres);
        } else {
            const double dxinv[3] = {1.0 / (p->dx[0] * p->dx[0]), 1.0 / (p->dx[1] * p->dx[1]), 1.0 / (p->dx[2] * p->dx[2])};
            oo_fillmtx(a, rhs, S->x, S->phase, S->mask, p->nx, p->ny, p->nz, dxinv, p->vlo, p->vhi, p->phase_id, p->direction);
            it = oo_solve_pcg(a, rhs, S->x, p->nx, p->ny, p->nz, p->eps, maxit, &relres);
            /* the device keeps exact zeros off the percolating cells (the reference's identity rows do too) */
            for (int64_t q = 0; q < S->n; ++q) if (!S->mask[q]) S->x[q] = 0.0;
        }
        S->info.iterations = it;
        S->info.rel_residual = relres;
        S->info.converged = isfinite(relres) && relres >= 0.0 && relres <= p->eps;
        free(a); free(rhs);
    } else if (p->problem == OI_PROBLEM_CELL) {
        S->info.converged = 1;
        S->info.rel_residual = 0.0;
    }
    S->solved = 1;
    if (info) *info = S->info;
    return OI_OK;
}

int oi_fluxes(oi_solver* S, double* fin, double* fout, int64_t* nin, int64_t* nout) {
    if (!S || !S->mask_built) return fail("call oi_build_mask first");
    double a = 0.0, b = 0.0;
    int64_t ca = 0, cb = 0;
    if (S->n_active > 0)
        oo_fluxes(S->x, S->mask, S->prm.nx, S->prm.ny, S->prm.nz, S->prm.direction, S->prm.dx, &a, &b, &ca, &cb);
    if (fin) *fin = a;
    if (fout) *fout = b;
    if (nin) *nin = ca;
    if (nout) *nout = cb;
    return OI_OK;
}

int oi_cell_gradient_sums(oi_solver* S, double* sums3, int64_t* n_active) {
    if (!S || !sums3 || S->prm.problem != OI_PROBLEM_CELL) return fail("cell-problem handle required");
    int64_t na = 0;
    oo_effdiff_gradient_sums(S->x, S->phase, S->prm.phase_id, S->prm.nx, S->prm.ny, S->prm.nz, S->prm.dx, sums3, &na);
    if (n_active) *n_active = na;
    return OI_OK;
}

int oi_check_matrix_properties(oi_solver* S, int32_t* ok) {
    if (!S || !S->mask_built) return fail("call oi_build_mask first");
    if (ok) *ok = 1;
    return OI_OK;
}

int oi_get_mask_u8(oi_solver* S, uint8_t* host) {
    if (!S || !host || !S->mask_built) return fail("mask not built");
    memcpy(host, S->mask, (size_t)S->n);
    return OI_OK;
}
int oi_get_solution(oi_solver* S, double* host) {
    if (!S || !host) return fail("null argument");
    memcpy(host, S->x, sizeof(double) * (size_t)S->n);
    return OI_OK;
}
