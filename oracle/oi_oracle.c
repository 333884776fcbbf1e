/*
 * oi_oracle.c -- plain-C CPU restatement of the reference TortuosityHypre path.
 *
 * TEST INFRASTRUCTURE ONLY: nothing in the product (openimpala_b200/) may call
 * or link this file.  Users: tests/, __graft_entry__.smoke(), and bench.py's
 * cpu_baseline / --impl reference legs (as the thing timed on the host cores,
 * never as the GPU arm).
 *
 * Parity status: "unpinned" against real HYPRE output (the reference cannot be
 * built here: no MPI / gfortran / AMReX / HYPRE).  Pinned against: the
 * reference's checkMatrixProperties invariants, analytic known answers, and
 * the numpy/scipy restatement (oracle/oi_numpy.py) on the sample image.
 *
 * Each function cites the reference file:line it follows (paths relative to
 * the reference tree).  Layout: x fastest, z slowest; cell (i,j,k) at
 * (k*ny + j)*nx + i.
 *
 * The solver is the one deliberate difference: the reference hands the stored
 * 7-coefficient matrix to HYPRE FlexGMRES + SMG (src/props/
 * TortuosityHypre.cpp:664-683, HYPRE v2.32.0, not vendored); here the same
 * stored matrix (7 doubles per cell, identity rows included) is solved by a
 * Jacobi-preconditioned CG on the rows' symmetric part with the reference's
 * stopping rule ||b - Ax||_2 <= eps * ||b||_2 (b = full rhs).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define IDX(i, j, k) (((int64_t)(k) * ny + (j)) * nx + (i))

int oo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Explicit thread count for the timed CPU arm: torchrun exports OMP_NUM_THREADS=1 to its workers,
 * which would silently turn the baseline into a one-core run.  Returns the count now in force. */
int oo_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

/* VolumeFraction::value, src/props/VolumeFraction.cpp:22-66 */
int64_t oo_count_phase_i32(const int32_t* f, int64_t n, int32_t phase) {
    int64_t c = 0;
#pragma omp parallel for reduction(+ : c) schedule(static)
    for (int64_t i = 0; i < n; ++i) c += (f[i] == phase);
    return c;
}

/* tortuosity_remspot, src/props/Tortuosity_filcc.F90:88-177: in-place, i fastest;
 * a voxel none of whose in-domain 6 neighbours holds its value flips 0 <-> 1
 * (any non-zero value becomes 0).  One pass over the whole domain as one box. */
void oo_remspot(int32_t* q, int nx, int ny, int nz) {
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                const int64_t c = IDX(i, j, k);
                const int32_t v = q[c];
                int connected = 0;
                if (i > 0 && q[c - 1] == v) connected = 1;
                else if (i + 1 < nx && q[c + 1] == v) connected = 1;
                else if (j > 0 && q[c - nx] == v) connected = 1;
                else if (j + 1 < ny && q[c + nx] == v) connected = 1;
                else if (k > 0 && q[c - (int64_t)nx * ny] == v) connected = 1;
                else if (k + 1 < nz && q[c + (int64_t)nx * ny] == v) connected = 1;
                if (!connected) q[c] = (v == 0) ? 1 : 0;
            }
}

/* parallelFloodFill, src/props/TortuosityHypre.cpp:297-389: the literal sweep --
 * in-place lexicographic (i fastest) pass over the whole box, repeated until a
 * pass changes nothing or `max_iter` passes were made (max_iter <= 0: no cap;
 * the reference caps at nx+ny+nz+2, :328).  Returns the number of passes. */
int oo_flood_fill(const int32_t* phase, int32_t phase_id, int nx, int ny, int nz, int dir,
                  int seed_plane, uint8_t* reached, int max_iter) {
    const int64_t n = (int64_t)nx * ny * nz;
    memset(reached, 0, (size_t)n);
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                const int d = dir == 0 ? i : (dir == 1 ? j : k);
                if (d == seed_plane && phase[IDX(i, j, k)] == phase_id) reached[IDX(i, j, k)] = 1; /* :313-324 */
            }
    int iter = 0, changed = 1;
    while (changed && (max_iter <= 0 || iter < max_iter)) {                                   /* :336 */
        ++iter;
        changed = 0;
        for (int k = 0; k < nz; ++k)
            for (int j = 0; j < ny; ++j)
                for (int i = 0; i < nx; ++i) {
                    const int64_t c = IDX(i, j, k);
                    if (reached[c] || phase[c] != phase_id) continue;                          /* :353 */
                    int hit = 0;                                                               /* :357-365 */
                    if (i + 1 < nx && reached[c + 1]) hit = 1;
                    else if (i > 0 && reached[c - 1]) hit = 1;
                    else if (j + 1 < ny && reached[c + nx]) hit = 1;
                    else if (j > 0 && reached[c - nx]) hit = 1;
                    else if (k + 1 < nz && reached[c + (int64_t)nx * ny]) hit = 1;
                    else if (k > 0 && reached[c - (int64_t)nx * ny]) hit = 1;
                    if (hit) { reached[c] = 1; changed = 1; }                                  /* :366-369 */
                }
    }
    return iter;
}

/* generateActivityMask, src/props/TortuosityHypre.cpp:394-558.  Returns the
 * active-cell count (sum of the mask, :549); mask is all zero when either face
 * has no seed (:508-514). */
int64_t oo_activity_mask(const int32_t* phase, int32_t phase_id, int nx, int ny, int nz, int dir,
                         uint8_t* mask, int capped) {
    const int64_t n = (int64_t)nx * ny * nz;
    const int nd = dir == 0 ? nx : (dir == 1 ? ny : nz);
    int64_t seeds_in = 0, seeds_out = 0;
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                const int d = dir == 0 ? i : (dir == 1 ? j : k);
                if (phase[IDX(i, j, k)] != phase_id) continue;
                if (d == 0) ++seeds_in;
                if (d == nd - 1) ++seeds_out;
            }
    memset(mask, 0, (size_t)n);
    if (seeds_in == 0 || seeds_out == 0) return 0;
    uint8_t* a = (uint8_t*)malloc((size_t)n);
    uint8_t* b = (uint8_t*)malloc((size_t)n);
    const int cap = capped ? nx + ny + nz + 2 : 0;
    oo_flood_fill(phase, phase_id, nx, ny, nz, dir, 0, a, cap);
    oo_flood_fill(phase, phase_id, nx, ny, nz, dir, nd - 1, b, cap);
    int64_t cnt = 0;
    for (int64_t c = 0; c < n; ++c) { mask[c] = a[c] & b[c]; cnt += mask[c]; }                /* :531-536 */
    free(a); free(b);
    return cnt;
}

/* tortuosity_fillmtx, src/props/TortuosityHypreFill.F90:44-314, on the whole
 * domain as one box.  a[7*m+s], s = C,-x,+x,-y,+y,-z,+z (F90:20-26).  The
 * out-of-domain ghost layer of the mask is inactive (TortuosityHypre.cpp:309,
 * 522; non-periodic FillBoundary never writes it).  xinit must be zeroed by
 * the caller like std::vector::resize does (TortuosityHypre.cpp:606). */
void oo_fillmtx(double* a, double* rhs, double* xinit, const int32_t* p, const uint8_t* mask,
                int nx, int ny, int nz, const double* dxinv, double vlo, double vhi, int32_t phase,
                int dir) {
    const double small_real = 1.0e-15;
    const int dom_hi[3] = {nx - 1, ny - 1, nz - 1};
#define ACTIVE(i, j, k) ((i) >= 0 && (i) < nx && (j) >= 0 && (j) < ny && (k) >= 0 && (k) < nz && \
                         p[IDX(i, j, k)] == phase && mask[IDX(i, j, k)] == 1)
#pragma omp parallel for schedule(static)
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                const int64_t m = IDX(i, j, k);
                double* row = a + 7 * m;
                for (int s = 0; s < 7; ++s) row[s] = 0.0;
                if (!ACTIVE(i, j, k)) {                       /* F90:111-118 */
                    row[0] = 1.0; rhs[m] = 0.0; xinit[m] = 0.0;
                    continue;
                }
                double diag = 0.0;                            /* F90:126-166 */
                if (ACTIVE(i - 1, j, k)) { row[1] = -dxinv[0]; diag += dxinv[0]; }
                if (ACTIVE(i + 1, j, k)) { row[2] = -dxinv[0]; diag += dxinv[0]; }
                if (ACTIVE(i, j - 1, k)) { row[3] = -dxinv[1]; diag += dxinv[1]; }
                if (ACTIVE(i, j + 1, k)) { row[4] = -dxinv[1]; diag += dxinv[1]; }
                if (ACTIVE(i, j, k - 1)) { row[5] = -dxinv[2]; diag += dxinv[2]; }
                if (ACTIVE(i, j, k + 1)) { row[6] = -dxinv[2]; diag += dxinv[2]; }
                row[0] = diag;
                if (fabs(diag) < small_real) {                /* F90:172-181 */
                    for (int s = 0; s < 7; ++s) row[s] = 0.0;
                    row[0] = 1.0; rhs[m] = 0.0; xinit[m] = 0.0;
                    continue;
                }
                rhs[m] = 0.0;
                int on_dirichlet = 0;                         /* F90:191-228 */
                const int d = dir == 0 ? i : (dir == 1 ? j : k);
                if (d == 0) {
                    for (int s = 0; s < 7; ++s) row[s] = 0.0;
                    row[0] = 1.0; rhs[m] = vlo; on_dirichlet = 1;
                } else if (d == dom_hi[dir]) {
                    for (int s = 0; s < 7; ++s) row[s] = 0.0;
                    row[0] = 1.0; rhs[m] = vhi; on_dirichlet = 1;
                }
                if (fabs(row[0] - 1.0) > small_real || on_dirichlet) {   /* F90:233-262 */
                    const double ext = (double)(dom_hi[dir] - 0);
                    const double factor = fabs(ext) < small_real ? 0.0 : 1.0 / ext;
                    xinit[m] = vlo + (vhi - vlo) * (double)d * factor;
                }
            }
#undef ACTIVE
}

/* y = A x for the stored 7-point struct matrix (what HYPRE's struct matvec
 * does with the values set at TortuosityHypre.cpp:635). */
static void matvec7(const double* a, const double* x, double* y, int nx, int ny, int nz) {
    const int64_t sx = 1, sy = nx, sz = (int64_t)nx * ny;
#pragma omp parallel for schedule(static)
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                const int64_t m = IDX(i, j, k);
                const double* r = a + 7 * m;
                double s = r[0] * x[m];
                if (r[1] != 0.0) s += r[1] * x[m - sx];
                if (r[2] != 0.0) s += r[2] * x[m + sx];
                if (r[3] != 0.0) s += r[3] * x[m - sy];
                if (r[4] != 0.0) s += r[4] * x[m + sy];
                if (r[5] != 0.0) s += r[5] * x[m - sz];
                if (r[6] != 0.0) s += r[6] * x[m + sz];
                y[m] = s;
            }
}

static double dot(const double* a, const double* b, int64_t n) {
    double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

/* solve(), src/props/TortuosityHypre.cpp:654-756, restated as Jacobi-PCG.
 * The identity rows are held fixed (x = rhs there), which makes the remaining
 * block symmetric positive definite; search directions are zero on those rows.
 * Stop: ||b - A x||_2 <= eps * ||b||_2 (HYPRE rule, ||b|| = 0 -> ||r0||).
 * Returns iterations; *relres = final ||r||/||b||. */
int oo_solve_pcg(const double* a, const double* rhs, double* x, int nx, int ny, int nz, double eps,
                 int maxiter, double* relres) {
    const int64_t n = (int64_t)nx * ny * nz;
    double* r = (double*)malloc(sizeof(double) * (size_t)n);
    double* z = (double*)malloc(sizeof(double) * (size_t)n);
    double* p = (double*)malloc(sizeof(double) * (size_t)n);
    double* q = (double*)malloc(sizeof(double) * (size_t)n);
    uint8_t* unk = (uint8_t*)malloc((size_t)n);
#pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < n; ++m) {
        const double* row = a + 7 * m;
        int off = 0;
        for (int s = 1; s < 7; ++s) off |= (row[s] != 0.0);
        unk[m] = (uint8_t)off;
        if (!off) x[m] = rhs[m] / row[0];          /* identity rows: exact */
    }
    matvec7(a, x, q, nx, ny, nz);
#pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < n; ++m) r[m] = unk[m] ? rhs[m] - q[m] : 0.0;
    const double bnorm = sqrt(dot(rhs, rhs, n));
    double rn = sqrt(dot(r, r, n));
    const double den = bnorm > 0.0 ? bnorm : rn;
    const double tol = eps * den;
    int it = 0;
    if (rn > tol) {
#pragma omp parallel for schedule(static)
        for (int64_t m = 0; m < n; ++m) { z[m] = unk[m] ? r[m] / a[7 * m] : 0.0; p[m] = z[m]; }
        double rz = dot(r, z, n);
        while (it < maxiter) {
            ++it;
            matvec7(a, p, q, nx, ny, nz);
#pragma omp parallel for schedule(static)
            for (int64_t m = 0; m < n; ++m) if (!unk[m]) q[m] = 0.0;
            const double alpha = rz / dot(p, q, n);
            double rr = 0.0;
#pragma omp parallel for reduction(+ : rr) schedule(static)
            for (int64_t m = 0; m < n; ++m) {
                x[m] += alpha * p[m];
                r[m] -= alpha * q[m];
                rr += r[m] * r[m];
            }
            rn = sqrt(rr);
            if (!(rn > tol)) break;
            double rzn = 0.0;
#pragma omp parallel for reduction(+ : rzn) schedule(static)
            for (int64_t m = 0; m < n; ++m) {
                z[m] = unk[m] ? r[m] / a[7 * m] : 0.0;
                rzn += r[m] * z[m];
            }
            const double beta = rzn / rz;
            rz = rzn;
#pragma omp parallel for schedule(static)
            for (int64_t m = 0; m < n; ++m) p[m] = z[m] + beta * p[m];
        }
    }
    if (relres) *relres = den > 0.0 ? rn / den : 0.0;
    free(r); free(z); free(p); free(q); free(unk);
    return it;
}

/* global_fluxes, src/props/TortuosityHypre.cpp:1000-1134 (x face area). */
void oo_fluxes(const double* x, const uint8_t* mask, int nx, int ny, int nz, int dir,
               const double* dx, double* fin, double* fout, int64_t* n_in, int64_t* n_out) {
    const int nd = dir == 0 ? nx : (dir == 1 ? ny : nz);
    const int64_t sd = dir == 0 ? 1 : (dir == 1 ? nx : (int64_t)nx * ny);
    double a = 0.0, b = 0.0;
    int64_t ca = 0, cb = 0;
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                const int d = dir == 0 ? i : (dir == 1 ? j : k);
                const int64_t c = IDX(i, j, k);
                if (d == 0 && mask[c]) {                                   /* :1067-1083 */
                    ++ca;
                    if (nd > 1 && mask[c + sd]) a += -((x[c + sd] - x[c]) / dx[dir]);
                }
                if (d == nd - 1 && mask[c]) {                              /* :1086-1103 */
                    ++cb;
                    if (nd > 1 && mask[c - sd]) b += -((x[c] - x[c - sd]) / dx[dir]);
                }
            }
    const double area = dir == 0 ? dx[1] * dx[2] : (dir == 1 ? dx[0] * dx[2] : dx[0] * dx[1]);
    *fin = a * area; *fout = b * area;                                     /* :1123-1133 */
    if (n_in) *n_in = ca;
    if (n_out) *n_out = cb;
}

/* value() tail, src/props/TortuosityHypre.cpp:782-877.  NaN/Inf conventions. */
double oo_tau(double fin, double fout, double active_vf, int nx, int ny, int nz, int dir,
              const double* dx, double vlo, double vhi, int converged, double* deff_out) {
    const double tiny = 1.0e-15, eps = 2.220446049250313e-16;
    if (deff_out) *deff_out = 0.0;
    if (active_vf <= eps || !converged) return NAN;                       /* :764-787 */
    const double mi = fabs(fin), mo = fabs(fout), avg = 0.5 * (mi + mo);
    if (avg > tiny && fabs(mi - mo) / avg > 1.0e-6) return NAN;           /* :794-823 */
    const double ext[3] = {nx * dx[0], ny * dx[1], nz * dx[2]};
    const double L = ext[dir];
    const double A = dir == 0 ? ext[1] * ext[2] : (dir == 1 ? ext[0] * ext[2] : ext[0] * ext[1]);
    const double grad = (vhi - vlo) / L;                                   /* :841 */
    if (avg < tiny) return INFINITY;                                       /* :846-851 */
    if (fabs(grad) < tiny) return INFINITY;                                /* :860-864 */
    const double deff = (avg / A) / fabs(grad);                            /* :868 */
    if (deff_out) *deff_out = deff;
    if (fabs(deff) < tiny) return INFINITY;                                /* :869-873 */
    return active_vf / deff;                                               /* :876 */
}

/* Whole path for one direction: mask -> matrix -> solve -> flux -> tau.
 * out[0]=tau out[1]=deff out[2]=active_vf out[3]=fin out[4]=fout out[5]=iters
 * out[6]=relres out[7]=n_active out[8]=solve seconds (wall, omp) */
int oo_tortuosity(const int32_t* phase, int nx, int ny, int nz, int32_t phase_id, int dir,
                  double vlo, double vhi, double eps, int maxiter, double* out) {
    const int64_t n = (int64_t)nx * ny * nz;
    const double dx[3] = {1.0, 1.0, 1.0}, dxinv[3] = {1.0, 1.0, 1.0};
    uint8_t* mask = (uint8_t*)malloc((size_t)n);
    const int64_t na = oo_activity_mask(phase, phase_id, nx, ny, nz, dir, mask, 0);
    const double avf = n > 0 ? (double)na / (double)n : 0.0;
    for (int q = 0; q < 9; ++q) out[q] = 0.0;
    out[2] = avf; out[7] = (double)na;
    if (na == 0) { out[0] = NAN; free(mask); return 0; }
    double* a = (double*)malloc(sizeof(double) * 7 * (size_t)n);
    double* rhs = (double*)malloc(sizeof(double) * (size_t)n);
    double* x = (double*)calloc((size_t)n, sizeof(double));
    oo_fillmtx(a, rhs, x, phase, mask, nx, ny, nz, dxinv, vlo, vhi, phase_id, dir);
    double relres = 0.0;
#ifdef _OPENMP
    const double t0 = omp_get_wtime();
#endif
    const int it = oo_solve_pcg(a, rhs, x, nx, ny, nz, eps, maxiter, &relres);
#ifdef _OPENMP
    out[8] = omp_get_wtime() - t0;
#endif
    double fin, fout, deff;
    oo_fluxes(x, mask, nx, ny, nz, dir, dx, &fin, &fout, NULL, NULL);
    const int conv = isfinite(relres) && relres >= 0.0 && relres <= eps;
    out[0] = oo_tau(fin, fout, avf, nx, ny, nz, dir, dx, vlo, vhi, conv, &deff);
    out[1] = deff; out[3] = fin; out[4] = fout; out[5] = (double)it; out[6] = relres;
    free(a); free(rhs); free(x); free(mask);
    return 0;
}

/* ===========================================================================
 * Homogenisation cell problem (row f-1): second, independent restatement of
 * effdiff_fillmtx + the periodic Krylov solve + the D_eff tensor sums.
 * =========================================================================== */

/* effdiff_fillmtx, src/props/EffDiffFillMtx.F90:109-258, on the whole periodic box
 * (the mask ghosts are filled periodically, EffectiveDiffusivityHypre.cpp:478).
 * a[7*m+s], slots C,-x,+x,-y,+y,-z,+z (F90:31-37); dx = cell sizes. */
void oo_effdiff_fillmtx(double* a, double* rhs, double* xinit, const int32_t* phase, int32_t phase_id,
                        int nx, int ny, int nz, const double* dx, int dir_k) {
    const double inv2[3] = {1.0 / (dx[0] * dx[0]), 1.0 / (dx[1] * dx[1]), 1.0 / (dx[2] * dx[2])};
    const double inv_2d[3] = {1.0 / (2.0 * dx[0]), 1.0 / (2.0 * dx[1]), 1.0 / (2.0 * dx[2])};
#pragma omp parallel for schedule(static)
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                const int64_t m = IDX(i, j, k);
                double* row = a + 7 * m;
                for (int s = 0; s < 7; ++s) row[s] = 0.0;
                rhs[m] = 0.0;
                xinit[m] = 0.0;
                if (phase[m] != phase_id) { row[0] = 1.0; continue; }             /* F90:124-129 */
                const int64_t nb[6] = {IDX((i + nx - 1) % nx, j, k), IDX((i + 1) % nx, j, k),
                                       IDX(i, (j + ny - 1) % ny, k), IDX(i, (j + 1) % ny, k),
                                       IDX(i, j, (k + nz - 1) % nz), IDX(i, j, (k + 1) % nz)};
                double diag = 0.0, flux = 0.0;
                int act[6];
                for (int s = 0; s < 6; ++s) {
                    const int axis = s / 2;
                    act[s] = phase[nb[s]] == phase_id;
                    if (act[s]) row[1 + s] = -inv2[axis];                           /* F90:152-154 */
                    diag += inv2[axis];                                             /* both branches, F90:153,156 */
                    if (!act[s] && axis == dir_k) flux += (s & 1) ? -(1.0 / dx[axis]) : (1.0 / dx[axis]);
                }
                row[0] = diag;
                const double div = -((double)act[2 * dir_k + 1] - (double)act[2 * dir_k]) * inv_2d[dir_k];   /* F90:226-234 */
                rhs[m] = div + flux;
            }
}

static void matvec7_periodic(const double* a, const double* x, double* y, int nx, int ny, int nz) {
#pragma omp parallel for schedule(static)
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                const int64_t m = IDX(i, j, k);
                const double* r = a + 7 * m;
                double s = r[0] * x[m];
                if (r[1] != 0.0) s += r[1] * x[IDX((i + nx - 1) % nx, j, k)];
                if (r[2] != 0.0) s += r[2] * x[IDX((i + 1) % nx, j, k)];
                if (r[3] != 0.0) s += r[3] * x[IDX(i, (j + ny - 1) % ny, k)];
                if (r[4] != 0.0) s += r[4] * x[IDX(i, (j + 1) % ny, k)];
                if (r[5] != 0.0) s += r[5] * x[IDX(i, j, (k + nz - 1) % nz)];
                if (r[6] != 0.0) s += r[6] * x[IDX(i, j, (k + 1) % nz)];
                y[m] = s;
            }
}

/* EffectiveDiffusivityHypre::solve (src/props/EffectiveDiffusivityHypre.cpp:543-676) restated as
 * Jacobi-PCG on the periodic struct matrix; stop ||b - A x|| <= eps ||b||.  Returns iterations. */
int oo_effdiff_solve(const double* a, const double* rhs, double* x, int nx, int ny, int nz, double eps,
                     int maxiter, double* relres) {
    const int64_t n = (int64_t)nx * ny * nz;
    double* r = (double*)malloc(sizeof(double) * (size_t)n);
    double* z = (double*)malloc(sizeof(double) * (size_t)n);
    double* p = (double*)malloc(sizeof(double) * (size_t)n);
    double* q = (double*)malloc(sizeof(double) * (size_t)n);
    matvec7_periodic(a, x, q, nx, ny, nz);
#pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < n; ++m) r[m] = rhs[m] - q[m];
    const double bnorm = sqrt(dot(rhs, rhs, n));
    double rn = sqrt(dot(r, r, n));
    const double tol = eps * bnorm;
    int it = 0;
    if (bnorm > 0.0 && rn > tol) {
#pragma omp parallel for schedule(static)
        for (int64_t m = 0; m < n; ++m) { z[m] = r[m] / a[7 * m]; p[m] = z[m]; }
        double rz = dot(r, z, n);
        while (it < maxiter) {
            ++it;
            matvec7_periodic(a, p, q, nx, ny, nz);
            const double alpha = rz / dot(p, q, n);
            double rr = 0.0;
#pragma omp parallel for reduction(+ : rr) schedule(static)
            for (int64_t m = 0; m < n; ++m) { x[m] += alpha * p[m]; r[m] -= alpha * q[m]; rr += r[m] * r[m]; }
            rn = sqrt(rr);
            if (!(rn > tol)) break;
            double rzn = 0.0;
#pragma omp parallel for reduction(+ : rzn) schedule(static)
            for (int64_t m = 0; m < n; ++m) { z[m] = r[m] / a[7 * m]; rzn += r[m] * z[m]; }
            const double beta = rzn / rz;
            rz = rzn;
#pragma omp parallel for schedule(static)
            for (int64_t m = 0; m < n; ++m) p[m] = z[m] + beta * p[m];
        }
    }
    if (relres) *relres = bnorm > 0.0 ? rn / bnorm : 0.0;
    free(r); free(z); free(p); free(q);
    return it;
}

/* One column of calculate_Deff_tensor_homogenization (src/props/Diffusion.cpp:97-141): sums over the
 * active cells of the central differences of chi_k (periodic; chi = 0 in the solid). */
void oo_effdiff_gradient_sums(const double* chi, const int32_t* phase, int32_t phase_id, int nx, int ny, int nz,
                              const double* dx, double* sums3, int64_t* n_active) {
    double sx = 0.0, sy = 0.0, sz = 0.0;
    int64_t na = 0;
#pragma omp parallel for reduction(+ : sx, sy, sz, na) schedule(static)
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                const int64_t m = IDX(i, j, k);
                if (phase[m] != phase_id) continue;
                ++na;
                sx += (chi[IDX((i + 1) % nx, j, k)] - chi[IDX((i + nx - 1) % nx, j, k)]) * (1.0 / (2.0 * dx[0]));
                sy += (chi[IDX(i, (j + 1) % ny, k)] - chi[IDX(i, (j + ny - 1) % ny, k)]) * (1.0 / (2.0 * dx[1]));
                sz += (chi[IDX(i, j, (k + 1) % nz)] - chi[IDX(i, j, (k + nz - 1) % nz)]) * (1.0 / (2.0 * dx[2]));
            }
    sums3[0] = sx; sums3[1] = sy; sums3[2] = sz;
    if (n_active) *n_active = na;
}

/* ===========================================================================
 * CPU port of the GPU arm's solver (benchmark baseline, and a third independent
 * solve for the parity tests): PCG preconditioned by one geometric V-cycle on the
 * system with its identity rows eliminated -- the algorithm of
 * openimpala_b200/csrc/oi_solver.cu restated for host cores in plain C + OpenMP,
 * all in fp64.  It stands where HYPRE's FlexGMRES + SMG V-cycle stands in the
 * reference (src/props/TortuosityHypre.cpp:666-691): a multigrid-preconditioned
 * Krylov solve with the reference's stopping rule, so its time is a far fairer CPU
 * baseline than Jacobi-PCG.  Coarse operators: 2x2x2 aggregation, couplings summed
 * over coarse faces and scaled by 1/2 (the rediscretisation-equivalent operator),
 * diagonal = outward couplings + sink terms; smoother: Jacobi with the reciprocals
 * of the fourth-kind Chebyshev roots as weights (degree d0 on level 0, 4 on level 1, dc below), mirrored
 * before / after the coarse correction; coarsest level: 8 sweeps.
 * =========================================================================== */
typedef struct {
    int nx, ny, nz;
    double *cx, *cy, *cz, *dg;      /* coupling to the +x,+y,+z neighbour; diagonal (0 = no unknown) */
    double *x, *b, *t;
} mg_level;

static void mg_free(mg_level* L) { free(L->cx); free(L->cy); free(L->cz); free(L->dg); free(L->x); free(L->b); free(L->t); }

/* Jacobi weights = reciprocals of the roots of the fourth-kind Chebyshev smoother polynomial on (0, 2]
 * (oi_solver.cu cheb_weights, the GPU arm's default), largest root first */
static void mg_cheb(int deg, double lo_frac, double* w) {
    (void)lo_frac;
    for (int k = 1; k <= deg; ++k) {
        const double sn = sin(3.14159265358979323846 * (double)(deg + 1 - k) / (2.0 * deg + 1.0));
        w[k - 1] = 1.0 / (2.0 * sn * sn);
    }
}

/* out = x + w (b - A x) / dg  (res == 1: out = b - A x; res == 2: out = A x);  x == NULL means a zero guess */
static void mg_sweep(const mg_level* L, const double* x, const double* b, double* out, double w, int res) {
    const int nx = L->nx, ny = L->ny, nz = L->nz;
    const int64_t sy = nx, sz = (int64_t)nx * ny;
#pragma omp parallel for schedule(static)
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                const int64_t m = ((int64_t)k * ny + j) * nx + i;
                const double d = L->dg[m];
                if (!(d > 0.0)) { out[m] = 0.0; continue; }
                if (!x) { out[m] = res ? b[m] : w * b[m] / d; continue; }
                double ax = d * x[m];
                if (i + 1 < nx) ax -= L->cx[m] * x[m + 1];
                if (i > 0) ax -= L->cx[m - 1] * x[m - 1];
                if (j + 1 < ny) ax -= L->cy[m] * x[m + sy];
                if (j > 0) ax -= L->cy[m - sy] * x[m - sy];
                if (k + 1 < nz) ax -= L->cz[m] * x[m + sz];
                if (k > 0) ax -= L->cz[m - sz] * x[m - sz];
                out[m] = res == 2 ? ax : (res ? b[m] - ax : x[m] + w * (b[m] - ax) / d);
            }
}

static void mg_coarsen(const mg_level* F, mg_level* C) {
    const int nx = (F->nx + 1) / 2, ny = (F->ny + 1) / 2, nz = (F->nz + 1) / 2;
    const size_t n = (size_t)nx * ny * nz;
    C->nx = nx; C->ny = ny; C->nz = nz;
    C->cx = (double*)calloc(n, sizeof(double)); C->cy = (double*)calloc(n, sizeof(double));
    C->cz = (double*)calloc(n, sizeof(double)); C->dg = (double*)calloc(n, sizeof(double));
    C->x = (double*)calloc(n, sizeof(double)); C->b = (double*)calloc(n, sizeof(double)); C->t = (double*)calloc(n, sizeof(double));
#pragma omp parallel for schedule(static)
    for (int K = 0; K < nz; ++K)
        for (int J = 0; J < ny; ++J)
            for (int I = 0; I < nx; ++I) {
                double sx = 0, sy = 0, sz = 0, sd = 0, in = 0;
                for (int k = 2 * K; k < 2 * K + 2 && k < F->nz; ++k)
                    for (int j = 2 * J; j < 2 * J + 2 && j < F->ny; ++j)
                        for (int i = 2 * I; i < 2 * I + 2 && i < F->nx; ++i) {
                            const int64_t m = ((int64_t)k * F->ny + j) * F->nx + i;
                            sd += F->dg[m];
                            if (i + 1 < F->nx) { if (i + 1 < 2 * I + 2) in += F->cx[m]; else sx += F->cx[m]; }
                            if (j + 1 < F->ny) { if (j + 1 < 2 * J + 2) in += F->cy[m]; else sy += F->cy[m]; }
                            if (k + 1 < F->nz) { if (k + 1 < 2 * K + 2) in += F->cz[m]; else sz += F->cz[m]; }
                        }
                const int64_t M = ((int64_t)K * ny + J) * nx + I;
                C->cx[M] = 0.5 * sx; C->cy[M] = 0.5 * sy; C->cz[M] = 0.5 * sz;
                C->dg[M] = 0.5 * (sd - 2.0 * in);
            }
}

static void mg_restrict(const mg_level* F, const double* res, mg_level* C) {
#pragma omp parallel for schedule(static)
    for (int K = 0; K < C->nz; ++K)
        for (int J = 0; J < C->ny; ++J)
            for (int I = 0; I < C->nx; ++I) {
                double s = 0.0;
                for (int k = 2 * K; k < 2 * K + 2 && k < F->nz; ++k)
                    for (int j = 2 * J; j < 2 * J + 2 && j < F->ny; ++j)
                        for (int i = 2 * I; i < 2 * I + 2 && i < F->nx; ++i)
                            s += res[((int64_t)k * F->ny + j) * F->nx + i];
                C->b[((int64_t)K * C->ny + J) * C->nx + I] = s;
            }
}

static void mg_prolong_add(const mg_level* F, double* x, const mg_level* C) {
#pragma omp parallel for schedule(static)
    for (int k = 0; k < F->nz; ++k)
        for (int j = 0; j < F->ny; ++j)
            for (int i = 0; i < F->nx; ++i) {
                const int64_t m = ((int64_t)k * F->ny + j) * F->nx + i;
                if (F->dg[m] > 0.0) x[m] += C->x[((int64_t)(k >> 1) * C->ny + (j >> 1)) * C->nx + (i >> 1)];
            }
}

/* L[l].x = M^-1 L[l].b by one V-cycle (zero initial guess); w0: level-0 weights, wm: levels below, wc: coarsest */
static const double* g_w1 = NULL;   /* level-1 weights (degree g_d1), set by oo_solve_mgpcg */
static int g_d1 = 0;
static void mg_vcycle(mg_level* L, int l, int nl, const double* w0, int d0, const double* wm, int dm, const double* wc, int dc) {
    mg_level* A = &L[l];
    const int last = (l + 1 == nl);
    const double* w = last ? wc : (l == 0 ? w0 : (l == 1 && g_w1 ? g_w1 : wm));
    const int deg = last ? dc : (l == 0 ? d0 : (l == 1 && g_w1 ? g_d1 : dm));
    double *cur = A->x, *oth = A->t, *tmp;
    mg_sweep(A, NULL, A->b, cur, w[0], 0);
    for (int s = 1; s < deg; ++s) { mg_sweep(A, cur, A->b, oth, w[s], 0); tmp = cur; cur = oth; oth = tmp; }
    if (!last) {
        mg_sweep(A, cur, A->b, oth, 0.0, 1);
        mg_restrict(A, oth, &L[l + 1]);
        mg_vcycle(L, l + 1, nl, w0, d0, wm, dm, wc, dc);
        mg_prolong_add(A, cur, &L[l + 1]);
        for (int s = 0; s < deg; ++s) { mg_sweep(A, cur, A->b, oth, w[deg - 1 - s], 0); tmp = cur; cur = oth; oth = tmp; }
    }
    if (cur != A->x) { A->t = A->x; A->x = cur; }
}

/* Solve the eliminated system of the tortuosity problem by MG-PCG.  mask: activity mask (oo_activity_mask);
 * x: out, the full potential field (Dirichlet values on the active plane cells, 0 on inactive cells).
 * Stop rule and return values as oo_solve_pcg.  d0 / dc: smoothing degree on level 0 / below (0 = 5 / 8). */
int oo_solve_mgpcg(const uint8_t* mask, int nx, int ny, int nz, int dir, double vlo, double vhi, double eps,
                   int maxiter, int d0, int dc, double* x, double* relres) {
    const int64_t n = (int64_t)nx * ny * nz, sy = nx, sz = (int64_t)nx * ny;
    const int nd = dir == 0 ? nx : (dir == 1 ? ny : nz);
    if (d0 <= 0) d0 = 5;
    if (dc <= 0) dc = 8;
    if (d0 > 16) d0 = 16;
    if (dc > 16) dc = 16;
    static const double lo_tab[] = {0.4, 0.4, 0.25, 0.2, 0.10, 0.05, 0.05, 0.05, 0.05};     /* as oi_solver.cu */
    double w0[16], wm[16], wc[8];
    static double w1[16];
    mg_cheb(4, 0.1, w1);                 /* MG level 1: degree 4, as the GPU arm's default */
    g_w1 = w1; g_d1 = 4;
    mg_cheb(d0, d0 <= 8 ? lo_tab[d0] : 0.07, w0);
    mg_cheb(dc, dc <= 8 ? lo_tab[dc] : 0.07, wm);
    mg_cheb(8, 0.05, wc);
    mg_level L[20];
    int nl = 1;
    mg_level* F = &L[0];
    F->nx = nx; F->ny = ny; F->nz = nz;
    F->cx = (double*)calloc((size_t)n, sizeof(double)); F->cy = (double*)calloc((size_t)n, sizeof(double));
    F->cz = (double*)calloc((size_t)n, sizeof(double)); F->dg = (double*)calloc((size_t)n, sizeof(double));
    F->x = (double*)calloc((size_t)n, sizeof(double)); F->b = (double*)calloc((size_t)n, sizeof(double)); F->t = (double*)calloc((size_t)n, sizeof(double));
    double* bvec = (double*)calloc((size_t)n, sizeof(double));       /* rhs of the eliminated system */
    int64_t n_in = 0, n_out = 0;
    /* level 0 from the mask (tortuosity_fillmtx with its identity rows eliminated, F90:111-228) */
#pragma omp parallel for schedule(static) reduction(+ : n_in, n_out)
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                const int64_t m = ((int64_t)k * ny + j) * nx + i;
                x[m] = 0.0;
                if (!mask[m]) continue;
                const int d = dir == 0 ? i : (dir == 1 ? j : k);
                if (d == 0) { x[m] = vlo; ++n_in; continue; }
                if (d == nd - 1) { x[m] = vhi; ++n_out; continue; }
                x[m] = vlo + (vhi - vlo) * (double)d * (nd > 1 ? 1.0 / (double)(nd - 1) : 0.0);   /* ramp, F90:233-262 */
                double diag = 0.0, rhs = 0.0;
                const int64_t nb[6] = {m - 1, m + 1, m - sy, m + sy, m - sz, m + sz};
                const int ok[6] = {i > 0, i + 1 < nx, j > 0, j + 1 < ny, k > 0, k + 1 < nz};
                const int dn[6] = {dir == 0 ? d - 1 : d, dir == 0 ? d + 1 : d, dir == 1 ? d - 1 : d, dir == 1 ? d + 1 : d,
                                   dir == 2 ? d - 1 : d, dir == 2 ? d + 1 : d};
                for (int s = 0; s < 6; ++s) {
                    if (!ok[s] || !mask[nb[s]]) continue;
                    diag += 1.0;
                    if (dn[s] == 0) rhs += vlo;                        /* Dirichlet neighbour: to the rhs */
                    else if (dn[s] == nd - 1) rhs += vhi;
                    else if (s == 1) F->cx[m] = 1.0;
                    else if (s == 3) F->cy[m] = 1.0;
                    else if (s == 5) F->cz[m] = 1.0;
                }
                F->dg[m] = diag;
                bvec[m] = rhs;
            }
    while (nl < 20 && (L[nl - 1].nx >= 3 || L[nl - 1].ny >= 3 || L[nl - 1].nz >= 3) &&
           (int64_t)L[nl - 1].nx * L[nl - 1].ny * L[nl - 1].nz > 64) {
        mg_coarsen(&L[nl - 1], &L[nl]);
        ++nl;
    }
    double* r = (double*)calloc((size_t)n, sizeof(double));
    double* p = (double*)calloc((size_t)n, sizeof(double));
    double* q = (double*)calloc((size_t)n, sizeof(double));
    double* u = (double*)calloc((size_t)n, sizeof(double));            /* unknown part of x */
#pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < n; ++m) u[m] = F->dg[m] > 0.0 ? x[m] : 0.0;
    mg_sweep(F, u, bvec, r, 0.0, 1);                                  /* r = b - A u */
    const double bnorm = sqrt((double)n_in * vlo * vlo + (double)n_out * vhi * vhi);
    double rn = sqrt(dot(r, r, n));
    const double den = bnorm > 0.0 ? bnorm : rn;
    const double tol = eps * den;
    int it = 0;
    if (rn > tol) {
        memcpy(F->b, r, sizeof(double) * (size_t)n);
        mg_vcycle(L, 0, nl, w0, d0, wm, dc, wc, 8);
        memcpy(p, F->x, sizeof(double) * (size_t)n);
        double rz = dot(r, p, n);
        while (it < maxiter) {
            ++it;
            mg_sweep(F, p, NULL, q, 0.0, 2);
            const double alpha = rz / dot(p, q, n);
            double rr = 0.0;
#pragma omp parallel for reduction(+ : rr) schedule(static)
            for (int64_t m = 0; m < n; ++m) { u[m] += alpha * p[m]; r[m] -= alpha * q[m]; rr += r[m] * r[m]; }
            rn = sqrt(rr);
            if (!(rn > tol)) break;
            memcpy(F->b, r, sizeof(double) * (size_t)n);
            mg_vcycle(L, 0, nl, w0, d0, wm, dc, wc, 8);
            const double rzn = dot(r, F->x, n);
            const double beta = rzn / rz;
            rz = rzn;
            const double* z = F->x;
#pragma omp parallel for schedule(static)
            for (int64_t m = 0; m < n; ++m) p[m] = z[m] + beta * p[m];
        }
    }
#pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < n; ++m) if (F->dg[m] > 0.0) x[m] = u[m];
    if (relres) *relres = den > 0.0 ? rn / den : 0.0;
    for (int l = 0; l < nl; ++l) mg_free(&L[l]);
    free(bvec); free(r); free(p); free(q); free(u);
    return it;
}

/* oo_tortuosity with the MG-PCG solve (same mask, flux and tau tail) */
int oo_tortuosity_mg(const int32_t* phase, int nx, int ny, int nz, int32_t phase_id, int dir,
                     double vlo, double vhi, double eps, int maxiter, int d0, int dc, double* out) {
    const int64_t n = (int64_t)nx * ny * nz;
    const double dx[3] = {1.0, 1.0, 1.0};
    uint8_t* mask = (uint8_t*)malloc((size_t)n);
#ifdef _OPENMP
    const double tm = omp_get_wtime();
#endif
    const int64_t na = oo_activity_mask(phase, phase_id, nx, ny, nz, dir, mask, 0);
    const double avf = n > 0 ? (double)na / (double)n : 0.0;
    for (int q = 0; q < 10; ++q) out[q] = 0.0;
    out[2] = avf; out[7] = (double)na;
    if (na == 0) { out[0] = NAN; free(mask); return 0; }
    double* x = (double*)calloc((size_t)n, sizeof(double));
    double relres = 0.0;
#ifdef _OPENMP
    const double t0 = omp_get_wtime();
    out[9] = t0 - tm;                                                   /* mask seconds */
#endif
    const int it = oo_solve_mgpcg(mask, nx, ny, nz, dir, vlo, vhi, eps, maxiter, d0, dc, x, &relres);
#ifdef _OPENMP
    out[8] = omp_get_wtime() - t0;
#endif
    double fin, fout, deff;
    oo_fluxes(x, mask, nx, ny, nz, dir, dx, &fin, &fout, NULL, NULL);
    const int conv = isfinite(relres) && relres >= 0.0 && relres <= eps;
    out[0] = oo_tau(fin, fout, avf, nx, ny, nz, dir, dx, vlo, vhi, conv, &deff);
    out[1] = deff; out[3] = fin; out[4] = fout; out[5] = (double)it; out[6] = relres;
    free(x); free(mask);
    return 0;
}
