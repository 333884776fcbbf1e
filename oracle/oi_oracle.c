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
