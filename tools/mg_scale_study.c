/* CPU study: per-level scaling of the aggregation coarse operators of the MG-PCG solve.
 *
 * The shipped hierarchy scales every aggregated operator by 1/2 (the over-correction that makes piecewise-constant
 * aggregation equivalent to the rediscretised operator on a uniform grid).  On a porous geometry the energy of the
 * staircase interpolant of a smooth mode exceeds the energy of the mode by MORE than 2 once aggregates are larger than
 * the pore features (the aggregated faces count open connections, the true conductance includes the tortuosity inside
 * the aggregates), so the deep levels under-correct.  This program measures that: it solves the tortuosity system of a
 * given activity mask with (a) the 1/2 hierarchy, (b) per-level scalars s_l = e^T A_l e / e_c^T (P^T A_l P) e_c matched
 * on a smooth test vector e (the error after a few iterations, or the first preconditioned residual).
 *
 *   gcc -O3 -fopenmp -o /tmp/mg_scale_study tools/mg_scale_study.c -lm
 *   /tmp/mg_scale_study mask.u8 n [threads]
 * The solver is the one of oracle/oi_oracle.c (oo_solve_mgpcg) with the scaling made a parameter.
 *
 * RESULT (128^3 packing, R 12, 6 levels): the hypothesis does not hold.  1/2 everywhere: 11 iterations.  Scalars
 * matched on the error after 3 iterations come out ABOVE 1/2 (0.75 0.60 0.60 0.59 0.74) and cost 14 iterations, matched
 * on the error of the ramp (0.54 .. 0.65) 12, on the preconditioned residual 16; 1/2 on level 1 and 0.45 / 0.42 / 0.40 /
 * 0.35 below: 12 / 12 / 13 / 15.  The 1/2 of the rediscretisation argument is the optimum on this geometry as well.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

typedef struct {
    int nx, ny, nz;
    double *cx, *cy, *cz, *dg, *x, *b, *t;
} level;

static double* dalloc(size_t n) { return (double*)calloc(n, sizeof(double)); }
static void lfree(level* L) { free(L->cx); free(L->cy); free(L->cz); free(L->dg); free(L->x); free(L->b); free(L->t); }
static size_t lcells(const level* L) { return (size_t)L->nx * L->ny * L->nz; }

static double dot(const double* a, const double* b, int64_t n) {
    double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

static void cheb4(int deg, double* w) {
    for (int k = 1; k <= deg; ++k) {
        const double sn = sin(3.14159265358979323846 * (double)(deg + 1 - k) / (2.0 * deg + 1.0));
        w[k - 1] = 1.0 / (2.0 * sn * sn);
    }
}

/* res 0: out = x + w (b - A x)/d; 1: out = b - A x; 2: out = A x; x == NULL: zero guess */
static void sweep(const level* L, const double* x, const double* b, double* out, double w, int res) {
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

/* aggregated operator, scaled by s (all couplings and the diagonal) */
static void coarsen(const level* F, level* C, double s) {
    const int nx = (F->nx + 1) / 2, ny = (F->ny + 1) / 2, nz = (F->nz + 1) / 2;
    const size_t n = (size_t)nx * ny * nz;
    C->nx = nx; C->ny = ny; C->nz = nz;
    C->cx = dalloc(n); C->cy = dalloc(n); C->cz = dalloc(n); C->dg = dalloc(n);
    C->x = dalloc(n); C->b = dalloc(n); C->t = dalloc(n);
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
                C->cx[M] = s * sx; C->cy[M] = s * sy; C->cz[M] = s * sz;
                C->dg[M] = s * (sd - 2.0 * in);
            }
}

static void rescale(level* C, double f) {
    const size_t n = lcells(C);
#pragma omp parallel for schedule(static)
    for (size_t m = 0; m < n; ++m) { C->cx[m] *= f; C->cy[m] *= f; C->cz[m] *= f; C->dg[m] *= f; }
}

static void restrict_sum(const level* F, const double* res, level* C, double* out) {
#pragma omp parallel for schedule(static)
    for (int K = 0; K < C->nz; ++K)
        for (int J = 0; J < C->ny; ++J)
            for (int I = 0; I < C->nx; ++I) {
                double s = 0.0;
                for (int k = 2 * K; k < 2 * K + 2 && k < F->nz; ++k)
                    for (int j = 2 * J; j < 2 * J + 2 && j < F->ny; ++j)
                        for (int i = 2 * I; i < 2 * I + 2 && i < F->nx; ++i)
                            s += res[((int64_t)k * F->ny + j) * F->nx + i];
                out[((int64_t)K * C->ny + J) * C->nx + I] = s;
            }
}

/* coarse representative of a fine vector: mean over the children that are unknowns */
static void restrict_mean(const level* F, const double* v, level* C, double* out) {
#pragma omp parallel for schedule(static)
    for (int K = 0; K < C->nz; ++K)
        for (int J = 0; J < C->ny; ++J)
            for (int I = 0; I < C->nx; ++I) {
                double s = 0.0; int c = 0;
                for (int k = 2 * K; k < 2 * K + 2 && k < F->nz; ++k)
                    for (int j = 2 * J; j < 2 * J + 2 && j < F->ny; ++j)
                        for (int i = 2 * I; i < 2 * I + 2 && i < F->nx; ++i) {
                            const int64_t m = ((int64_t)k * F->ny + j) * F->nx + i;
                            if (F->dg[m] > 0.0) { s += v[m]; ++c; }
                        }
                out[((int64_t)K * C->ny + J) * C->nx + I] = c ? s / c : 0.0;
            }
}

static void prolong_add(const level* F, double* x, const level* C) {
#pragma omp parallel for schedule(static)
    for (int k = 0; k < F->nz; ++k)
        for (int j = 0; j < F->ny; ++j)
            for (int i = 0; i < F->nx; ++i) {
                const int64_t m = ((int64_t)k * F->ny + j) * F->nx + i;
                if (F->dg[m] > 0.0) x[m] += C->x[((int64_t)(k >> 1) * C->ny + (j >> 1)) * C->nx + (i >> 1)];
            }
}

static double W0[16], W1[16], WM[16], WC[16];
static int D0 = 5, D1 = 4, DM = 8, DC = 8;

static void vcycle(level* L, int l, int nl) {
    level* A = &L[l];
    const int last = (l + 1 == nl);
    const double* w = last ? WC : (l == 0 ? W0 : (l == 1 ? W1 : WM));
    const int deg = last ? DC : (l == 0 ? D0 : (l == 1 ? D1 : DM));
    double *cur = A->x, *oth = A->t, *tmp;
    sweep(A, NULL, A->b, cur, w[0], 0);
    for (int s = 1; s < deg; ++s) { sweep(A, cur, A->b, oth, w[s], 0); tmp = cur; cur = oth; oth = tmp; }
    if (!last) {
        sweep(A, cur, A->b, oth, 0.0, 1);
        restrict_sum(A, oth, &L[l + 1], L[l + 1].b);
        vcycle(L, l + 1, nl);
        prolong_add(A, cur, &L[l + 1]);
        for (int s = 0; s < deg; ++s) { sweep(A, cur, A->b, oth, w[deg - 1 - s], 0); tmp = cur; cur = oth; oth = tmp; }
    }
    if (cur != A->x) { A->t = A->x; A->x = cur; }
}

static double energy(const level* L, const double* v, double* tmp) {
    sweep(L, v, NULL, tmp, 0.0, 2);
    return dot(v, tmp, (int64_t)lcells(L));
}

/* level 0 of the eliminated tortuosity system in z from the activity mask; u0 = ramp, bvec = rhs */
static void level0(const uint8_t* mask, int n, level* F, double* u0, double* bvec, int64_t* n_in, int64_t* n_out) {
    const int nx = n, ny = n, nz = n;
    const size_t N = (size_t)n * n * n;
    const int64_t sy = nx, sz = (int64_t)nx * ny;
    F->nx = nx; F->ny = ny; F->nz = nz;
    F->cx = dalloc(N); F->cy = dalloc(N); F->cz = dalloc(N); F->dg = dalloc(N);
    F->x = dalloc(N); F->b = dalloc(N); F->t = dalloc(N);
    int64_t ni = 0, no = 0;
    const double vlo = 0.0, vhi = 1.0;
#pragma omp parallel for schedule(static) reduction(+ : ni, no)
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                const int64_t m = ((int64_t)k * ny + j) * nx + i;
                u0[m] = 0.0; bvec[m] = 0.0;
                if (!mask[m]) continue;
                if (k == 0) { ++ni; continue; }
                if (k == nz - 1) { ++no; continue; }
                u0[m] = vlo + (vhi - vlo) * (double)k / (double)(nz - 1);
                double diag = 0.0, rhs = 0.0;
                const int64_t nb[6] = {m - 1, m + 1, m - sy, m + sy, m - sz, m + sz};
                const int ok[6] = {i > 0, i + 1 < nx, j > 0, j + 1 < ny, k > 0, k + 1 < nz};
                const int dn[6] = {k, k, k, k, k - 1, k + 1};
                for (int s = 0; s < 6; ++s) {
                    if (!ok[s] || !mask[nb[s]]) continue;
                    diag += 1.0;
                    if (dn[s] == 0) rhs += vlo;
                    else if (dn[s] == nz - 1) rhs += vhi;
                    else if (s == 1) F->cx[m] = 1.0;
                    else if (s == 3) F->cy[m] = 1.0;
                    else if (s == 5) F->cz[m] = 1.0;
                }
                F->dg[m] = diag;
                bvec[m] = rhs;
            }
    *n_in = ni; *n_out = no;
}

static int build(level* L, const double* scale) {
    int nl = 1;
    while (nl < 20 && (L[nl - 1].nx >= 3 || L[nl - 1].ny >= 3 || L[nl - 1].nz >= 3) && lcells(&L[nl - 1]) > 64) {
        coarsen(&L[nl - 1], &L[nl], scale ? scale[nl] : 0.5);
        ++nl;
    }
    return nl;
}

/* PCG; snap_it >= 0: copy u after that many iterations into snap_u, and the preconditioned residual of that iteration into snap_z */
static int pcg(level* L, int nl, const double* u0, const double* bvec, double bnorm, double eps, int maxiter, double* u,
               int snap_it, double* snap_u, double* snap_z, double* relres) {
    level* F = &L[0];
    const int64_t n = (int64_t)lcells(F);
    double *r = dalloc(n), *p = dalloc(n), *q = dalloc(n);
    memcpy(u, u0, sizeof(double) * n);
    sweep(F, u, bvec, r, 0.0, 1);
    double rn = sqrt(dot(r, r, n));
    const double den = bnorm > 0.0 ? bnorm : rn, tol = eps * den;
    int it = 0;
    if (rn > tol) {
        memcpy(F->b, r, sizeof(double) * n);
        vcycle(L, 0, nl);
        memcpy(p, F->x, sizeof(double) * n);
        if (snap_it == 0) { if (snap_u) memcpy(snap_u, u, sizeof(double) * n); if (snap_z) memcpy(snap_z, F->x, sizeof(double) * n); }
        double rz = dot(r, p, n);
        while (it < maxiter) {
            ++it;
            sweep(F, p, NULL, q, 0.0, 2);
            const double alpha = rz / dot(p, q, n);
            double rr = 0.0;
#pragma omp parallel for reduction(+ : rr) schedule(static)
            for (int64_t m = 0; m < n; ++m) { u[m] += alpha * p[m]; r[m] -= alpha * q[m]; rr += r[m] * r[m]; }
            rn = sqrt(rr);
            if (!(rn > tol)) break;
            memcpy(F->b, r, sizeof(double) * n);
            vcycle(L, 0, nl);
            if (snap_it == it) { if (snap_u) memcpy(snap_u, u, sizeof(double) * n); if (snap_z) memcpy(snap_z, F->x, sizeof(double) * n); }
            const double rzn = dot(r, F->x, n);
            const double beta = rzn / rz;
            rz = rzn;
            const double* z = F->x;
#pragma omp parallel for schedule(static)
            for (int64_t m = 0; m < n; ++m) p[m] = z[m] + beta * p[m];
        }
    }
    if (relres) *relres = rn / den;
    free(r); free(p); free(q);
    return it;
}

/* per-level scalars matched on the test vector e (level 0): s_l = E_{l-1}(e_{l-1}) / E_l^{unscaled}(e_l); the hierarchy
 * L[1..] is rebuilt level by level with them */
static int build_matched(level* L, const double* e, double* scale, double lo, double hi) {
    int nl = 1;
    double* v = dalloc(lcells(&L[0]));
    memcpy(v, e, sizeof(double) * lcells(&L[0]));
    while (nl < 20 && (L[nl - 1].nx >= 3 || L[nl - 1].ny >= 3 || L[nl - 1].nz >= 3) && lcells(&L[nl - 1]) > 64) {
        level* F = &L[nl - 1];
        level* C = &L[nl];
        coarsen(F, C, 1.0);
        double* tmp = dalloc(lcells(F));
        const double ef = energy(F, v, tmp);
        free(tmp);
        double* vc = dalloc(lcells(C));
        restrict_mean(F, v, C, vc);
        double* tc = dalloc(lcells(C));
        const double ec = energy(C, vc, tc);
        free(tc);
        double s = ec > 0.0 ? ef / ec : 0.5;
        if (!(s > lo)) s = lo;
        if (s > hi) s = hi;
        scale[nl] = s;
        rescale(C, s);
        free(v); v = vc;
        ++nl;
    }
    free(v);
    return nl;
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s mask.u8 n [threads]\n", argv[0]); return 2; }
    const int n = atoi(argv[2]);
    if (argc > 3) omp_set_num_threads(atoi(argv[3]));
    const size_t N = (size_t)n * n * n;
    uint8_t* mask = (uint8_t*)malloc(N);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(mask, 1, N, f) != N) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    fclose(f);
    cheb4(D0, W0); cheb4(D1, W1); cheb4(DM, WM); cheb4(DC, WC);
    level L[20];
    double *u0 = dalloc(N), *bvec = dalloc(N), *u = dalloc(N), *ustar = dalloc(N), *uk = dalloc(N), *zk = dalloc(N), *e = dalloc(N);
    int64_t n_in, n_out;
    level0(mask, n, &L[0], u0, bvec, &n_in, &n_out);
    const double bnorm = sqrt((double)n_out);
    double relres;
    /* (a) the shipped hierarchy */
    int nl = build(L, NULL);
    double t0 = omp_get_wtime();
    int it = pcg(L, nl, u0, bvec, bnorm, 1e-9, 200, u, -1, NULL, NULL, &relres);
    printf("n %d levels %d  half: %d iterations (relres %.2e, %.1f s)\n", n, nl, it, relres, omp_get_wtime() - t0);
    fflush(stdout);
    /* reference solution, and snapshots after 3 iterations */
    pcg(L, nl, u0, bvec, bnorm, 1e-13, 400, ustar, 3, uk, zk, &relres);
    for (int l = 1; l < nl; ++l) lfree(&L[l]);
    /* (b) matched on the true error after 3 iterations */
    double sc[20];
    for (size_t m = 0; m < N; ++m) e[m] = ustar[m] - uk[m];
    nl = build_matched(L, e, sc, 0.1, 1.0);
    printf("scales (error after 3 iterations):");
    for (int l = 1; l < nl; ++l) printf(" %.3f", sc[l]);
    it = pcg(L, nl, u0, bvec, bnorm, 1e-9, 200, u, -1, NULL, NULL, &relres);
    printf("  -> %d iterations (relres %.2e)\n", it, relres);
    fflush(stdout);
    for (int l = 1; l < nl; ++l) lfree(&L[l]);
    /* (c) matched on the true error of the initial guess */
    for (size_t m = 0; m < N; ++m) e[m] = ustar[m] - u0[m];
    nl = build_matched(L, e, sc, 0.1, 1.0);
    printf("scales (error of the ramp):");
    for (int l = 1; l < nl; ++l) printf(" %.3f", sc[l]);
    it = pcg(L, nl, u0, bvec, bnorm, 1e-9, 200, u, -1, NULL, NULL, &relres);
    printf("  -> %d iterations (relres %.2e)\n", it, relres);
    fflush(stdout);
    for (int l = 1; l < nl; ++l) lfree(&L[l]);
    /* (d) matched on the preconditioned residual of iteration 3 (available in practice) */
    nl = build_matched(L, zk, sc, 0.1, 1.0);
    printf("scales (z of iteration 3):");
    for (int l = 1; l < nl; ++l) printf(" %.3f", sc[l]);
    it = pcg(L, nl, u0, bvec, bnorm, 1e-9, 200, u, -1, NULL, NULL, &relres);
    printf("  -> %d iterations (relres %.2e)\n", it, relres);
    fflush(stdout);
    for (int l = 1; l < nl; ++l) lfree(&L[l]);
    /* (e) fixed tables: 1/2 on level 1, g below */
    const double gs[] = {0.45, 0.42, 0.40, 0.35};
    for (int q = 0; q < 4; ++q) {
        for (int l = 0; l < 20; ++l) sc[l] = l <= 1 ? 0.5 : gs[q];
        nl = build(L, sc);
        it = pcg(L, nl, u0, bvec, bnorm, 1e-9, 200, u, -1, NULL, NULL, &relres);
        printf("fixed 0.5 / %.2f below level 1 -> %d iterations\n", gs[q], it);
        fflush(stdout);
        for (int l = 1; l < nl; ++l) lfree(&L[l]);
    }
    return 0;
}
