/* osqp_port.c -- CPU port of the reference's per-cycle MPC path.  TEST INFRASTRUCTURE ONLY
 * (see oracle/__init__.py): the checker and the CPU baseline of bench.py, never the product.
 *
 * What it restates, per robot and per cycle, in the reference's own formulation:
 *   gait.py:26-37                 contact table
 *   com_trajectory.py:221-286     continuous + ZOH-discrete dynamics (closed form, A_c^2 = 0)
 *   centroidal_mpc.py:122-176     variable bounds
 *   centroidal_mpc.py:235-303     g, A = [A_eq ; F], lba/uba   (sparse QP, 24N variables)
 *   centroidal_mpc.py:324-359     friction rows
 *   centroidal_mpc.py:98          CasADi conic -> OSQP: A_osqp = [I ; A], l = [lbx ; lba], u = [ubx ; uba]
 * The OSQP algorithm is restated from the published paper (Stellato et al. 2020) with the option
 * values of centroidal_mpc.py:24-35; oracle/osqp_ref.py is the NumPy twin this file is checked
 * against.  Parity unpinned against a real OSQP binary (none is installable here).
 *
 * Linear system: OSQP factors the quasi-definite KKT matrix with QDLDL; here the equivalent reduced
 * system (P + sigma I + A' diag(rho) A) x = rhs is factored by a banded Cholesky in the stage order
 * (u_k, x_{k+1}), half bandwidth 35 -- the same x~ and z~ = A x~ in exact arithmetic.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

#define NX 12
#define NU 12
#define BW 35
#define OSQP_INFTY 1e30
#define MIN_SCALING 1e-4
#define MAX_SCALING 1e4
#define RHO_MIN 1e-6
#define RHO_MAX 1e6
#define RHO_TOL 1e-4
#define RHO_EQ_OVER_RHO_INEQ 1e3

typedef struct {
    double Q[12], R[12];
    double mu, fz_min;
    double eps_abs, eps_rel, rho, sigma, alpha, adaptive_rho_tolerance;
    int max_iter, check_termination, adaptive_rho_interval, scaling;
} port_opts;

void port_default_opts(port_opts* o) {
    static const double Q[12] = {1, 1, 50, 10, 20, 1, 2, 2, 1, 1, 1, 1}; /* centroidal_mpc.py:12 */
    for (int i = 0; i < 12; ++i) { o->Q[i] = Q[i]; o->R[i] = 1e-5; }       /* :13 */
    o->mu = 0.8; o->fz_min = 10.0;                                         /* :15, :127 */
    o->eps_abs = 1e-4; o->eps_rel = 1e-4; o->max_iter = 1000;              /* :25-27 */
    o->check_termination = 10; o->adaptive_rho_interval = 25; o->scaling = 5; /* :31-33 */
    o->rho = 0.1; o->sigma = 1e-6; o->alpha = 1.6; o->adaptive_rho_tolerance = 5.0; /* OSQP defaults */
}
int port_opts_size(void) { return (int)sizeof(port_opts); }

/* ---- gait.py:26-37 ---------------------------------------------------------------------- */
static void contact_table(double t0, double dt, int N, double hz, double duty, const double* off, int* ct /*4xN*/) {
    const double T = 1 / hz;
    for (int leg = 0; leg < 4; ++leg)
        for (int k = 0; k < N; ++k) {
            volatile double t = t0 + (double)k * dt;
            t = t + dt / 2;
            volatile double ph = off[leg] + t / T;
            double r = fmod(ph, 1.0);
            if (r < 0.0) r += 1.0;
            ct[leg * N + k] = r < duty ? 1 : 0;
        }
}

/* ---- com_trajectory.py:221-286 (closed-form ZOH) ------------------------------------------ */
static void dynamics(int N, const double* x_ref, const double* r_foot, const double* I, double m, double dt,
                     double* Ad /*144*/, double* Bd /*N*144*/, double* gd /*12*/) {
    double s = 0;
    for (int i = 0; i < N; ++i) s += x_ref[5 * N + i];
    const double yaw = s / N, cy = cos(yaw), sy = sin(yaw);
    const double a = I[0], b = I[1], c = I[2], d = I[3], e = I[4], f = I[5], g = I[6], h = I[7], k = I[8];
    const double A_ = e * k - f * h, B_ = -(d * k - f * g), C_ = d * h - e * g;
    const double id = 1.0 / (a * A_ + b * B_ + c * C_);
    const double Ii[9] = {A_ * id, -(b * k - c * h) * id, (b * f - c * e) * id,
                          B_ * id, (a * k - c * g) * id, -(a * f - c * d) * id,
                          C_ * id, -(a * h - b * g) * id, (a * e - b * d) * id};
    memset(Ad, 0, 144 * sizeof(double));
    for (int i = 0; i < 12; ++i) Ad[i * 12 + i] = 1.0;
    for (int i = 0; i < 3; ++i) Ad[i * 12 + 6 + i] = dt;
    const double RzT[9] = {cy, sy, 0, -sy, cy, 0, 0, 0, 1};
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Ad[(3 + i) * 12 + 9 + j] = dt * RzT[i * 3 + j];
    memset(gd, 0, 12 * sizeof(double));
    gd[2] = -9.81 * (dt * dt / 2.0);
    gd[8] = -9.81 * dt;
    const double hh = dt * dt / 2.0;
    memset(Bd, 0, (size_t)N * 144 * sizeof(double));
    for (int kk = 0; kk < N; ++kk)
        for (int leg = 0; leg < 4; ++leg) {
            const double r[3] = {r_foot[(leg * 3 + 0) * N + kk], r_foot[(leg * 3 + 1) * N + kk], r_foot[(leg * 3 + 2) * N + kk]};
            const double S[9] = {0, -r[2], r[1], r[2], 0, -r[0], -r[1], r[0], 0};
            double W[9];
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
                W[i * 3 + j] = Ii[i * 3] * S[j] + Ii[i * 3 + 1] * S[3 + j] + Ii[i * 3 + 2] * S[6 + j];
            double* Bk = Bd + (size_t)kk * 144;
            for (int j = 0; j < 3; ++j) {
                const int col = 3 * leg + j;
                Bk[j * 12 + col] = hh / m;
                Bk[(6 + j) * 12 + col] = dt / m;
                for (int i = 0; i < 3; ++i) {
                    Bk[(9 + i) * 12 + col] = dt * W[i * 3 + j];
                    Bk[(3 + i) * 12 + col] = hh * (RzT[i * 3] * W[j] + RzT[i * 3 + 1] * W[3 + j] + RzT[i * 3 + 2] * W[6 + j]);
                }
            }
        }
}

/* ---- CSR matrix with the structural pattern CasADi hands OSQP ------------------------------- */
typedef struct { int m, n, nnz; int* rp; int* ci; double* v; } csr;

static void csr_mv(const csr* A, const double* x, double* y) {
    for (int r = 0; r < A->m; ++r) {
        double s = 0;
        for (int p = A->rp[r]; p < A->rp[r + 1]; ++p) s += A->v[p] * x[A->ci[p]];
        y[r] = s;
    }
}
static void csr_mtv(const csr* A, const double* y, double* x) {
    memset(x, 0, A->n * sizeof(double));
    for (int r = 0; r < A->m; ++r) {
        const double yr = y[r];
        if (yr == 0.0) continue;
        for (int p = A->rp[r]; p < A->rp[r + 1]; ++p) x[A->ci[p]] += A->v[p] * yr;
    }
}

static double limit_scaling(double v) { return v < MIN_SCALING ? 1.0 : (v > MAX_SCALING ? MAX_SCALING : v); }
static double vmax_abs(const double* v, int n) { double m = 0; for (int i = 0; i < n; ++i) { double a = fabs(v[i]); if (a > m) m = a; } return m; }

/* stage ordering (u_k, x_{k+1}) of the 24N variables -> banded K */
static inline int perm_of(int c, int N) {
    if (c < 12 * N) { int k = c / 12; return 24 * k + 12 + (c - 12 * k); }
    c -= 12 * N;
    { int k = c / 12; return 24 * k + (c - 12 * k); }
}

typedef struct {
    int N, n, m;
    csr A;
    double *Pd, *q, *l, *u, *D, *E, *rv, *Kb, *x, *y, *z, *xt, *zt, *rhs, *t1, *t2, *t3, *ncol;
    int* perm;
    int* ct;
    double *Ad, *Bd, *gd;
} work;

static work* work_new(int N) {
    work* w = (work*)calloc(1, sizeof(work));
    w->N = N; w->n = 24 * N; w->m = 52 * N;
    const int n = w->n, m = w->m;
    const int nnz = n + (12 * N + (N - 1) * 144 + N * 144) + 32 * N;
    w->A.m = m; w->A.n = n; w->A.nnz = nnz;
    w->A.rp = (int*)malloc((m + 1) * sizeof(int));
    w->A.ci = (int*)malloc(nnz * sizeof(int));
    w->A.v = (double*)malloc(nnz * sizeof(double));
#define DA(name, cnt) w->name = (double*)malloc((size_t)(cnt) * sizeof(double))
    DA(Pd, n); DA(q, n); DA(l, m); DA(u, m); DA(D, n); DA(E, m); DA(rv, m); DA(Kb, (size_t)n * (BW + 1));
    DA(x, n); DA(y, m); DA(z, m); DA(xt, n); DA(zt, m); DA(rhs, n); DA(t1, m > n ? m : n); DA(t2, m > n ? m : n);
    DA(t3, m > n ? m : n); DA(ncol, n); DA(Ad, 144); DA(Bd, (size_t)N * 144); DA(gd, 12);
#undef DA
    w->perm = (int*)malloc(n * sizeof(int));
    for (int c = 0; c < n; ++c) w->perm[c] = perm_of(c, N);
    w->ct = (int*)malloc(4 * N * sizeof(int));
    return w;
}
static void work_free(work* w) {
    free(w->A.rp); free(w->A.ci); free(w->A.v);
    free(w->Pd); free(w->q); free(w->l); free(w->u); free(w->D); free(w->E); free(w->rv); free(w->Kb);
    free(w->x); free(w->y); free(w->z); free(w->xt); free(w->zt); free(w->rhs); free(w->t1); free(w->t2); free(w->t3);
    free(w->ncol); free(w->Ad); free(w->Bd); free(w->gd); free(w->perm); free(w->ct); free(w);
}

/* ---- centroidal_mpc.py: assemble the QP in OSQP form --------------------------------------- */
static void assemble(work* w, const port_opts* o, const double* x0, const double* x_ref) {
    const int N = w->N, n = w->n;
    csr* A = &w->A;
    int p = 0, r = 0;
    /* H = blkdiag(2Q xN, 2R xN)  (centroidal_mpc.py:183-201);  g (:248-253) */
    for (int k = 0; k < N; ++k)
        for (int i = 0; i < 12; ++i) {
            w->Pd[12 * k + i] = 2 * o->Q[i];
            w->q[12 * k + i] = -2 * o->Q[i] * x_ref[i * N + k];
            w->Pd[12 * N + 12 * k + i] = 2 * o->R[i];
            w->q[12 * N + 12 * k + i] = 0.0;
        }
    /* identity rows = variable bounds (centroidal_mpc.py:122-176) */
    for (int c = 0; c < n; ++c) { A->rp[r] = p; A->ci[p] = c; A->v[p] = 1.0; ++p; w->l[r] = -OSQP_INFTY; w->u[r] = OSQP_INFTY; ++r; }
    for (int k = 0; k < N; ++k)
        for (int leg = 0; leg < 4; ++leg) {
            const int j = 12 * N + 12 * k + 3 * leg;
            if (w->ct[leg * N + k] == 1) { w->l[j + 2] = o->fz_min; }
            else for (int c = 0; c < 3; ++c) { w->l[j + c] = 0.0; w->u[j + c] = 0.0; }
        }
    /* dynamics rows (centroidal_mpc.py:287-303): +I at x_{k+1}, -Ad at x_k (k>=1), -Bd[k] at u_k */
    for (int k = 0; k < N; ++k)
        for (int i = 0; i < 12; ++i) {
            A->rp[r] = p;
            if (k >= 1) for (int j = 0; j < 12; ++j) { A->ci[p] = 12 * (k - 1) + j; A->v[p] = -w->Ad[i * 12 + j]; ++p; }
            A->ci[p] = 12 * k + i; A->v[p] = 1.0; ++p;
            for (int j = 0; j < 12; ++j) { A->ci[p] = 12 * N + 12 * k + j; A->v[p] = -w->Bd[(size_t)k * 144 + i * 12 + j]; ++p; }
            double beq = w->gd[i];                                         /* :257-261 */
            if (k == 0) for (int j = 0; j < 12; ++j) beq += w->Ad[i * 12 + j] * x0[j];
            w->l[r] = beq; w->u[r] = beq; ++r;
        }
    /* friction rows (centroidal_mpc.py:324-359), bounds (:264-279) */
    for (int k = 0; k < N; ++k)
        for (int leg = 0; leg < 4; ++leg) {
            const int fx = 12 * N + 12 * k + 3 * leg, fy = fx + 1, fz = fx + 2;
            const int cols[4] = {fx, fx, fy, fy};
            const double sg[4] = {1.0, -1.0, 1.0, -1.0};
            for (int f = 0; f < 4; ++f) {
                A->rp[r] = p;
                A->ci[p] = cols[f]; A->v[p] = sg[f]; ++p;
                A->ci[p] = fz; A->v[p] = -o->mu; ++p;
                w->l[r] = -OSQP_INFTY;
                w->u[r] = (w->ct[leg * N + k] == 1) ? 0.0 : OSQP_INFTY;
                ++r;
            }
        }
    A->rp[r] = p;
}

/* OSQP scale_data (Ruiz equilibration + cost scaling), in place.  Returns c. */
static double ruiz(work* w, int iters) {
    const int n = w->n, m = w->m;
    csr* A = &w->A;
    double c = 1.0;
    for (int i = 0; i < n; ++i) w->D[i] = 1.0;
    for (int i = 0; i < m; ++i) w->E[i] = 1.0;
    double* Dt = w->t1;
    double* Et = w->t2;
    for (int it = 0; it < iters; ++it) {
        for (int j = 0; j < n; ++j) Dt[j] = fabs(w->Pd[j]);
        for (int r = 0; r < m; ++r) {
            double mx = 0;
            for (int p = A->rp[r]; p < A->rp[r + 1]; ++p) {
                const double a = fabs(A->v[p]);
                if (a > mx) mx = a;
                if (a > Dt[A->ci[p]]) Dt[A->ci[p]] = a;
            }
            Et[r] = mx;
        }
        for (int j = 0; j < n; ++j) Dt[j] = 1.0 / sqrt(limit_scaling(Dt[j]));
        for (int r = 0; r < m; ++r) Et[r] = 1.0 / sqrt(limit_scaling(Et[r]));
        for (int j = 0; j < n; ++j) { w->Pd[j] *= Dt[j] * Dt[j]; w->q[j] *= Dt[j]; w->D[j] *= Dt[j]; }
        for (int r = 0; r < m; ++r) {
            for (int p = A->rp[r]; p < A->rp[r + 1]; ++p) A->v[p] *= Et[r] * Dt[A->ci[p]];
            w->E[r] *= Et[r];
        }
        double mean = 0;
        for (int j = 0; j < n; ++j) mean += fabs(w->Pd[j]);
        mean /= n;
        double ct = limit_scaling(mean);
        double nq = vmax_abs(w->q, n);
        nq = nq < MIN_SCALING ? 1.0 : (nq > MAX_SCALING ? MAX_SCALING : nq);
        ct = 1.0 / (ct > nq ? ct : nq);
        for (int j = 0; j < n; ++j) { w->Pd[j] *= ct; w->q[j] *= ct; }
        c *= ct;
    }
    for (int r = 0; r < m; ++r) { w->l[r] *= w->E[r]; w->u[r] *= w->E[r]; }
    return c;
}

static void rho_vec(work* w, double rho) {
    for (int r = 0; r < w->m; ++r) {
        if (w->l[r] < -OSQP_INFTY * MIN_SCALING && w->u[r] > OSQP_INFTY * MIN_SCALING) w->rv[r] = RHO_MIN;
        else if (w->u[r] - w->l[r] < RHO_TOL) w->rv[r] = RHO_EQ_OVER_RHO_INEQ * rho;
        else w->rv[r] = rho;
    }
}

/* K = P + sigma I + A' diag(rv) A in band storage Kb[i*(BW+1)+d] = K[i][i-d], then Cholesky in place */
static int factor(work* w, double sigma) {
    const int n = w->n, m = w->m;
    const csr* A = &w->A;
    double* Kb = w->Kb;
    memset(Kb, 0, (size_t)n * (BW + 1) * sizeof(double));
    for (int c = 0; c < n; ++c) Kb[(size_t)w->perm[c] * (BW + 1)] = w->Pd[c] + sigma;
    for (int r = 0; r < m; ++r) {
        const double rr = w->rv[r];
        for (int p = A->rp[r]; p < A->rp[r + 1]; ++p) {
            const int ia = w->perm[A->ci[p]];
            const double va = rr * A->v[p];
            for (int p2 = A->rp[r]; p2 < A->rp[r + 1]; ++p2) {
                const int ib = w->perm[A->ci[p2]];
                if (ib <= ia) Kb[(size_t)ia * (BW + 1) + (ia - ib)] += va * A->v[p2];
            }
        }
    }
    for (int i = 0; i < n; ++i) {
        const int j0 = i - BW > 0 ? i - BW : 0;
        double* Li = Kb + (size_t)i * (BW + 1);
        for (int j = j0; j <= i; ++j) {
            const double* Lj = Kb + (size_t)j * (BW + 1);
            const int k0 = (j - BW > j0) ? j - BW : j0;
            double s = Li[i - j];
            for (int k = k0; k < j; ++k) s -= Li[i - k] * Lj[j - k];
            if (j < i) Li[i - j] = s / Lj[0];
            else { if (!(s > 0.0)) return 1; Li[0] = sqrt(s); }
        }
    }
    return 0;
}

/* solve K xt = rhs (both in the reference variable order) */
static void kkt_solve(work* w, const double* rhs, double* xt) {
    const int n = w->n;
    double* b = w->t3;
    for (int c = 0; c < n; ++c) b[w->perm[c]] = rhs[c];
    const double* Kb = w->Kb;
    for (int i = 0; i < n; ++i) {
        const double* Li = Kb + (size_t)i * (BW + 1);
        const int j0 = i - BW > 0 ? i - BW : 0;
        double s = b[i];
        for (int j = j0; j < i; ++j) s -= Li[i - j] * b[j];
        b[i] = s / Li[0];
    }
    for (int i = n - 1; i >= 0; --i) {
        const double bi = b[i] / Kb[(size_t)i * (BW + 1)];
        b[i] = bi;
        const int j0 = i - BW > 0 ? i - BW : 0;
        const double* Li = Kb + (size_t)i * (BW + 1);
        for (int j = j0; j < i; ++j) b[j] -= Li[i - j] * bi;
    }
    for (int c = 0; c < n; ++c) xt[c] = b[w->perm[c]];
}

/* One robot, one cycle.  w_io (24N) primal in/out, y_io (52N) duals in/out ([lam_x ; lam_a]),
 * rho_io adapted rho in/out (<= 0: start from opts.rho).  Returns iterations; *status 1 solved, -2 max_iter. */
static int solve_one(work* w, const port_opts* o, const double* x0, const double* x_ref, const double* r_foot,
                     const double* I_world, double mass, double t0, double dt, double hz, double duty,
                     const double* off, int warm, double* w_io, double* y_io, double* rho_io, int* status,
                     double* obj, int* nfac_out) {
    const int N = w->N, n = w->n, m = w->m;
    contact_table(t0, dt, N, hz, duty, off, w->ct);
    dynamics(N, x_ref, r_foot, I_world, mass, dt, w->Ad, w->Bd, w->gd);
    assemble(w, o, x0, x_ref);
    const double c = ruiz(w, o->scaling);
    double rho = (warm && *rho_io > 0.0) ? *rho_io : o->rho;
    rho_vec(w, rho);
    int nfac = 1;
    if (factor(w, o->sigma)) { *status = -7; return 0; }
    if (warm) {
        for (int i = 0; i < n; ++i) w->x[i] = w_io[i] / w->D[i];
        for (int r = 0; r < m; ++r) w->y[r] = c * y_io[r] / w->E[r];
    } else {
        memset(w->x, 0, n * sizeof(double));
        memset(w->y, 0, m * sizeof(double));
    }
    csr_mv(&w->A, w->x, w->z);
    int it = 0;
    *status = -2;
    for (it = 1; it <= o->max_iter; ++it) {
        for (int r = 0; r < m; ++r) w->t1[r] = w->rv[r] * w->z[r] - w->y[r];
        csr_mtv(&w->A, w->t1, w->rhs);
        for (int i = 0; i < n; ++i) w->rhs[i] += o->sigma * w->x[i] - w->q[i];
        kkt_solve(w, w->rhs, w->xt);
        csr_mv(&w->A, w->xt, w->zt);
        for (int i = 0; i < n; ++i) w->x[i] = o->alpha * w->xt[i] + (1 - o->alpha) * w->x[i];
        for (int r = 0; r < m; ++r) {
            const double zh = o->alpha * w->zt[r] + (1 - o->alpha) * w->z[r];
            double zn = zh + w->y[r] / w->rv[r];
            zn = zn < w->l[r] ? w->l[r] : (zn > w->u[r] ? w->u[r] : zn);
            w->y[r] += w->rv[r] * (zh - zn);
            w->z[r] = zn;
        }
        const int check = (it % o->check_termination) == 0;
        const int adapt = o->adaptive_rho_interval > 0 && (it % o->adaptive_rho_interval) == 0;
        if (!(check || adapt)) continue;
        csr_mv(&w->A, w->x, w->t1);                 /* Ax */
        csr_mtv(&w->A, w->y, w->t2);                /* A'y */
        double rp = 0, rd = 0, nAx = 0, nz = 0, nPx = 0, nAty = 0;
        for (int r = 0; r < m; ++r) {
            const double a = fabs(w->t1[r] - w->z[r]); if (a > rp) rp = a;
            if (fabs(w->t1[r]) > nAx) nAx = fabs(w->t1[r]);
            if (fabs(w->z[r]) > nz) nz = fabs(w->z[r]);
        }
        for (int i = 0; i < n; ++i) {
            const double px = w->Pd[i] * w->x[i];
            const double a = fabs(px + w->q[i] + w->t2[i]); if (a > rd) rd = a;
            if (fabs(px) > nPx) nPx = fabs(px);
            if (fabs(w->t2[i]) > nAty) nAty = fabs(w->t2[i]);
        }
        const double nq = vmax_abs(w->q, n);
        const double np_ = nAx > nz ? nAx : nz;
        double nd_ = nPx > nAty ? nPx : nAty; if (nq > nd_) nd_ = nq;
        if (check && rp <= o->eps_abs + o->eps_rel * np_ && rd <= o->eps_abs + o->eps_rel * nd_) { *status = 1; break; }
        if (adapt) {
            double rn = rho * sqrt((rp / (np_ + 1e-10)) / (rd / (nd_ + 1e-10) + 1e-10));
            rn = rn < RHO_MIN ? RHO_MIN : (rn > RHO_MAX ? RHO_MAX : rn);
            if (rn > rho * o->adaptive_rho_tolerance || rn < rho / o->adaptive_rho_tolerance) {
                rho = rn;
                rho_vec(w, rho);
                ++nfac;
                if (factor(w, o->sigma)) { *status = -7; return it; }
            }
        }
    }
    if (it > o->max_iter) it = o->max_iter;
    double J = 0;
    for (int i = 0; i < n; ++i) {
        const double xs = w->D[i] * w->x[i];
        w_io[i] = xs;
    }
    for (int r = 0; r < m; ++r) y_io[r] = w->E[r] * w->y[r] / c;
    /* objective in unscaled units: 1/2 w'Hw + g'w */
    for (int k = 0; k < N; ++k)
        for (int i = 0; i < 12; ++i) {
            const double xv = w_io[12 * k + i], uv = w_io[12 * N + 12 * k + i];
            J += o->Q[i] * xv * xv - 2 * o->Q[i] * x_ref[i * N + k] * xv + o->R[i] * uv * uv;
        }
    *obj = J;
    *rho_io = rho;
    *nfac_out = nfac;
    return it;
}

/* Batched entry: B independent robots, `nthreads` POSIX threads (each with its own workspace,
 * robots handed out one at a time from an atomic counter).
 * Layouts as include/cmpc.h: x0 (B,12), x_ref (B,12,N), r_foot (B,4,3,N), I_world (B,3,3), mass (B), t0 (B).
 * w_io (B,24N), y_io (B,52N), rho_io (B), iters (B), status (B), obj (B), nfac (B).                */
typedef struct {
    int B, N, warm;
    const double *x0, *x_ref, *r_foot, *I_world, *mass, *t0, *off;
    double dt, hz, duty;
    const port_opts* opts;
    double *w_io, *y_io, *rho_io, *obj;
    int32_t *iters, *status, *nfac;
    atomic_int next;
} job;

static void* worker(void* arg) {
    job* j = (job*)arg;
    const int N = j->N;
    work* w = work_new(N);
    for (;;) {
        const int b = atomic_fetch_add(&j->next, 1);
        if (b >= j->B) break;
        int st = 0, nf = 0;
        double J = 0;
        j->iters[b] = solve_one(w, j->opts, j->x0 + (size_t)b * 12, j->x_ref + (size_t)b * 12 * N,
                                j->r_foot + (size_t)b * 12 * N, j->I_world + (size_t)b * 9, j->mass[b], j->t0[b],
                                j->dt, j->hz, j->duty, j->off, j->warm, j->w_io + (size_t)b * 24 * N,
                                j->y_io + (size_t)b * 52 * N, j->rho_io + b, &st, &J, &nf);
        j->status[b] = st; j->obj[b] = J; j->nfac[b] = nf;
    }
    work_free(w);
    return NULL;
}

int port_solve_batch(int B, int N, const double* x0, const double* x_ref, const double* r_foot,
                     const double* I_world, const double* mass, const double* t0, double dt, double hz,
                     double duty, const double* phase_offset, const port_opts* opts, int warm, int nthreads,
                     double* w_io, double* y_io, double* rho_io, int32_t* iters, int32_t* status, double* obj,
                     int32_t* nfac) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    job j = {B, N, warm, x0, x_ref, r_foot, I_world, mass, t0, phase_offset, dt, hz, duty, opts,
             w_io, y_io, rho_io, obj, iters, status, nfac, 0};
    if (nthreads == 1) { worker(&j); return 0; }
    pthread_t th[256];
    int started = 0;
    for (int i = 0; i < nthreads; ++i) if (pthread_create(&th[started], NULL, worker, &j) == 0) ++started;
    if (started == 0) worker(&j);
    for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
    return 0;
}

int port_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
