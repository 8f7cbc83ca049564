// cmpc_core.cuh -- per-robot convex-MPC QP: dynamics, condensing, factorisation, solve.
//
// One cooperative thread array (CTA) owns one robot's QP; every routine below is written as
// "CTA-strided loops separated by barriers" over a workspace that lives in shared memory.
// The same source also compiles for the host with a one-thread CTA (tests/_emul, used by the CPU
// test-suite to check the kernel logic against the oracle; it is NOT part of the product library).
//
// What it replaces in the reference (ltinphan/convex-mpc-unitree-go2, convex_mpc/):
//   com_trajectory.py:221-286   continuous + ZOH-discrete centroidal dynamics   -> dyn_*()
//   centroidal_mpc.py:235-303   per-cycle QP update (g, A, bounds)              -> recursions(), build_H()
//   centroidal_mpc.py:122-176   variable bounds (swing = 0, stance fz >= fz_min) -> feet elimination + rows
//   centroidal_mpc.py:324-359   friction pyramid rows                            -> row_def()
//   centroidal_mpc.py:98        CasADi conic -> OSQP solve                       -> solve_active_set(), admm()
//
// Formulation.  The reference keeps states as decision variables and the dynamics as equality
// rows (24N variables).  Here the states are eliminated (condensed QP in the 12N forces) and the
// swing-leg forces, which the reference pins to zero with lbx = ubx = 0, are eliminated too, so
// the working problem has n = 3 * (number of stance foot-steps) variables and 5 inequality rows
// per stance foot-step: fz >= fz_min and the four pyramid faces.  Both eliminations are exact.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define CMPC_HD __host__ __device__ __forceinline__
#define CMPC_HDN __host__ __device__ __noinline__
#else
#define CMPC_HD inline
#define CMPC_HDN inline
#endif

namespace cmpc {

// ----------------------------------------------------------------------------------------------
// CTA abstraction
// ----------------------------------------------------------------------------------------------
struct Cta {
    int tid;   // thread index inside the group
    int nt;    // group size
    int warp;  // 1: the group is a single warp (barrier = __syncwarp)
};

CMPC_HD void cta_sync(const Cta& c) {
#if defined(__CUDA_ARCH__)
    if (c.warp) __syncwarp(); else __syncthreads();
#else
    (void)c;
#endif
}

#define CTA_FOR(i, lo, hi) for (int i = (lo) + c.tid; i < (hi); i += c.nt)

CMPC_HD int tri(int i) { return (i * (i + 1)) >> 1; }

struct Params {
    double Q[12];
    double R[12];
    double mu, fz_min;
    double eps_abs, eps_rel;
    double rho0, sigma, alpha;
    int max_iter, mode, polish, check_termination, adaptive_rho_interval;
    int pdas_max_iter;
};

// status / path codes (mirrors include/cmpc.h)
enum { ST_SOLVED = 1, ST_INACCURATE = 2, ST_MAX_ITER = -2, ST_NON_CVX = -7, ST_TOO_MANY_FEET = -20 };
enum { PATH_UNCONSTRAINED = 0, PATH_ACTIVE_SET = 1, PATH_ADMM = 2, PATH_ADMM_POLISH = 3, PATH_RICCATI = 4 };
enum { NSTAT = 8 };

// Shared-memory workspace of one CTA (all pointers into one dynamic smem allocation).
struct Ws {
    double* Hp;     // packed lower triangle, (nmax+1)(nmax+2)/2 doubles: H -> L -> W=L^-1 -> M=H^-1
    double* P;      // N*144 cost-to-go matrices;  aliased by S (active-set Schur complement)
    double* Bcol;   // nmax*12: column of B_d[k] for every free variable
    double* Ad;     // 144
    double* T144;   // 144 scratch
    double* gd;     // 12
    double* x0;     // 12
    double* E;      // 12N free-response error  e_i = x^free_{i+1} - xref_i
    double* q;      // 12N backward gradient recursion, later co-states
    double* X;      // 12N rolled-out states
    double* g;      // nmax
    double* u0;     // nmax unconstrained minimiser
    double* x;      // nmax current iterate
    double* t1;     // nmax+1
    double* t2;     // nmax+1
    double* hx;     // nmax
    double* lam;    // 5*nfmax multipliers (>=0), row form a'x <= b
    double* viol;   // 5*nfmax
    double* z;      // 5*nfmax (ADMM)
    double* yv;     // 5*nfmax (ADMM duals, OSQP sign)
    double* red;    // 40 reduction scratch
    double* sc;     // 16 scalar doubles
    int* fk;        // nfmax foot -> step
    int* fl;        // nfmax foot -> leg
    int* vstart;    // N+1 first variable of each step
    int* aidx;      // kcap active row ids
    int* isc;       // 16 scalar ints
    unsigned char* act;       // 5*nfmax
    unsigned char* act_prev;  // 5*nfmax
    unsigned char* act_prev2; // 5*nfmax
};

// capacity of the active-set Schur complement: packed kcap x kcap triangle aliased onto P (N*144)
CMPC_HD int kcap_for(int N) {
    int k = 1;
    while ((k + 1) * (k + 2) / 2 <= N * 144) ++k;   // packed lower triangle
    return k;
}

// Carve the workspace out of `base` (16-byte aligned).  Returns bytes used.
// `hp_ext` != null: the packed matrix lives there (global memory) instead of in the carve.
CMPC_HD size_t ws_carve(Ws& w, unsigned char* base, int N, int nfmax, double* hp_ext) {
    const int nmax = 3 * nfmax;
    double* p = reinterpret_cast<double*>(base);
    auto take = [&](size_t n) { double* r = p; p += (n + 1) & ~(size_t)1; return r; };
    w.Hp = hp_ext ? hp_ext : take((size_t)(nmax + 1) * (nmax + 2) / 2);
    w.P = take((size_t)N * 144);
    w.Bcol = take((size_t)nmax * 12);
    w.Ad = take(144);
    w.T144 = take(144);
    w.gd = take(12);
    w.x0 = take(12);
    w.E = take((size_t)12 * N);
    w.q = take((size_t)12 * N);
    w.X = take((size_t)12 * N);
    w.g = take(nmax);
    w.u0 = take(nmax);
    w.x = take(nmax);
    w.t1 = take(nmax + 2);
    w.t2 = take(nmax + 2);
    w.hx = take(nmax);
    w.lam = take((size_t)5 * nfmax);
    w.viol = take((size_t)5 * nfmax);
    w.z = take((size_t)5 * nfmax);
    w.yv = take((size_t)5 * nfmax);
    w.red = take(40);
    w.sc = take(16);
    int* ip = reinterpret_cast<int*>(p);
    auto itake = [&](size_t n) { int* r = ip; ip += (n + 3) & ~(size_t)3; return r; };
    w.fk = itake(nfmax);
    w.fl = itake(nfmax);
    w.vstart = itake(N + 1);
    w.aidx = itake(kcap_for(N));
    w.isc = itake(16);
    unsigned char* cp = reinterpret_cast<unsigned char*>(ip);
    auto ctake = [&](size_t n) { unsigned char* r = cp; cp += (n + 15) & ~(size_t)15; return r; };
    w.act = ctake((size_t)5 * nfmax);
    w.act_prev = ctake((size_t)5 * nfmax);
    w.act_prev2 = ctake((size_t)5 * nfmax);
    return (size_t)(cp - base);
}

// ----------------------------------------------------------------------------------------------
// CTA-wide reductions (device: warp shuffles + one smem hop; host: identity)
// ----------------------------------------------------------------------------------------------
CMPC_HD double cta_max(const Cta& c, double v, double* red) {
#if defined(__CUDA_ARCH__)
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (c.warp) return v;
    const int wid = c.tid >> 5, nw = (c.nt + 31) >> 5;
    __syncthreads();
    if ((c.tid & 31) == 0) red[wid] = v;
    __syncthreads();
    double r = red[0];
    for (int i = 1; i < nw; ++i) r = fmax(r, red[i]);
    return r;
#else
    (void)c; (void)red;
    return v;
#endif
}

CMPC_HD double cta_sum(const Cta& c, double v, double* red) {
#if defined(__CUDA_ARCH__)
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (c.warp) return v;
    const int wid = c.tid >> 5, nw = (c.nt + 31) >> 5;
    __syncthreads();
    if ((c.tid & 31) == 0) red[wid] = v;
    __syncthreads();
    double r = red[0];
    for (int i = 1; i < nw; ++i) r += red[i];
    return r;
#else
    (void)c; (void)red;
    return v;
#endif
}

// ----------------------------------------------------------------------------------------------
// Contact schedule (gait.py:26-37) -- bit-exact: every operation is a single correctly rounded
// IEEE double op in the reference's order; no FMA contraction is allowed here.
// ----------------------------------------------------------------------------------------------
CMPC_HD double dmul_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    volatile double r = a * b; return r;
#endif
}
CMPC_HD double dadd_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double r = a + b; return r;
#endif
}
CMPC_HD double ddiv_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(a, b);
#else
    volatile double r = a / b; return r;
#endif
}

// period = 1 / frequency is formed by the caller exactly as gait.py:17 does.
CMPC_HD int stance_bit(double t0, double dt, int k, double period, double offset, double duty) {
    double t = dadd_rn(t0, dmul_rn((double)k, dt));   // gait.py:29   t0 + arange(N)*dt
    t = dadd_rn(t, ddiv_rn(dt, 2.0));                 // gait.py:30   + dt/2
    double ph = dadd_rn(offset, ddiv_rn(t, period));  // gait.py:33   offset + t/T
    double r = fmod(ph, 1.0);                         // np.mod(.,1.0): fmod, then sign fix-up
    if (r != 0.0 && r < 0.0) r = dadd_rn(r, 1.0);
    return r < duty ? 1 : 0;                          // gait.py:36
}

// ----------------------------------------------------------------------------------------------
// Dynamics (com_trajectory.py:221-286), closed-form ZOH because A_c^2 = 0.
// ----------------------------------------------------------------------------------------------
struct DynCommon {
    double cy, sy;       // cos/sin of the horizon-average yaw (com_trajectory.py:226-232)
    double Iinv[9];      // inverse world inertia (com_trajectory.py:255)
    double minv;         // 1/m
    double dt;
};

CMPC_HD void dyn_common(DynCommon& d, const double* x_ref, int N, const double* I_world, double mass, double dt) {
    double s = 0.0;
    for (int i = 0; i < N; ++i) s += x_ref[5 * N + i];
    const double yaw = s / (double)N;
    d.cy = cos(yaw);
    d.sy = sin(yaw);
    const double a = I_world[0], b = I_world[1], cc = I_world[2];
    const double dd = I_world[3], e = I_world[4], f = I_world[5];
    const double g = I_world[6], h = I_world[7], k = I_world[8];
    const double A = e * k - f * h, Bc = -(dd * k - f * g), C = dd * h - e * g;
    const double det = a * A + b * Bc + cc * C;
    const double id = 1.0 / det;
    d.Iinv[0] = A * id;
    d.Iinv[1] = -(b * k - cc * h) * id;
    d.Iinv[2] = (b * f - cc * e) * id;
    d.Iinv[3] = Bc * id;
    d.Iinv[4] = (a * k - cc * g) * id;
    d.Iinv[5] = -(a * f - cc * dd) * id;
    d.Iinv[6] = C * id;
    d.Iinv[7] = -(a * h - b * g) * id;
    d.Iinv[8] = (a * e - b * dd) * id;
    d.minv = 1.0 / mass;
    d.dt = dt;
}

// A_d = I + dt A_c  (row-major 12x12)
CMPC_HD double dyn_Ad(const DynCommon& d, int r, int c) {
    if (r == c) return 1.0;
    if (r < 3 && c == r + 6) return d.dt;
    if (r >= 3 && r < 6 && c >= 9) {
        // rows 3..5 = dt * Rz^T, Rz^T = [[cy, sy, 0], [-sy, cy, 0], [0, 0, 1]]
        const int rr = r - 3, cc = c - 9;
        const double RzT[9] = {d.cy, d.sy, 0.0, -d.sy, d.cy, 0.0, 0.0, 0.0, 1.0};
        return d.dt * RzT[rr * 3 + cc];
    }
    return 0.0;
}

CMPC_HD void dyn_gd(const DynCommon& d, double* gd) {
    for (int i = 0; i < 12; ++i) gd[i] = 0.0;
    gd[2] = -9.81 * (d.dt * d.dt / 2.0);
    gd[8] = -9.81 * d.dt;
}

// Column `comp` (0..2) of the 12x3 block of B_d[k] that belongs to one leg with lever arm r.
CMPC_HD void dyn_Bd_col(const DynCommon& d, const double r[3], int comp, double out[12]) {
    double sk[3];
    if (comp == 0) { sk[0] = 0.0;  sk[1] = r[2];  sk[2] = -r[1]; }
    else if (comp == 1) { sk[0] = -r[2]; sk[1] = 0.0;  sk[2] = r[0]; }
    else { sk[0] = r[1];  sk[1] = -r[0]; sk[2] = 0.0; }
    double w[3];
    for (int i = 0; i < 3; ++i) w[i] = d.Iinv[3 * i] * sk[0] + d.Iinv[3 * i + 1] * sk[1] + d.Iinv[3 * i + 2] * sk[2];
    const double h = d.dt * d.dt / 2.0;
    for (int i = 0; i < 12; ++i) out[i] = 0.0;
    out[comp] = h * d.minv;
    out[3] = h * (d.cy * w[0] + d.sy * w[1]);
    out[4] = h * (-d.sy * w[0] + d.cy * w[1]);
    out[5] = h * w[2];
    out[6 + comp] = d.dt * d.minv;
    out[9] = d.dt * w[0];
    out[10] = d.dt * w[1];
    out[11] = d.dt * w[2];
}

// ----------------------------------------------------------------------------------------------
// Per-robot input view
// ----------------------------------------------------------------------------------------------
struct QpIn {
    const double* Ad;       // 144 or null
    const double* Bd;       // N*144 or null
    const double* gd;       // 12 or null
    const double* x0;       // 12
    const double* x_ref;    // 12*N (row-major (12,N))
    const double* r_foot;   // 4*3*N or null
    const double* I_world;  // 9 or null
    double mass, dt;
    const uint64_t* mask;   // W words (null = all stance)
    int N;
};

CMPC_HD int mask_bit(const uint64_t* mask, int N, int leg, int k) {
    if (!mask) return 1;
    const int b = leg * N + k;
    return (int)((mask[b >> 6] >> (b & 63)) & 1ull);
}

// Stance foot-steps in (step, leg) order -> variables 3j..3j+2.  Returns the true count (may exceed nfmax).
CMPC_HD int setup_feet(const Cta& c, const QpIn& in, Ws& w, int nfmax) {
    if (c.tid == 0) {
        int nf = 0;
        for (int k = 0; k < in.N; ++k) {
            w.vstart[k] = 3 * (nf < nfmax ? nf : nfmax);
            for (int leg = 0; leg < 4; ++leg)
                if (mask_bit(in.mask, in.N, leg, k)) {
                    if (nf < nfmax) { w.fk[nf] = k; w.fl[nf] = leg; }
                    ++nf;
                }
        }
        w.vstart[in.N] = 3 * (nf < nfmax ? nf : nfmax);
        w.isc[0] = nf;
    }
    cta_sync(c);
    return w.isc[0];
}

// Load / compute A_d, g_d, x0 and the B_d column of every free variable.
CMPC_HD void load_dynamics(const Cta& c, const QpIn& in, Ws& w, int n) {
    const int N = in.N;
    if (in.Ad) {
        CTA_FOR(i, 0, 144) w.Ad[i] = in.Ad[i];
        CTA_FOR(i, 0, 12) { w.gd[i] = in.gd[i]; w.x0[i] = in.x0[i]; }
        CTA_FOR(v, 0, n) {
            const int j = v / 3, comp = v - 3 * j;
            const double* B = in.Bd + (size_t)w.fk[j] * 144 + 3 * w.fl[j] + comp;
            for (int r = 0; r < 12; ++r) w.Bcol[v * 12 + r] = B[r * 12];
        }
    } else {
        DynCommon d;
        dyn_common(d, in.x_ref, N, in.I_world, in.mass, in.dt);
        CTA_FOR(i, 0, 144) w.Ad[i] = dyn_Ad(d, i / 12, i % 12);
        if (c.tid == 0) dyn_gd(d, w.gd);
        CTA_FOR(i, 0, 12) w.x0[i] = in.x0[i];
        CTA_FOR(v, 0, n) {
            const int j = v / 3, comp = v - 3 * j;
            const int k = w.fk[j], leg = w.fl[j];
            double r[3];
            for (int a = 0; a < 3; ++a) r[a] = in.r_foot[(size_t)(leg * 3 + a) * N + k];
            double col[12];
            dyn_Bd_col(d, r, comp, col);
            for (int a = 0; a < 12; ++a) w.Bcol[v * 12 + a] = col[a];
        }
    }
    cta_sync(c);
}

// B_d column of an arbitrary (step, leg, comp) -- used for the multipliers of eliminated variables.
CMPC_HD void any_Bd_col(const QpIn& in, int k, int leg, int comp, double out[12]) {
    if (in.Ad) {
        const double* B = in.Bd + (size_t)k * 144 + 3 * leg + comp;
        for (int r = 0; r < 12; ++r) out[r] = B[r * 12];
    } else {
        DynCommon d;
        dyn_common(d, in.x_ref, in.N, in.I_world, in.mass, in.dt);
        double r[3];
        for (int a = 0; a < 3; ++a) r[a] = in.r_foot[(size_t)(leg * 3 + a) * in.N + k];
        dyn_Bd_col(d, r, comp, out);
    }
}

// ----------------------------------------------------------------------------------------------
// Recursions that replace the explicit A_qp power chain / B_qp block-Toeplitz products:
//   e_i   = A_d^{i+1} x0 + G_i - xref_i                       (forward free response)
//   P_a   = Q + A_d' P_{a+1} A_d,  P_{N-1} = Q               (=> H[a,b] = 2 Bd[a]' P_a A_d^{a-b} Bd[b])
//   q_a   = Q e_a + A_d' q_{a+1}                              (=> g_a   = 2 Bd[a]' q_a)
// ----------------------------------------------------------------------------------------------
CMPC_HD void free_response(const Cta& c, const QpIn& in, Ws& w) {
    const int N = in.N;
    double* cur = w.t1;
    double* nxt = w.t2;
    CTA_FOR(r, 0, 12) cur[r] = w.x0[r];
    cta_sync(c);
    for (int i = 0; i < N; ++i) {
        CTA_FOR(r, 0, 12) {
            double s = w.gd[r];
            for (int k = 0; k < 12; ++k) s += w.Ad[r * 12 + k] * cur[k];
            nxt[r] = s;
            w.E[i * 12 + r] = s - in.x_ref[r * N + i];
        }
        cta_sync(c);
        double* t = cur; cur = nxt; nxt = t;
    }
}

CMPC_HD void cost_to_go(const Cta& c, const Params& p, int N, Ws& w, bool with_q) {
    double* PN = w.P + (size_t)(N - 1) * 144;
    CTA_FOR(i, 0, 144) PN[i] = (i / 12 == i % 12) ? p.Q[i / 12] : 0.0;
    if (with_q) { CTA_FOR(r, 0, 12) w.q[(N - 1) * 12 + r] = p.Q[r] * w.E[(N - 1) * 12 + r]; }
    cta_sync(c);
    for (int a = N - 2; a >= 0; --a) {
        const double* Pn = w.P + (size_t)(a + 1) * 144;
        CTA_FOR(i, 0, 144) {
            const int r = i / 12, cc = i % 12;
            double s = 0.0;
            for (int k = 0; k < 12; ++k) s += Pn[r * 12 + k] * w.Ad[k * 12 + cc];
            w.T144[i] = s;
        }
        if (with_q) {
            // q_a = Q e_a + A_d' q_{a+1}; uses threads at the top of the CTA when there are enough
            const int base = (c.nt >= 160) ? 144 : 0;
            for (int r = c.tid - base; r >= 0 && r < 12; r += c.nt) {
                double s = p.Q[r] * w.E[a * 12 + r];
                for (int k = 0; k < 12; ++k) s += w.Ad[k * 12 + r] * w.q[(a + 1) * 12 + k];
                w.q[a * 12 + r] = s;
            }
        }
        cta_sync(c);
        double* Pa = w.P + (size_t)a * 144;
        CTA_FOR(i, 0, 144) {
            const int r = i / 12, cc = i % 12;
            double s = (r == cc) ? p.Q[r] : 0.0;
            for (int k = 0; k < 12; ++k) s += w.Ad[k * 12 + r] * w.T144[k * 12 + cc];
            Pa[i] = s;
        }
        cta_sync(c);
    }
}

CMPC_HD void gradient_vec(const Cta& c, Ws& w, int n) {
    CTA_FOR(v, 0, n) {
        const int k = w.fk[v / 3];
        double s = 0.0;
        for (int r = 0; r < 12; ++r) s += w.Bcol[v * 12 + r] * w.q[k * 12 + r];
        w.g[v] = 2.0 * s;
    }
    cta_sync(c);
}

// Packed lower triangle of  H + diag(shift),  shift_v = sigma + rho * d_c  (d = A'A diagonal).
CMPC_HD void build_H(const Cta& c, const Params& p, int N, Ws& w, int n, double sigma, double rho) {
    const double dz = 1.0 + 4.0 * p.mu * p.mu;
    CTA_FOR(v, 0, n) {
        const int j = v / 3, comp = v - 3 * j;
        const int k0 = w.fk[j];
        double e[12], t[12];
        for (int r = 0; r < 12; ++r) e[r] = w.Bcol[v * 12 + r];
        for (int a = k0; a < N; ++a) {
            const double* Pa = w.P + (size_t)a * 144;
            for (int r = 0; r < 12; ++r) {
                double s = 0.0;
                for (int k = 0; k < 12; ++k) s += Pa[r * 12 + k] * e[k];
                t[r] = s;
            }
            const int lo = (a == k0) ? v : w.vstart[a];
            const int hi = w.vstart[a + 1];
            for (int wv = lo; wv < hi; ++wv) {
                double s = 0.0;
                for (int r = 0; r < 12; ++r) s += w.Bcol[wv * 12 + r] * t[r];
                s *= 2.0;
                if (wv == v) s += 2.0 * p.R[3 * w.fl[j] + comp] + sigma + rho * (comp == 2 ? dz : 2.0);
                w.Hp[tri(wv) + v] = s;
            }
            if (a + 1 < N) {
                for (int r = 0; r < 12; ++r) {
                    double s = 0.0;
                    for (int k = 0; k < 12; ++k) s += w.Ad[r * 12 + k] * e[k];
                    t[r] = s;
                }
                for (int r = 0; r < 12; ++r) e[r] = t[r];
            }
        }
    }
    cta_sync(c);
}

// ----------------------------------------------------------------------------------------------
// Dense kernels on the packed lower triangle
// ----------------------------------------------------------------------------------------------
// Cholesky H = L L' in place over rows 0..nrows-1 (nrows = n, or n+1 when row n carries g' so that
// the factorisation leaves L^-1 g there).  Returns 0 on success, 1 if a pivot is not positive.
CMPC_HD int chol_packed(const Cta& c, double* Hp, int n, int nrows, int* flag) {
    if (c.tid == 0) *flag = 0;
    cta_sync(c);
    for (int j = 0; j < n; ++j) {
        const double* rj = Hp + tri(j);
        CTA_FOR(i, j, nrows) {
            double* ri = Hp + tri(i);
            double s = ri[j];
            for (int k = 0; k < j; ++k) s -= ri[k] * rj[k];
            if (i == j) {
                if (!(s > 0.0)) { *flag = 1; s = 1.0; }
                s = sqrt(s);
            }
            ri[j] = s;
        }
        cta_sync(c);
        const double dinv = 1.0 / rj[j];
        CTA_FOR(i, j + 1, nrows) Hp[tri(i) + j] *= dinv;
        cta_sync(c);
    }
    return *flag;
}

// Solve L' x = rhs (in place in x), executed by one warp (or the host's single thread).
CMPC_HD void backsolve_warp(const Cta& c, const double* Hp, int n, double* x) {
    for (int j = n - 1; j >= 0; --j) {
        const double* rj = Hp + tri(j);
        const double xj = x[j] / rj[j];
        cta_sync(c);
        if (c.tid == 0) x[j] = xj;
        CTA_FOR(i, 0, j) x[i] -= rj[i] * xj;
        cta_sync(c);
    }
}

// W = L^-1 in place (row by row from the top; row i of L is staged in tmp).
CMPC_HD void trtri_packed(const Cta& c, double* Hp, int n, double* tmp) {
    for (int i = 0; i < n; ++i) {
        double* ri = Hp + tri(i);
        CTA_FOR(k, 0, i + 1) tmp[k] = ri[k];
        cta_sync(c);
        const double wii = 1.0 / tmp[i];
        CTA_FOR(j, 0, i + 1) {
            if (j == i) { ri[i] = wii; }
            else {
                double s = 0.0;
                for (int k = j; k < i; ++k) s += tmp[k] * Hp[tri(k) + j];
                ri[j] = -wii * s;
            }
        }
        cta_sync(c);
    }
}

// M = W' W in place (lower triangle), rows ascending.
CMPC_HD void lauum_packed(const Cta& c, double* Hp, int n, double* tmp) {
    for (int i = 0; i < n; ++i) {
        CTA_FOR(j, 0, i + 1) {
            double s = 0.0;
            for (int k = i; k < n; ++k) s += Hp[tri(k) + i] * Hp[tri(k) + j];
            tmp[j] = s;
        }
        cta_sync(c);
        CTA_FOR(j, 0, i + 1) Hp[tri(i) + j] = tmp[j];
        cta_sync(c);
    }
}

// y = M v for a packed symmetric M.
CMPC_HD void symv_packed(const Cta& c, const double* Hp, int n, const double* v, double* y) {
    CTA_FOR(i, 0, n) {
        const double* ri = Hp + tri(i);
        double s = 0.0;
        for (int j = 0; j <= i; ++j) s += ri[j] * v[j];
        for (int j = i + 1; j < n; ++j) s += Hp[tri(j) + i] * v[j];
        y[i] = s;
    }
    cta_sync(c);
}

CMPC_HD double sym_at(const double* Hp, int i, int j) { return i >= j ? Hp[tri(i) + j] : Hp[tri(j) + i]; }

// ----------------------------------------------------------------------------------------------
// Inequality rows, 5 per stance foot-step f (variables 3f, 3f+1, 3f+2 = fx, fy, fz), as a'x <= b:
//   t=0:  -fz <= -fz_min          (centroidal_mpc.py:163-170)
//   t=1:  +fx - mu fz <= 0        (centroidal_mpc.py:337-341)
//   t=2:  -fx - mu fz <= 0        (:342-346)
//   t=3:  +fy - mu fz <= 0        (:347-351)
//   t=4:  -fy - mu fz <= 0        (:352-356)
// ----------------------------------------------------------------------------------------------
struct RowDef { int c1, c2; double s1, s2, b; };

CMPC_HD RowDef row_def(int row, double mu, double fz_min) {
    const int f = row / 5, t = row - 5 * f;
    RowDef r;
    r.c2 = 3 * f + 2;
    if (t == 0) { r.c1 = 3 * f + 2; r.s1 = -1.0; r.s2 = 0.0; r.b = -fz_min; }
    else {
        r.c1 = 3 * f + ((t - 1) >> 1);
        r.s1 = ((t - 1) & 1) ? -1.0 : 1.0;
        r.s2 = -mu;
        r.b = 0.0;
    }
    return r;
}

CMPC_HD void foot_viol(const double* x, int f, double mu, double fz_min, double v[5]) {
    const double fx = x[3 * f], fy = x[3 * f + 1], fz = x[3 * f + 2];
    v[0] = fz_min - fz;
    v[1] = fx - mu * fz;
    v[2] = -fx - mu * fz;
    v[3] = fy - mu * fz;
    v[4] = -fy - mu * fz;
}

// viol for all rows; returns the largest violation.
CMPC_HD double all_viol(const Cta& c, const Params& p, Ws& w, const double* x, int nf) {
    double m = -1e300;
    CTA_FOR(f, 0, nf) {
        double v[5];
        foot_viol(x, f, p.mu, p.fz_min, v);
        for (int t = 0; t < 5; ++t) { w.viol[5 * f + t] = v[t]; m = fmax(m, v[t]); }
    }
    return cta_max(c, m, w.red);
}

// v = A' lam  (lam >= 0 in the a'x <= b row form)
CMPC_HD void At_lam(const Cta& c, const Params& p, const double* lam, double* v, int nf) {
    CTA_FOR(f, 0, nf) {
        const double* l = lam + 5 * f;
        v[3 * f] = l[1] - l[2];
        v[3 * f + 1] = l[3] - l[4];
        v[3 * f + 2] = -l[0] - p.mu * (l[1] + l[2] + l[3] + l[4]);
    }
}

// Small dense Cholesky solve S lam = rhs (S packed lower triangle, k x k), in place.
// Returns 1 on failure (S not positive definite).
CMPC_HD int small_chol_solve(const Cta& c, double* S, int k, double* rhs, int* flag) {
    if (c.tid == 0) *flag = 0;
    cta_sync(c);
    for (int j = 0; j < k; ++j) {
        if (c.tid == 0) {
            double d = S[tri(j) + j];
            if (!(d > 0.0)) { *flag = 1; d = 1.0; }
            S[tri(j) + j] = sqrt(d);
        }
        cta_sync(c);
        const double dinv = 1.0 / S[tri(j) + j];
        CTA_FOR(i, j + 1, k) S[tri(i) + j] *= dinv;
        cta_sync(c);
        // trailing update, lower part only
        const int m = k - j - 1;
        CTA_FOR(e, 0, m * m) {
            const int a = j + 1 + e / m, b = j + 1 + e % m;
            if (b <= a) S[tri(a) + b] -= S[tri(a) + j] * S[tri(b) + j];
        }
        cta_sync(c);
    }
    if (*flag) return 1;
    // forward then backward substitution by thread 0 of the group (k is small)
    if (c.tid == 0) {
        for (int i = 0; i < k; ++i) {
            double s = rhs[i];
            for (int j = 0; j < i; ++j) s -= S[tri(i) + j] * rhs[j];
            rhs[i] = s / S[tri(i) + i];
        }
        for (int i = k - 1; i >= 0; --i) {
            double s = rhs[i];
            for (int j = i + 1; j < k; ++j) s -= S[tri(j) + i] * rhs[j];
            rhs[i] = s / S[tri(i) + i];
        }
    }
    cta_sync(c);
    return 0;
}

// Equality-constrained solve on the working set w.act:  lam = S^-1 (A_act u0 - b_act),
// x = u0 - M A_act' lam  with  S = A_act M A_act'  (M = H^-1 explicit, packed in w.Hp).
// Returns 0 on success, 1 if the set is too large for the scratch or S is not positive definite.
CMPC_HD int working_set_solve(const Cta& c, const Params& p, Ws& w, int n, int nf, int kcap) {
    const int m = 5 * nf;
    double* S = w.P;  // alias: the cost-to-go matrices are dead once H has been built
    if (c.tid == 0) {
        int k = 0;
        for (int r = 0; r < m; ++r)
            if (w.act[r]) { if (k < kcap) w.aidx[k] = r; ++k; }
        w.isc[1] = k;
    }
    cta_sync(c);
    const int k = w.isc[1];
    if (k > kcap) return 1;
    if (k == 0) {
        CTA_FOR(i, 0, n) w.x[i] = w.u0[i];
        CTA_FOR(r, 0, m) w.lam[r] = 0.0;
        cta_sync(c);
        return 0;
    }
    CTA_FOR(e, 0, k * k) {
        const int a = e / k, b = e - a * k;
        if (b <= a) {
            const RowDef ra = row_def(w.aidx[a], p.mu, p.fz_min);
            const RowDef rb = row_def(w.aidx[b], p.mu, p.fz_min);
            double s = ra.s1 * rb.s1 * sym_at(w.Hp, ra.c1, rb.c1);
            if (rb.s2 != 0.0) s += ra.s1 * rb.s2 * sym_at(w.Hp, ra.c1, rb.c2);
            if (ra.s2 != 0.0) s += ra.s2 * rb.s1 * sym_at(w.Hp, ra.c2, rb.c1);
            if (ra.s2 != 0.0 && rb.s2 != 0.0) s += ra.s2 * rb.s2 * sym_at(w.Hp, ra.c2, rb.c2);
            S[tri(a) + b] = s;
        }
    }
    CTA_FOR(a, 0, k) {
        const RowDef ra = row_def(w.aidx[a], p.mu, p.fz_min);
        w.t1[a] = ra.s1 * w.u0[ra.c1] + ra.s2 * w.u0[ra.c2] - ra.b;
    }
    cta_sync(c);
    if (small_chol_solve(c, S, k, w.t1, &w.isc[4])) return 1;
    CTA_FOR(r, 0, m) w.lam[r] = 0.0;
    cta_sync(c);
    CTA_FOR(a, 0, k) w.lam[w.aidx[a]] = w.t1[a];
    cta_sync(c);
    At_lam(c, p, w.lam, w.t2, nf);
    cta_sync(c);
    symv_packed(c, w.Hp, n, w.t2, w.hx);
    CTA_FOR(i, 0, n) w.x[i] = w.u0[i] - w.hx[i];
    cta_sync(c);
    return 0;
}

// ----------------------------------------------------------------------------------------------
// Exact active-set solve on the Schur complement.
//   Phase 1 (primal-dual active set): the working set is re-chosen wholesale from lam + viol;
//            converges in a handful of iterations when it converges, but may cycle.
//   Phase 2 (single exchange): from the last consistent set, drop the most negative multiplier,
//            else add the most violated row; one change per solve, monotone in practice.
// Start: w.x (primal guess), w.lam (multiplier guess).  On success w.x / w.lam hold the KKT point.
// Returns the number of solves (>0) on convergence, 0 if it gave up (caller falls back to ADMM).
// ----------------------------------------------------------------------------------------------
CMPC_HD int solve_active_set(const Cta& c, const Params& p, Ws& w, int n, int nf, int kcap, int* n_active) {
    const int m = 5 * nf;
    const double tol = 1e-10;
    const int max_total = p.pdas_max_iter + 8 * p.pdas_max_iter + 32;
    CTA_FOR(r, 0, m) { w.act_prev[r] = 0; w.act_prev2[r] = 2; }
    cta_sync(c);
    int single = 0;
    for (int it = 1; it <= max_total; ++it) {
        if (!single) {
            // candidate set from s = lam + viol, at most one of each opposite face pair
            CTA_FOR(f, 0, nf) {
                double v[5];
                foot_viol(w.x, f, p.mu, p.fz_min, v);
                double s[5];
                for (int t = 0; t < 5; ++t) s[t] = w.lam[5 * f + t] + v[t];
                w.act[5 * f] = s[0] > tol;
                w.act[5 * f + 1] = (s[1] > tol) && (s[1] >= s[2]);
                w.act[5 * f + 2] = (s[2] > tol) && (s[2] > s[1]);
                w.act[5 * f + 3] = (s[3] > tol) && (s[3] >= s[4]);
                w.act[5 * f + 4] = (s[4] > tol) && (s[4] > s[3]);
            }
            cta_sync(c);
            if (c.tid == 0) {
                int same = 1, same2 = 1, k = 0;
                for (int r = 0; r < m; ++r) {
                    if (w.act[r] != w.act_prev[r]) same = 0;
                    if (w.act[r] != w.act_prev2[r]) same2 = 0;
                    k += w.act[r];
                }
                w.isc[1] = k;
                w.isc[2] = (it > 1 && same) ? 1 : 0;
                w.isc[3] = ((it > 2 && same2 && !same) || it > p.pdas_max_iter) ? 1 : 0;   // cycle / budget
            }
            cta_sync(c);
            if (w.isc[2]) { *n_active = w.isc[1]; return it - 1; }
            if (w.isc[3]) {
                if (it == 1) return 0;
                single = 1;   // x, lam belong to act_prev: continue from there, one change at a time
                CTA_FOR(r, 0, m) w.act[r] = w.act_prev[r];
                cta_sync(c);
            }
        }
        if (single) {
            CTA_FOR(f, 0, nf) {
                double v[5];
                foot_viol(w.x, f, p.mu, p.fz_min, v);
                for (int t = 0; t < 5; ++t) w.viol[5 * f + t] = v[t];
            }
            cta_sync(c);
            if (c.tid == 0) {
                int drop = -1, add = -1, k = 0;
                double worst = -tol, most = 1e-9;
                for (int r = 0; r < m; ++r) {
                    if (w.act[r]) {
                        ++k;
                        if (w.lam[r] < worst) { worst = w.lam[r]; drop = r; }
                    } else {
                        const int t = r % 5;
                        // never activate the face opposite to an active one (rows would be dependent)
                        const int opp = (t == 0) ? -1 : (((t - 1) ^ 1) + 1);
                        if (opp >= 0 && w.act[r - t + opp]) continue;
                        if (w.viol[r] > most) { most = w.viol[r]; add = r; }
                    }
                }
                if (drop >= 0) { w.act[drop] = 0; --k; }
                else if (add >= 0) { w.act[add] = 1; ++k; }
                w.isc[1] = k;
                w.isc[2] = (drop < 0 && add < 0) ? 1 : 0;
            }
            cta_sync(c);
            if (w.isc[2]) { *n_active = w.isc[1]; return it - 1 > 0 ? it - 1 : 1; }
        }
        CTA_FOR(r, 0, m) { w.act_prev2[r] = w.act_prev[r]; w.act_prev[r] = w.act[r]; }
        cta_sync(c);
        if (working_set_solve(c, p, w, n, nf, kcap)) return 0;
    }
    return 0;
}

// ----------------------------------------------------------------------------------------------
// Roll-out, co-states and the KKT residuals of the *original* problem, independent of any factor:
//   x_{k+1} = A_d x_k + B_d[k] u_k + g_d;   nu_{N-1} = -2Q(x_N - xref_N),  nu_{k-1} = -2Q(x_k - xref_k) + A_d' nu_k
//   dJ/du_v = 2 R u_v - Bcol[v]' nu_{k(v)}   
// Leaves X in w.X, nu in w.q, the gradient of the objective in w.hx.  Returns the objective value in
// the reference's convention (constant xref'Q xref dropped, centroidal_mpc.py:248-253).
// ----------------------------------------------------------------------------------------------
CMPC_HD double rollout_costate(const Cta& c, const Params& p, const QpIn& in, Ws& w, int n, const double* x) {
    const int N = in.N;
    for (int k = 0; k < N; ++k) {
        const double* prev = k ? (w.X + (k - 1) * 12) : w.x0;
        CTA_FOR(r, 0, 12) {
            double s = w.gd[r];
            for (int j = 0; j < 12; ++j) s += w.Ad[r * 12 + j] * prev[j];
            for (int v = w.vstart[k]; v < w.vstart[k + 1]; ++v) s += w.Bcol[v * 12 + r] * x[v];
            w.X[k * 12 + r] = s;
        }
        cta_sync(c);
    }
    for (int k = N - 1; k >= 0; --k) {
        CTA_FOR(r, 0, 12) {
            double s = -2.0 * p.Q[r] * (w.X[k * 12 + r] - in.x_ref[r * N + k]);
            if (k < N - 1) for (int j = 0; j < 12; ++j) s += w.Ad[j * 12 + r] * w.q[(k + 1) * 12 + j];
            w.q[k * 12 + r] = s;
        }
        cta_sync(c);
    }
    double part = 0.0;
    CTA_FOR(v, 0, n) {
        const int j = v / 3, comp = v - 3 * j;
        const int k = w.fk[j];
        double s = 0.0;
        for (int r = 0; r < 12; ++r) s += w.Bcol[v * 12 + r] * w.q[k * 12 + r];
        const double Rv = p.R[3 * w.fl[j] + comp];
        w.hx[v] = 2.0 * Rv * x[v] - s;
        part += Rv * x[v] * x[v];
    }
    CTA_FOR(i, 0, 12 * N) {
        const int k = i / 12, r = i - 12 * k;
        const double xr = in.x_ref[r * N + k];
        const double d = w.X[i] - xr;
        part += p.Q[r] * (d * d - xr * xr);   // = 1/2 w'Hw + g'w of the reference's sparse QP (CasADi "cost")
    }
    return cta_sum(c, part, w.red);
}

// ----------------------------------------------------------------------------------------------
// OSQP-style ADMM on the reduced QP (rows l <= A x <= u with A = [ +e_z ; pyramid faces ]).
// M = (H + sigma I + rho A'A)^-1 explicit in w.Hp.  Duals w.yv in OSQP sign convention.
// ----------------------------------------------------------------------------------------------
CMPC_HD void admm_rows(const double* x, int f, double mu, double z[5]) {
    const double fx = x[3 * f], fy = x[3 * f + 1], fz = x[3 * f + 2];
    z[0] = fz;
    z[1] = fx - mu * fz;
    z[2] = -fx - mu * fz;
    z[3] = fy - mu * fz;
    z[4] = -fy - mu * fz;
}

CMPC_HD double admm_clip(int t, double v, double fz_min) {
    return t == 0 ? fmax(v, fz_min) : fmin(v, 0.0);
}

// A' y for OSQP-sign duals
CMPC_HD void At_y(const double* l, double mu, double out[3]) {
    out[0] = l[1] - l[2];
    out[1] = l[3] - l[4];
    out[2] = l[0] - mu * (l[1] + l[2] + l[3] + l[4]);
}

struct AdmmResult { int iters; int status; double rho; double rp, rd; int nfac; };

// Forward declaration: (re)build P, H + shift, invert in place.  Returns 0 on success.
CMPC_HD int factor_inverse(const Cta& c, const Params& p, const QpIn& in, Ws& w, int n, double sigma, double rho,
                           bool need_cost_to_go);

CMPC_HD AdmmResult admm(const Cta& c, const Params& p, const QpIn& in, Ws& w, int n, int nf, double rho,
                        double eps_abs, double eps_rel, int max_iter) {
    AdmmResult res;
    res.iters = 0; res.status = ST_MAX_ITER; res.rho = rho; res.rp = 0; res.rd = 0; res.nfac = 1;
    const int m = 5 * nf;
    const double dz = 1.0 + 4.0 * p.mu * p.mu;
    if (factor_inverse(c, p, in, w, n, p.sigma, rho, true)) { res.status = ST_NON_CVX; return res; }
    // hx = H x for the starting point (gradient routine gives H x + g)
    rollout_costate(c, p, in, w, n, w.x);
    CTA_FOR(i, 0, n) w.hx[i] -= w.g[i];
    CTA_FOR(f, 0, nf) {
        double zz[5];
        admm_rows(w.x, f, p.mu, zz);
        for (int t = 0; t < 5; ++t) w.z[5 * f + t] = admm_clip(t, zz[t], p.fz_min);
    }
    cta_sync(c);
    for (int it = 1; it <= max_iter; ++it) {
        // rhs = sigma x - g + A'(rho z - y)
        CTA_FOR(f, 0, nf) {
            double tmp[5], a[3];
            for (int t = 0; t < 5; ++t) tmp[t] = rho * w.z[5 * f + t] - w.yv[5 * f + t];
            At_y(tmp, p.mu, a);
            for (int cc = 0; cc < 3; ++cc) w.t1[3 * f + cc] = p.sigma * w.x[3 * f + cc] - w.g[3 * f + cc] + a[cc];
        }
        cta_sync(c);
        symv_packed(c, w.Hp, n, w.t1, w.t2);   // x~ = M rhs
        // H x~ = rhs - sigma x~ - rho D x~ ; relax
        CTA_FOR(i, 0, n) {
            const double d = (i % 3 == 2) ? dz : 2.0;
            const double hxt = w.t1[i] - p.sigma * w.t2[i] - rho * d * w.t2[i];
            w.hx[i] = p.alpha * hxt + (1.0 - p.alpha) * w.hx[i];
        }
        CTA_FOR(f, 0, nf) {
            double zt[5];
            admm_rows(w.t2, f, p.mu, zt);
            for (int t = 0; t < 5; ++t) {
                const int r = 5 * f + t;
                const double zh = p.alpha * zt[t] + (1.0 - p.alpha) * w.z[r];
                const double zn = admm_clip(t, zh + w.yv[r] / rho, p.fz_min);
                w.yv[r] += rho * (zh - zn);
                w.z[r] = zn;
            }
        }
        cta_sync(c);
        CTA_FOR(i, 0, n) w.x[i] = p.alpha * w.t2[i] + (1.0 - p.alpha) * w.x[i];
        cta_sync(c);
        res.iters = it;
        const bool check = (it % p.check_termination == 0) || it == max_iter;
        const bool adapt = p.adaptive_rho_interval > 0 && (it % p.adaptive_rho_interval == 0);
        if (!(check || adapt)) continue;
        // residuals (unscaled): r_p = |Ax - z|, r_d = |Hx + g + A'y|
        double rp = 0, nAx = 0, nz = 0, rd = 0, nHx = 0, nAty = 0, ng = 0;
        CTA_FOR(f, 0, nf) {
            double ax[5], a[3];
            admm_rows(w.x, f, p.mu, ax);
            for (int t = 0; t < 5; ++t) {
                rp = fmax(rp, fabs(ax[t] - w.z[5 * f + t]));
                nAx = fmax(nAx, fabs(ax[t]));
                nz = fmax(nz, fabs(w.z[5 * f + t]));
            }
            At_y(w.yv + 5 * f, p.mu, a);
            for (int cc = 0; cc < 3; ++cc) {
                const int i = 3 * f + cc;
                rd = fmax(rd, fabs(w.hx[i] + w.g[i] + a[cc]));
                nHx = fmax(nHx, fabs(w.hx[i]));
                nAty = fmax(nAty, fabs(a[cc]));
                ng = fmax(ng, fabs(w.g[i]));
            }
        }
        rp = cta_max(c, rp, w.red);
        rd = cta_max(c, rd, w.red);
        const double np_ = cta_max(c, fmax(nAx, nz), w.red);
        const double nd_ = cta_max(c, fmax(fmax(nHx, nAty), ng), w.red);
        res.rp = rp; res.rd = rd;
        if (check && rp <= eps_abs + eps_rel * np_ && rd <= eps_abs + eps_rel * nd_) {
            res.status = ST_SOLVED;
            break;
        }
        if (adapt && it < max_iter) {
            const double a = rp / fmax(np_, 1e-30), b = rd / fmax(nd_, 1e-30);
            double rn = rho * sqrt(a / fmax(b, 1e-30));
            rn = fmin(fmax(rn, 1e-6), 1e6);
            if (rn > 5.0 * rho || rn < 0.2 * rho) {
                rho = rn;
                res.rho = rho;
                ++res.nfac;
                // the relaxation recursion for hx stays valid (it does not depend on rho);
                // factor_inverse only touches Hp, P, T144 and t1
                if (factor_inverse(c, p, in, w, n, p.sigma, rho, true)) { res.status = ST_NON_CVX; return res; }
            }
        }
        }
    (void)m;
    res.rho = rho;
    return res;
}

CMPC_HD int factor_inverse(const Cta& c, const Params& p, const QpIn& in, Ws& w, int n, double sigma, double rho,
                           bool need_cost_to_go) {
    if (need_cost_to_go) cost_to_go(c, p, in.N, w, false);
    build_H(c, p, in.N, w, n, sigma, rho);
    if (chol_packed(c, w.Hp, n, n, &w.isc[5])) return 1;
    trtri_packed(c, w.Hp, n, w.t1);
    lauum_packed(c, w.Hp, n, w.t1);
    return 0;
}

// ----------------------------------------------------------------------------------------------
// The whole per-robot solve.  Output pointers are per-robot views into global memory.
// ----------------------------------------------------------------------------------------------
struct QpOut {
    double* u;       // 12N in/out (warm start in, solution out)
    double* y;       // 28N in/out
    double* rho;     // 1 in/out
    double* X;       // 12N or null
    double* nu;      // 12N or null
    int32_t* status;
    int32_t* iters;
    double* stats;   // NSTAT
};

CMPC_HD void write_failure(const Cta& c, const QpIn& in, QpOut& o, int status, int nf) {
    CTA_FOR(i, 0, 12 * in.N) o.u[i] = 0.0;
    if (c.tid == 0) {
        *o.status = status;
        *o.iters = 0;
        for (int i = 0; i < NSTAT; ++i) o.stats[i] = 0.0;
        o.stats[3] = 3.0 * nf;
    }
}

CMPC_HD void solve_one(const Cta& c, const Params& p, const QpIn& in, QpOut& o, Ws& w, int nfmax, int warm) {
    const int N = in.N;
    const int nf = setup_feet(c, in, w, nfmax);
    if (nf > nfmax) { write_failure(c, in, o, ST_TOO_MANY_FEET, nf); return; }
    const int n = 3 * nf, m = 5 * nf;
    const int kcap = kcap_for(N);
    int status = ST_SOLVED, iters = 0, path = PATH_UNCONSTRAINED, as_iters = 0, n_active = 0;
    double rho = (warm && o.rho && *o.rho > 0.0) ? *o.rho : p.rho0;

    load_dynamics(c, in, w, n);
    free_response(c, in, w);
    cost_to_go(c, p, N, w, true);
    gradient_vec(c, w, n);

    // warm-start state (centroidal_mpc.py:92-95): previous forces and multipliers, re-indexed
    if (warm) {
        CTA_FOR(v, 0, n) { const int j = v / 3; w.x[v] = o.u[12 * w.fk[j] + 3 * w.fl[j] + (v - 3 * j)]; }
        CTA_FOR(f, 0, nf) {
            const int k = w.fk[f], leg = w.fl[f];
            const double yb = o.y[12 * k + 3 * leg + 2];
            w.yv[5 * f] = fmin(yb, 0.0);
            w.lam[5 * f] = fmax(-yb, 0.0);
            for (int t = 1; t < 5; ++t) {
                const double yf = fmax(o.y[12 * N + 16 * k + 4 * leg + (t - 1)], 0.0);
                w.yv[5 * f + t] = yf;
                w.lam[5 * f + t] = yf;
            }
        }
    } else {
        CTA_FOR(v, 0, n) w.x[v] = 0.0;
        CTA_FOR(r, 0, m) { w.yv[r] = 0.0; w.lam[r] = 0.0; }
    }
    cta_sync(c);

    bool done = (n == 0);
    bool need_admm = false;
    if (!done && p.mode == 1) {
        // ---- unconstrained minimiser: Cholesky with g as an extra row, then one back-substitution
        build_H(c, p, N, w, n, 0.0, 0.0);
        CTA_FOR(j, 0, n) w.Hp[tri(n) + j] = w.g[j];
        cta_sync(c);
        if (chol_packed(c, w.Hp, n, n + 1, &w.isc[5])) { write_failure(c, in, o, ST_NON_CVX, nf); return; }
        CTA_FOR(j, 0, n) w.u0[j] = -w.Hp[tri(n) + j];
        cta_sync(c);
        {
            Cta wc; wc.tid = c.tid; wc.nt = c.nt < 32 ? c.nt : 32; wc.warp = 1;
            if (c.tid < wc.nt) backsolve_warp(wc, w.Hp, n, w.u0);
            cta_sync(c);
        }
        const double mv = all_viol(c, p, w, w.u0, nf);
        if (mv <= 1e-9) {
            CTA_FOR(i, 0, n) w.x[i] = w.u0[i];
            CTA_FOR(r, 0, m) w.lam[r] = 0.0;
            cta_sync(c);
            done = true;
        } else {
            // ---- explicit inverse, then primal-dual active set
            trtri_packed(c, w.Hp, n, w.t1);
            lauum_packed(c, w.Hp, n, w.t1);
            if (!warm) { CTA_FOR(i, 0, n) w.x[i] = w.u0[i]; cta_sync(c); }
            as_iters = solve_active_set(c, p, w, n, nf, kcap, &n_active);
            if (as_iters > 0) { done = true; path = PATH_ACTIVE_SET; }
            else need_admm = true;
        }
    } else if (!done) {
        need_admm = true;
    }

    if (need_admm) {
        path = PATH_ADMM;
        if (p.mode == 1) {
            // fallback start: previous warm start if any, else the clipped unconstrained point
            CTA_FOR(r, 0, m) w.yv[r] = 0.0;
            CTA_FOR(f, 0, nf) {
                double fz = fmax(w.u0[3 * f + 2], p.fz_min);
                const double lim = p.mu * fz;
                w.x[3 * f] = fmin(fmax(w.u0[3 * f], -lim), lim);
                w.x[3 * f + 1] = fmin(fmax(w.u0[3 * f + 1], -lim), lim);
                w.x[3 * f + 2] = fz;
            }
            cta_sync(c);
        }
        const double ea = (p.mode == 1) ? fmin(p.eps_abs, 1e-6) : p.eps_abs;
        const double er = (p.mode == 1) ? fmin(p.eps_rel, 1e-6) : p.eps_rel;
        AdmmResult r = admm(c, p, in, w, n, nf, rho, ea, er, p.max_iter);
        if (r.status == ST_NON_CVX) { write_failure(c, in, o, ST_NON_CVX, nf); return; }
        status = r.status;
        iters = r.iters;
        rho = r.rho;
        // multipliers in the a'x <= b form
        CTA_FOR(f, 0, nf) {
            w.lam[5 * f] = fmax(-w.yv[5 * f], 0.0);
            for (int t = 1; t < 5; ++t) w.lam[5 * f + t] = fmax(w.yv[5 * f + t], 0.0);
        }
        cta_sync(c);
        if (p.mode == 1 || p.polish) {
            // polish: exact active-set solve started from the ADMM point (OSQP's polish idea,
            // centroidal_mpc.py:28 has it switched off in the reference)
            // Re-factor H (no shift) with g appended to regain u0 and M = H^-1
            cost_to_go(c, p, N, w, false);
            build_H(c, p, N, w, n, 0.0, 0.0);
            CTA_FOR(j, 0, n) w.Hp[tri(n) + j] = w.g[j];
            cta_sync(c);
            if (!chol_packed(c, w.Hp, n, n + 1, &w.isc[5])) {
                CTA_FOR(j, 0, n) w.u0[j] = -w.Hp[tri(n) + j];
                cta_sync(c);
                Cta wc; wc.tid = c.tid; wc.nt = c.nt < 32 ? c.nt : 32; wc.warp = 1;
                if (c.tid < wc.nt) backsolve_warp(wc, w.Hp, n, w.u0);
                cta_sync(c);
                trtri_packed(c, w.Hp, n, w.t1);
                lauum_packed(c, w.Hp, n, w.t1);
                // keep the ADMM point in case the polish fails
                CTA_FOR(i, 0, n) w.z[i] = w.x[i];           // z (5nf >= n) is free after ADMM
                CTA_FOR(r, 0, m) w.yv[r] = w.lam[r];
                cta_sync(c);
                const int ai = solve_active_set(c, p, w, n, nf, kcap, &n_active);
                if (ai > 0) { path = PATH_ADMM_POLISH; as_iters = ai; status = ST_SOLVED; }
                else {
                    CTA_FOR(i, 0, n) w.x[i] = w.z[i];
                    CTA_FOR(r, 0, m) w.lam[r] = w.yv[r];
                    cta_sync(c);
                }
            }
        }
    }

    // ---- epilogue: residuals from first principles, outputs in the reference's layouts
    double obj = 0.0, rp = 0.0, rd = 0.0;
    if (n > 0) {
        obj = rollout_costate(c, p, in, w, n, w.x);
        At_lam(c, p, w.lam, w.t2, nf);
        cta_sync(c);
        double a = 0.0, b = 0.0;
        int na = 0;
        CTA_FOR(i, 0, n) a = fmax(a, fabs(w.hx[i] + w.t2[i]));
        CTA_FOR(f, 0, nf) {
            double v[5];
            foot_viol(w.x, f, p.mu, p.fz_min, v);
            for (int t = 0; t < 5; ++t) { b = fmax(b, v[t]); if (w.lam[5 * f + t] > 0.0) ++na; }
        }
        rd = cta_max(c, a, w.red);
        rp = fmax(cta_max(c, b, w.red), 0.0);
        n_active = (int)(cta_sum(c, (double)na, w.red) + 0.5);
    } else {
        // no stance foot at all: x_k is the free response
        CTA_FOR(i, 0, 12 * N) { const int k = i / 12, r = i % 12; w.X[i] = w.E[i] + in.x_ref[r * N + k]; }
        cta_sync(c);
        for (int k = N - 1; k >= 0; --k) {
            CTA_FOR(r, 0, 12) {
                double s = -2.0 * p.Q[r] * w.E[k * 12 + r];
                if (k < N - 1) for (int j = 0; j < 12; ++j) s += w.Ad[j * 12 + r] * w.q[(k + 1) * 12 + j];
                w.q[k * 12 + r] = s;
            }
            cta_sync(c);
        }
    }
    // exact paths must certify themselves: residuals of the original problem, not of the factor
    if (status == ST_SOLVED && path != PATH_ADMM && (rp > 1e-6 || rd > 1e-6)) status = ST_INACCURATE;

    // forces and duals, full reference layouts; eliminated (swing) variables get u = 0 and the
    // multiplier that closes stationarity:  y_box = Bd[k][:,v]' nu_k   (2 R u = 0, no friction dual)
    CTA_FOR(i, 0, 12 * N) o.u[i] = 0.0;
    CTA_FOR(i, 0, 28 * N) o.y[i] = 0.0;
    cta_sync(c);
    CTA_FOR(v, 0, n) { const int j = v / 3; o.u[12 * w.fk[j] + 3 * w.fl[j] + (v - 3 * j)] = w.x[v]; }
    CTA_FOR(f, 0, nf) {
        const int k = w.fk[f], leg = w.fl[f];
        o.y[12 * k + 3 * leg + 2] = -w.lam[5 * f];
        for (int t = 1; t < 5; ++t) o.y[12 * N + 16 * k + 4 * leg + (t - 1)] = w.lam[5 * f + t];
    }
    CTA_FOR(i, 0, 12 * N) {
        const int k = i / 12, jj = i - 12 * k, leg = jj / 3, comp = jj - 3 * leg;
        if (!mask_bit(in.mask, N, leg, k)) {
            double col[12];
            any_Bd_col(in, k, leg, comp, col);
            double s = 0.0;
            for (int r = 0; r < 12; ++r) s += col[r] * w.q[k * 12 + r];
            o.y[i] = s;
        }
    }
    if (o.X) { CTA_FOR(i, 0, 12 * N) o.X[i] = w.X[i]; }
    if (o.nu) { CTA_FOR(i, 0, 12 * N) o.nu[i] = w.q[i]; }
    if (c.tid == 0) {
        if (o.rho) *o.rho = rho;
        *o.status = status;
        *o.iters = iters;
        o.stats[0] = rp;
        o.stats[1] = rd;
        o.stats[2] = obj;
        o.stats[3] = (double)n;
        o.stats[4] = (double)n_active;
        o.stats[5] = rho;
        o.stats[6] = (double)as_iters;
        o.stats[7] = (double)path;
    }
    cta_sync(c);
}

// ----------------------------------------------------------------------------------------------
// Diagnostic: dense H (12N x 12N, all variables free) and g for the parity tests (level L1).
// ----------------------------------------------------------------------------------------------
CMPC_HD void build_dense_one(const Cta& c, const Params& p, const QpIn& in_, Ws& w, int nfmax, double* H, double* g) {
    QpIn in = in_;
    in.mask = nullptr;   // every foot-step free
    const int N = in.N;
    const int nf = setup_feet(c, in, w, nfmax);
    const int n = 3 * nf;
    load_dynamics(c, in, w, n);
    free_response(c, in, w);
    cost_to_go(c, p, N, w, true);
    gradient_vec(c, w, n);
    build_H(c, p, N, w, n, 0.0, 0.0);
    CTA_FOR(e, 0, n * n) {
        const int i = e / n, j = e - i * n;
        H[e] = sym_at(w.Hp, i, j);
    }
    CTA_FOR(i, 0, n) g[i] = w.g[i];
    cta_sync(c);
}

}  // namespace cmpc
