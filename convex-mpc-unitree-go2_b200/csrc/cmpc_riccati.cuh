// cmpc_riccati.cuh -- nominal pre-pass: the unconstrained minimiser of one robot's QP by a Riccati recursion,
// one WARP per robot.
//
// Why.  On the headline workload 69 % of the QPs have no active inequality at the optimum (nominal trot:
// SURVEY.md section 8 f4).  For those the optimum of the reference QP (centroidal_mpc.py:69-120) is the
// unconstrained minimiser over the stance forces, and because the problem is a time-varying LQ problem in
//     x_{k+1} = A_d x_k + B_d[k] u_k + g_d        (com_trajectory.py:221-286)
// that minimiser follows from a backward Riccati sweep over the N steps with 12 x 12 blocks: ~0.15 MFLOP and a
// few KB of state per robot instead of the 0.9 MFLOP / 106 KB of the condensed factorisation, so a warp can own
// a robot and a dozen robots are in flight per SM (the condensed path holds two).  A robot whose Riccati
// solution violates a friction-pyramid / fz_min row (or whose certificate is not clean) is appended to a
// work-list and goes through the exact condensed active-set kernel (cmpc_fast.cuh) unchanged; everybody else
// is finished here, with the same outputs: forces, states by roll-out, co-states, box multipliers of the
// eliminated swing forces, KKT residuals and objective recomputed from first principles.
//
// Recursion (value function  V_k(x) = x'P x + 2 p'x,  P_N = Q, p_N = -Q xref_N;  stance columns only):
//     T = B'P      G = R + T B      S = T A      q = P g + p
//     G = L L'     W = inv(L)       K = W'W S    kff = W'W B'q         u_k = -(K x_k + kff)
//     P <- Q + A'P A - S'K          p <- -Q xref_k + A'q - S'kff
// The gains K, kff (3 nf x 12 doubles per robot) go through a per-warp global scratch slot that stays in L2.
//
// The code is written against the warp-sized Cta of cmpc_core.cuh ("virtual lanes" loops + cta_sync), so the
// same source runs under the host emulation of the CPU test-suite (tests/_emul) with a one-thread group.
#pragma once
#include "cmpc_core.cuh"

#if !defined(__CUDA_ARCH__) && defined(CMPC_RIC_DEBUG)
#include <cstdio>
#define RIC_DBG(...) fprintf(stderr, __VA_ARGS__)
#else
#define RIC_DBG(...)
#endif

namespace cmpc {
namespace ric {

constexpr int LD = 14;          // leading dimension of the 12 x 12 work matrices (row stride 112 B: 16-byte aligned
                                // rows, and the twelve row starts fall in distinct bank groups but for r / r+8)
constexpr int MAT = 12 * LD;

struct WsR {
    double* RF;     // 12N lever arms as given, index (leg*3 + a)*N + k
    double* XR;     // 12N reference, index i*12 + r  (state x_{i+1})
    double* X;      // 12N rolled-out states
    double* NU;     // 12N co-states
    double* U;      // 12N forces in the output layout 12k + 3 leg + comp (0 for swing legs)
    double* P;      // cost-to-go matrix
    double* Bm;     // B_d[k], stance columns compacted to the front, zero elsewhere
    double* TK;     // T = B'P, later the gain K
    double* GY;     // G -> L (lower) ; later Y = W S
    double* S;      // S = T A
    double* W;      // inv(L), later N = P A
    double* UW;     // 4 x 18: U and W of the four legs at the current step
    double* pv;     // 12 value-function gradient p
    double* q;      // 12
    double* bq;     // 12  B'q -> W B'q
    double* kff;    // 12
    double* xk;     // 12 current state of the forward sweep
    double* uk;     // 12 compact forces of the current step
    double* x0;     // 12
    double* dinv;   // 12  1 / L_aa
    double* red;    // 40
    DynCommon* dyn;
    int* vstart;    // N+1: first compact variable of every step
};

CMPC_HD size_t ws_carve_ric(WsR& w, unsigned char* base, int N) {
    double* p = reinterpret_cast<double*>(base);
    auto take = [&](size_t n) { double* r = p; p += (n + 1) & ~(size_t)1; return r; };
    w.RF = take((size_t)12 * N);
    w.XR = take((size_t)12 * N);
    w.X = take((size_t)12 * N);
    w.NU = take((size_t)12 * N);
    w.U = take((size_t)12 * N);
    w.P = take(MAT); w.Bm = take(MAT); w.TK = take(MAT); w.GY = take(MAT); w.S = take(MAT); w.W = take(MAT);
    w.UW = take(72);
    w.pv = take(12); w.q = take(12); w.bq = take(12); w.kff = take(12); w.xk = take(12); w.uk = take(12);
    w.x0 = take(12); w.dinv = take(12);
    w.red = take(40);
    w.dyn = reinterpret_cast<DynCommon*>(take((sizeof(DynCommon) + 7) / 8));
    int* ip = reinterpret_cast<int*>(p);
    w.vstart = ip;
    ip += (N + 1 + 3) & ~3;
    return (size_t)(reinterpret_cast<unsigned char*>(ip) - base);
}

// doubles of gain scratch one robot needs (K rows + kff)
CMPC_HD size_t gain_doubles(int nfmax) { return (size_t)3 * nfmax * 13; }

// O[r][:] = sum_{k<kmax} Xop[r][k] M[k][:]  for r < nrows;  Xop = X or X^T.  Lane -> (row r = lane & 15,
// column half h = lane >> 4): six outputs per lane, the M row segment is a broadcast load.
CMPC_HD void mm12(const Cta& c, double* O, const double* X, const double* M, int nrows, int kmax, bool transX) {
    CTA_FOR(vl, 0, 32) {
        const int r = vl & 15, h = vl >> 4;
        if (r >= nrows) continue;
        double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        for (int k = 0; k < kmax; ++k) {
            const double x = transX ? X[k * LD + r] : X[r * LD + k];
            const double* m = M + k * LD + 6 * h;
#pragma unroll
            for (int cc = 0; cc < 6; ++cc) acc[cc] += x * m[cc];
        }
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) O[r * LD + 6 * h + cc] = acc[cc];
    }
}

// (v A)[c] for a row vector v of length 12 (A_d = I + dt A_c):  columns 6..8 gain dt v[0..2], columns 9..11
// gain dt v[3..5] Rz^T
CMPC_HD double row_times_A(const DynCommon& d, const double* v, int st, int c) {
    double s = v[c * st];
    if (c >= 6 && c < 9) s += d.dt * v[(c - 6) * st];
    else if (c == 9) s += d.dt * (d.cy * v[3 * st] - d.sy * v[4 * st]);
    else if (c == 10) s += d.dt * (d.sy * v[3 * st] + d.cy * v[4 * st]);
    else if (c == 11) s += d.dt * v[5 * st];
    return s;
}

// (A^T v)[i] for a column vector v (stride st between entries): rows 6..8 gain dt v[0..2], rows 9..11 gain
// dt Rz v[3..5]   (A^T has the transposed coupling blocks)
CMPC_HD double At_times_col(const DynCommon& d, const double* v, int st, int i) { return row_times_A(d, v, st, i); }

// (A v)[i]: rows 0..2 gain dt v[6..8], rows 3..5 gain dt Rz^T v[9..11]
CMPC_HD double A_times_col(const DynCommon& d, const double* v, int i) {
    double s = v[i];
    if (i < 3) s += d.dt * v[6 + i];
    else if (i == 3) s += d.dt * (d.cy * v[9] + d.sy * v[10]);
    else if (i == 4) s += d.dt * (-d.sy * v[9] + d.cy * v[10]);
    else if (i == 5) s += d.dt * v[11];
    return s;
}

// stance legs of step k, compacted: returns m = 3 * count, legs[] holds the leg of every compact triple
CMPC_HD int stance_legs(const QpIn& in, int k, int legs[4]) {
    int cnt = 0;
    for (int leg = 0; leg < 4; ++leg)
        if (mask_bit(in.mask, in.N, leg, k)) legs[cnt++] = leg;
    for (int i = cnt; i < 4; ++i) legs[i] = -1;
    return 3 * cnt;
}

// UW of the four legs at step k (lanes 0..3), then the dense B_d[k] with the stance columns first
CMPC_HD void build_B(const Cta& c, const QpIn& in, WsR& w, int k, const int legs[4], int m) {
    const int N = in.N;
    const DynCommon& d = *w.dyn;
    CTA_FOR(leg, 0, 4) {
        double r[3];
        for (int a = 0; a < 3; ++a) r[a] = w.RF[(leg * 3 + a) * N + k];
        // U (rows of Rz^T W) then W = Iinv [r]x, row-major 3x3 each
        const double sk[9] = {0.0, -r[2], r[1], r[2], 0.0, -r[0], -r[1], r[0], 0.0};
        double Wm[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                Wm[i * 3 + j] = d.Iinv[i * 3] * sk[j] + d.Iinv[i * 3 + 1] * sk[3 + j] + d.Iinv[i * 3 + 2] * sk[6 + j];
        double* o = w.UW + 18 * leg;
        for (int j = 0; j < 3; ++j) {
            o[j] = d.cy * Wm[j] + d.sy * Wm[3 + j];
            o[3 + j] = -d.sy * Wm[j] + d.cy * Wm[3 + j];
            o[6 + j] = Wm[6 + j];
        }
        for (int i = 0; i < 9; ++i) o[9 + i] = Wm[i];
    }
    cta_sync(c);
    const double h = d.dt * d.dt / 2.0;
    CTA_FOR(e, 0, 144) {
        const int i = e / 12, a = e - 12 * i;
        double v = 0.0;
        if (a < m) {
            const int leg = legs[a / 3], cc = a % 3;
            const double* uw = w.UW + 18 * leg;
            if (i < 3) v = (i == cc) ? h * d.minv : 0.0;
            else if (i < 6) v = h * uw[(i - 3) * 3 + cc];
            else if (i < 9) v = (i - 6 == cc) ? d.dt * d.minv : 0.0;
            else v = d.dt * uw[9 + (i - 9) * 3 + cc];
        }
        w.Bm[i * LD + a] = v;
    }
    cta_sync(c);
}

CMPC_HD double rsqrt_d(double v) {
#if defined(__CUDA_ARCH__)
    return rsqrt(v);
#else
    return 1.0 / sqrt(v);
#endif
}

// In-place Cholesky of the leading m x m block of G (lower), then W = inv(L) (lower, zero above the diagonal,
// rows/columns >= m untouched).  Returns 1 if a pivot is not positive.
CMPC_HD int chol_inv_small(const Cta& c, double* G, double* W, double* dinv, int m) {
    int bad = 0;
    for (int j = 0; j < m; ++j) {
        const double dj = G[j * LD + j];
        if (!(dj > 0.0)) bad = 1;
        const double rj = rsqrt_d(dj > 0.0 ? dj : 1.0);
        cta_sync(c);                                         // everybody has read the pivot
        CTA_FOR(a, j, m) G[a * LD + j] *= rj;                // column j of L (the diagonal becomes sqrt(d))
        if (c.tid == 0) dinv[j] = rj;
        cta_sync(c);
        const int t = m - 1 - j;                             // trailing block: rows/cols j+1 .. m-1, lower part
        CTA_FOR(e, 0, t * t) {
            const int a = j + 1 + e / t, b = j + 1 + e % t;
            if (b <= a) G[a * LD + b] -= G[a * LD + j] * G[b * LD + j];
        }
        cta_sync(c);
    }
    // inverse, one column per lane by forward substitution
    CTA_FOR(b, 0, m) {
        for (int a = 0; a < m; ++a) {
            double s = (a == b) ? 1.0 : 0.0;
            if (a < b) { W[a * LD + b] = 0.0; continue; }
            for (int k = b; k < a; ++k) s -= G[a * LD + k] * W[k * LD + b];
            W[a * LD + b] = s * dinv[a];
        }
    }
    cta_sync(c);
    return bad;
}

// One robot.  Returns 1 when the robot is finished here (outputs written), 0 when it has to go through the
// condensed active-set path (nothing written).  `gains`: gain_doubles(nfmax) doubles of scratch.
CMPC_HD int riccati_one(const Cta& c, const Params& p, const QpIn& in, QpOut& o, WsR& w, int nfmax, int warm,
                        double* gains) {
    const int N = in.N;
    // ---- set-up: stage the record, dynamics constants, compact variable offsets
    if (c.tid == 0) {
        dyn_common(*w.dyn, in.x_ref, N, in.I_world, in.mass, in.dt);
        int nf = 0;
        for (int k = 0; k < N; ++k) {
            w.vstart[k] = 3 * nf;
            for (int leg = 0; leg < 4; ++leg) nf += mask_bit(in.mask, N, leg, k);
        }
        w.vstart[N] = 3 * nf;
    }
    CTA_FOR(i, 0, 12) w.x0[i] = in.x0[i];
    CTA_FOR(idx, 0, 12 * N) {
        const int r = idx / N, i = idx - r * N;
        w.XR[i * 12 + r] = in.x_ref[idx];
        w.RF[idx] = in.r_foot[idx];
    }
    cta_sync(c);
    const int n = w.vstart[N];
    if (n > 3 * nfmax || n == 0) return 0;
    const DynCommon& d = *w.dyn;
    double* Kst = gains;                   // (n x 12) gain rows
    double* kst = gains + (size_t)3 * nfmax * 12;
    const double gz2 = -9.81 * d.dt * d.dt / 2.0, gz8 = -9.81 * d.dt;   // g_d: entries 2 and 8

    // ---- backward sweep
    CTA_FOR(e, 0, 144) { const int i = e / 12, cc = e - 12 * i; w.P[i * LD + cc] = (i == cc) ? p.Q[i] : 0.0; }
    CTA_FOR(i, 0, 12) w.pv[i] = -p.Q[i] * w.XR[(N - 1) * 12 + i];
    cta_sync(c);
    int bad = 0;
    for (int k = N - 1; k >= 0; --k) {
        int legs[4];
        const int m = stance_legs(in, k, legs);
        const int voff = w.vstart[k];
        if (m > 0) {
            build_B(c, in, w, k, legs, m);
            // q = P g + p ;  T = B'P
            CTA_FOR(i, 0, 12) w.q[i] = w.P[i * LD + 2] * gz2 + w.P[i * LD + 8] * gz8 + w.pv[i];
            mm12(c, w.TK, w.Bm, w.P, m, 12, true);
            cta_sync(c);
            // G = T B + R ;  S = T A ;  bq = B'q
            mm12(c, w.GY, w.TK, w.Bm, m, 12, false);
            CTA_FOR(e, 0, m * 12) { const int a = e / 12, cc = e - 12 * a; w.S[a * LD + cc] = row_times_A(d, w.TK + a * LD, 1, cc); }
            CTA_FOR(a, 0, m) {
                double s = 0.0;
                for (int i = 0; i < 12; ++i) s += w.Bm[i * LD + a] * w.q[i];
                w.bq[a] = s;
            }
            cta_sync(c);
            CTA_FOR(a, 0, m) w.GY[a * LD + a] += p.R[3 * legs[a / 3] + a % 3];
            cta_sync(c);
            bad |= chol_inv_small(c, w.GY, w.W, w.dinv, m);
            // Y = W S ;  yv = W bq
            CTA_FOR(a, 0, m) {
                double yv = 0.0;
                for (int b = 0; b <= a; ++b) yv += w.W[a * LD + b] * w.bq[b];
                w.uk[a] = yv;                                // (uk is free during the backward sweep)
            }
            mm12(c, w.GY, w.W, w.S, m, m, false);
            cta_sync(c);
            // K = W'Y ;  kff = W'yv  -> shared memory (for the P update) and the gain scratch (for the forward sweep)
            mm12(c, w.TK, w.W, w.GY, m, m, true);
            CTA_FOR(a, 0, m) {
                double s = 0.0;
                for (int b = a; b < m; ++b) s += w.W[b * LD + a] * w.uk[b];
                w.kff[a] = s;
                kst[voff + a] = s;
            }
            cta_sync(c);
            CTA_FOR(e, 0, m * 12) { const int a = e / 12, cc = e - 12 * a; Kst[(size_t)(voff + a) * 12 + cc] = w.TK[a * LD + cc]; }
        } else {
            CTA_FOR(i, 0, 12) w.q[i] = w.P[i * LD + 2] * gz2 + w.P[i * LD + 8] * gz8 + w.pv[i];
            cta_sync(c);
        }
        if (k == 0) break;
        // N = P A (into W) ;  then  P <- Q + A'N - S'K ,  p <- -Q xref_k + A'q - S'kff
        CTA_FOR(e, 0, 144) { const int i = e / 12, cc = e - 12 * i; w.W[i * LD + cc] = row_times_A(d, w.P + i * LD, 1, cc); }
        cta_sync(c);
        CTA_FOR(vl, 0, 32) {
            const int i = vl & 15, h = vl >> 4;
            if (i >= 12) continue;
            double acc[6];
#pragma unroll
            for (int cc = 0; cc < 6; ++cc) acc[cc] = At_times_col(d, w.W + 6 * h + cc, LD, i) + ((i == 6 * h + cc) ? p.Q[i] : 0.0);
            for (int a = 0; a < m; ++a) {
                const double s = w.S[a * LD + i];
                const double* kr = w.TK + a * LD + 6 * h;
#pragma unroll
                for (int cc = 0; cc < 6; ++cc) acc[cc] -= s * kr[cc];
            }
#pragma unroll
            for (int cc = 0; cc < 6; ++cc) w.P[i * LD + 6 * h + cc] = acc[cc];
        }
        CTA_FOR(i, 0, 12) {
            double s = -p.Q[i] * w.XR[(k - 1) * 12 + i] + At_times_col(d, w.q, 1, i);
            for (int a = 0; a < m; ++a) s -= w.S[a * LD + i] * w.kff[a];
            w.pv[i] = s;
        }
        cta_sync(c);
    }
    if (bad) { RIC_DBG("bad pivot\n"); return 0; }

    // ---- forward sweep: u_k = -(K x_k + kff),  x_{k+1} = A x_k + B u_k + g
    CTA_FOR(i, 0, 12) w.xk[i] = w.x0[i];
    CTA_FOR(i, 0, 12 * N) w.U[i] = 0.0;
    cta_sync(c);
    for (int k = 0; k < N; ++k) {
        int legs[4];
        const int m = stance_legs(in, k, legs);
        const int voff = w.vstart[k];
        if (m > 0) {
            build_B(c, in, w, k, legs, m);
            CTA_FOR(a, 0, m) {
                const double* kr = Kst + (size_t)(voff + a) * 12;
                double s = kst[voff + a];
                for (int cc = 0; cc < 12; ++cc) s += kr[cc] * w.xk[cc];
                w.uk[a] = -s;
                w.U[12 * k + 3 * legs[a / 3] + a % 3] = -s;
            }
            cta_sync(c);
        }
        double xn = 0.0;
        CTA_FOR(i, 0, 12) {
            xn = A_times_col(d, w.xk, i) + (i == 2 ? gz2 : (i == 8 ? gz8 : 0.0));
            for (int a = 0; a < m; ++a) xn += w.Bm[i * LD + a] * w.uk[a];
            w.X[k * 12 + i] = xn;
        }
        cta_sync(c);
        CTA_FOR(i, 0, 12) w.xk[i] = w.X[k * 12 + i];
        cta_sync(c);
    }

    // ---- feasibility of the unconstrained minimiser: fz >= fz_min and the four pyramid faces
    double mv = -1e300;
    CTA_FOR(e, 0, 4 * N) {
        const int k = e >> 2, leg = e & 3;
        if (!mask_bit(in.mask, N, leg, k)) continue;
        const double* f = w.U + 12 * k + 3 * leg;
        const double lim = p.mu * f[2];
        mv = fmax(mv, fmax(p.fz_min - f[2], fmax(fabs(f[0]), fabs(f[1])) - lim));
    }
    mv = cta_max(c, mv, w.red);
    if (!(mv <= 1e-9)) { RIC_DBG("infeasible mv=%g\n", mv); return 0; }

    // ---- co-states nu_k = -2 Q (X_k - xref_k) + A' nu_{k+1}, stationarity, objective (first principles)
    double part = 0.0;
    CTA_FOR(idx, 0, 12 * N) {
        const int r = idx % 12;
        const double xr = w.XR[idx], dd = w.X[idx] - xr;
        part += p.Q[r] * (dd * dd - xr * xr);
    }
    for (int k = N - 1; k >= 0; --k) {
        CTA_FOR(i, 0, 12) {
            double s = -2.0 * p.Q[i] * (w.X[k * 12 + i] - w.XR[k * 12 + i]);
            if (k < N - 1) s += At_times_col(d, w.NU + (k + 1) * 12, 1, i);
            w.NU[k * 12 + i] = s;
        }
        cta_sync(c);
    }
    double rd = 0.0;
    const double h = d.dt * d.dt / 2.0;
    double* ybox = w.XR;                           // the reference is not needed any more: stage the box multipliers
    cta_sync(c);                                   // there (nothing may reach o.* before the robot is accepted)
    CTA_FOR(e, 0, 4 * N) {
        const int k = e >> 2, leg = e & 3;
        const int st = mask_bit(in.mask, N, leg, k);
        double r[3];
        for (int a = 0; a < 3; ++a) r[a] = w.RF[(leg * 3 + a) * N + k];
        const double sk[9] = {0.0, -r[2], r[1], r[2], 0.0, -r[0], -r[1], r[0], 0.0};
        double Wm[9], Um[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                Wm[i * 3 + j] = d.Iinv[i * 3] * sk[j] + d.Iinv[i * 3 + 1] * sk[3 + j] + d.Iinv[i * 3 + 2] * sk[6 + j];
        for (int j = 0; j < 3; ++j) {
            Um[j] = d.cy * Wm[j] + d.sy * Wm[3 + j];
            Um[3 + j] = -d.sy * Wm[j] + d.cy * Wm[3 + j];
            Um[6 + j] = Wm[6 + j];
        }
        const double* nu = w.NU + k * 12;
        for (int cc = 0; cc < 3; ++cc) {
            double s = h * d.minv * nu[cc] + d.dt * d.minv * nu[6 + cc];
            for (int q = 0; q < 3; ++q) s += h * Um[q * 3 + cc] * nu[3 + q] + d.dt * Wm[q * 3 + cc] * nu[9 + q];
            const int idx = 12 * k + 3 * leg + cc;
            if (st) {
                const double Rv = p.R[3 * leg + cc], f = w.U[idx];
                rd = fmax(rd, fabs(2.0 * Rv * f - s));
                part += Rv * f * f;
                ybox[idx] = (cc == 2) ? -0.0 : 0.0;
            } else {
                ybox[idx] = s;                     // box multiplier of the eliminated swing force: B_col' nu_k
            }
        }
    }
    rd = cta_max(c, rd, w.red);
    const double obj = cta_sum(c, part, w.red);
    if (!(rd <= 1e-6)) { RIC_DBG("rd=%g\n", rd); return 0; }                   // certificate not clean: let the exact path decide

    // ---- outputs in the reference's layouts
    CTA_FOR(i, 0, 12 * N) { o.u[i] = w.U[i]; o.y[i] = ybox[i]; }
    CTA_FOR(i, 0, 16 * N) o.y[12 * N + i] = 0.0;
    if (o.X) { CTA_FOR(i, 0, 12 * N) o.X[i] = w.X[i]; }
    if (o.nu) { CTA_FOR(i, 0, 12 * N) o.nu[i] = w.NU[i]; }
    if (c.tid == 0) {
        const double rho = (warm && o.rho && *o.rho > 0.0) ? *o.rho : p.rho0;
        if (o.rho) *o.rho = rho;
        *o.status = ST_SOLVED;
        *o.iters = 0;
        o.stats[0] = fmax(mv, 0.0);
        o.stats[1] = rd;
        o.stats[2] = obj;
        o.stats[3] = (double)n;
        o.stats[4] = 0.0;
        o.stats[5] = rho;
        o.stats[6] = 0.0;
        o.stats[7] = (double)PATH_RICCATI;
    }
    return 1;
}

}  // namespace ric
}  // namespace cmpc
