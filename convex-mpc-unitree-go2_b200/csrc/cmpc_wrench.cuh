// cmpc_wrench.cuh -- wrench-space projected Riccati + primal-dual active set: the whole QP of
// centroidal_mpc.py:69-120 (build + CasADi->OSQP solve) for one robot on FOUR threads ("quad"), eight robots per
// warp, every robot / stance pattern / working set running the same straight-line code.
//
// Structure used (NumPy twin with the derivation: oracle/wrench_riccati.py):
//   * B_d[k] = Bbar [C_1 .. C_4],  C_j = [I; W_j],  W_j = I_world^-1 [r_j]x  (com_trajectory.py:234-262): the forces
//     act on the state only through the 6-dimensional wrench  w = sum_j C_j f_j,
//       Bbar = [[h/m I, 0], [0, h Rz'], [dt/m I, 0], [0, dt I]],   h = dt^2/2,   A = I + dt E.
//   * Every inequality row (fz >= fz_min centroidal_mpc.py:163-170, friction faces :324-359, swing forces = 0
//     :150-161) touches one foot at one step, so a working set is eliminated foot by foot:
//       f_j = fhat_j - Pi_j C_j' mu,   Pi_j = Z (Z'RZ)^-1 Z' = diag(dx, dy, 0) + sg z z'   (0 for swing / pinned feet)
//     and the stage enters the recursion only through  Lam_k = sum_j C_j Pi_j C_j'  (6 x 6)  and  what_k = sum_j C_j fhat_j.
//   * Backward stage (cost-to-go x'Px + 2p'x, all factorizations 6 x 6):
//       Gbar = Bbar'P Bbar = L L',   Nn = I + L' Lam L = Ln Ln',   Phi = L^-T (I - Nn^-1) L^-1,   Gam = L Nn^-1 L^-1
//       P  <- Q + A'(P - (P Bbar) Phi (P Bbar)') A,      p <- -Q xref + A'(q - (P Bbar) Phi Bbar'q),   q = P ghat + p
//       mu_k = Ktil (A x_k) + kbar,   Ktil = Gam (P Bbar)',   kbar = Gam Bbar'q            (wrench co-state)
//   * Primal-dual active set: after the forward sweep every foot-step gets multipliers from its own 3 x 3 stationarity
//     system and the next working set is { rows with lam + violation > 0 } (the rule of solve_active_set_fast); the
//     loop ends when the set repeats.  A sweep costs the same whatever the size of the working set.
//
// Thread a of a quad owns the block row a in {p, rpy, v, omega} of P (3 x 12, registers) and, in its second role,
// foot a (its contribution to Lam_k, its force, multipliers and next working-set code).  Exchanges inside the quad go
// through ~2.9 KB of shared memory per robot and __syncwarp; the 6 x 6 factorizations are done redundantly by the four
// threads (no exchange, no idle lanes).  Gains go to an L2-resident scratch.  The certificate / remaining outputs of a
// robot whose working set has settled are written by finish_robot (inputs prefetched with cp.async).  Working sets
// that cycle under the primal-dual rule continue with single exchanges (one row dropped or added per sweep).
// The same source compiles for the host (tests/_emul): there the four threads of a quad run one after the other
// between the synchronisation points.
#pragma once
#include "cmpc_core.cuh"

namespace cmpc {
namespace wr {

enum { PATH_WRENCH = 5 };
enum { ST_PENDING = 3 };         // internal marker: working set settled by the sweep kernel, certificate kernel to follow
constexpr unsigned char SWING = 255;
constexpr int GAIN_D2 = 10;      // double2 slots per thread and stage: Ktil rows (9), kbar part (1)

#if defined(__CUDA_ARCH__)
#define WR_Q_BEGIN { const int q = qlane; TS& t = ts[0];
#define WR_Q_END }
#define WR_SYNC() __syncwarp()
#define WR_GQ 0          // e.gains already points at this thread's slots

#else
#define WR_Q_BEGIN for (int q = 0; q < 4; ++q) { TS& t = ts[q];
#define WR_Q_END }
#define WR_SYNC() ((void)0)
#define WR_GQ q

#endif

struct alignas(16) D2 { double x, y; };

// batch arrays (device pointers), layouts of include/cmpc.h
struct Bat {
    const double *x0, *x_ref, *r_foot, *I_world, *mass;
    const uint64_t* mask;
    double *u, *y, *rho, *X, *nu, *stats;
    int32_t *status, *iters;
    double dt;
    int N, W;
    const double* cst = nullptr;     // (B,12) cy, sy, 1/m, Iinv[9] formed ahead of the sweeps (robot_consts_kernel), or null
};

// CTA-wide tables (shared memory): indices depend on the thread, so they must not sit in the constant bank
struct Tab {
    double Q[12], R[12], Rinv[12];
    double sig[16];          // [leg][2 (fx tied) + (fy tied)] = 1 / (Rz + [fx tied] mu^2 Rx + [fy tied] mu^2 Ry)
};

// shared memory of one robot (sweep kernel)
struct Sh {
    double E1[72];           // backward: S = P Bbar, [row 12][6]; forward: x (12) | mu partials (24) | wrench partials (24) | candidates
    double E3[12];           // backward: p ; finish: co-state
    double D4[8];            // backward: d of the threads p, rpy
    double PK[4 * 36];       // backward: the block rows of P while the 6 x 6 work needs the registers, then the rows D A of the
                             // threads p, rpy for their partners v, omega; finish: prefetched inputs of a stage, 2 slots x 4 threads x 18
                             // (x 3, xref 3, lever arm 3, force 3, duals 5)
    double cst[12];          // cy, sy, 1/m, Iinv[9]
    double ring[2][48];      // lever arms of four consecutive stages, [row 12][stage 4], prefetched a group ahead
    int flag[8];             // [0..3] per-thread PDAS flags; [4] row to drop, [5] row to add (single exchange), -1 = none
    unsigned hq[4];          // per-thread hash of the next working set
    unsigned hist[8];        // hashes of the last eight working sets (cycle detection)
};
// An odd number of 16-byte units: the eight robots of a warp then start in eight different bank groups, so a 16-byte
// (or 8-byte) access at the same offset of every robot -- the broadcast reads of the 6 x 6 work -- is one wavefront.
CMPC_HD size_t robot_bytes(int N) { return ((sizeof(Sh) + (size_t)3 * 4 * N + 15) & ~(size_t)15) | (size_t)16; }
CMPC_HD unsigned char* codes_of(Sh* sh) { return reinterpret_cast<unsigned char*>(sh) + sizeof(Sh); }

// per-thread state that lives across synchronisation points
struct TS {
    double P[36];            // block row a of P, P[i*12 + c] (registers between phase 3 and phase 1; parked in Sh::PK otherwise)
    double pv[3];            // block a of p ; forward: block a of x
    double qa[3];            // d_a
    double pg[3];            // backward: block a of P g
    double xr[3];            // reference entries of the stage (own rows)
    double ax[3];            // forward: block a of A x_k ; finish: co-state block
    double red[4];           // finish: partial reductions
    D2 gk[GAIN_D2];          // forward: gains of the current stage (loaded one stage ahead)
    int chg, nact;           // forward: changed foot-steps; rows of the next working set (this thread's foot)
    unsigned hsh;            // forward: running hash of this thread's part of the next working set
    double dworst, amost;    // forward: single-exchange candidates of this thread's foot (most negative multiplier, largest violation)
    int didx, aidx;
};
CMPC_HD int lt(int i, int j) { return ((i * (i + 1)) >> 1) + j; }      // lower-triangle index, i >= j

CMPC_HD double wr_rsqrt(double v) {
#if defined(__CUDA_ARCH__)
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(v));
    const double tt = y0 * y0;
    const double e = fma(-tt, v, 1.0);
    const double pl = fma(e, 0.375, 0.5);
    const double u = y0 * e;
    return fma(pl, u, y0);
#else
    return 1.0 / sqrt(v);
#endif
}

// ---- asynchronous copies of 8 bytes global -> shared (device: cp.async; host: plain copy)
CMPC_HD void cp8(double* dst_smem, const double* src) {
#if defined(__CUDA_ARCH__)
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
#else
    *dst_smem = *src;
#endif
}
CMPC_HD void cp16(double* dst_smem, const double* src) {
#if defined(__CUDA_ARCH__)
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
#else
    dst_smem[0] = src[0]; dst_smem[1] = src[1];
#endif
}
CMPC_HD void cp_commit_wait() {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
#endif
}
// 16-byte shared-memory load of two adjacent doubles (the compiler cannot prove the alignment itself)
CMPC_HD D2 ld2(const double* p) { return *reinterpret_cast<const D2*>(p); }

// foot projection for a working-set code: bit 0 fz pinned at fz_min, bits 1-2 fx (0 free, 1 = +mu fz, 2 = -mu fz),
// bits 3-4 fy likewise; SWING = not in stance
struct FootP { double dx, dy, sg, zx, zy, fh; };
CMPC_HD FootP foot_proj(unsigned char code, int leg, const Tab& tb, double mu, double fz_min) {
    const bool st = code != SWING;
    const int c = st ? (int)code : 0;
    const int az = c & 1, ax = (c >> 1) & 3, ay = (c >> 3) & 3;
    FootP f;
    f.zx = ax == 0 ? 0.0 : (ax == 1 ? mu : -mu);
    f.zy = ay == 0 ? 0.0 : (ay == 1 ? mu : -mu);
    f.dx = (st && ax == 0) ? tb.Rinv[3 * leg] : 0.0;
    f.dy = (st && ay == 0) ? tb.Rinv[3 * leg + 1] : 0.0;
    f.sg = (st && !az) ? tb.sig[leg * 4 + (ax != 0 ? 2 : 0) + (ay != 0 ? 1 : 0)] : 0.0;
    f.fh = (st && az) ? fz_min : 0.0;
    return f;
}

// W = Iinv [r]x   (row-major 3 x 3)
CMPC_HD void foot_W(const double* Ii, const double* r, double* W) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        W[i * 3 + 0] = Ii[i * 3 + 1] * r[2] - Ii[i * 3 + 2] * r[1];
        W[i * 3 + 1] = -Ii[i * 3 + 0] * r[2] + Ii[i * 3 + 2] * r[0];
        W[i * 3 + 2] = Ii[i * 3 + 0] * r[1] - Ii[i * 3 + 1] * r[0];
    }
}

// The 6 x 6 kernels below work on register arrays: every index has to be a compile-time constant, so the triangular
// loops are written column by column with the column as a template parameter (constant trip counts for the unroller).
template <int J>
CMPC_HD void chol6_col(double* G, double* di, double& pmin) {
    double s = G[lt(J, J)];
#pragma unroll
    for (int m = 0; m < J; ++m) s -= G[lt(J, m)] * G[lt(J, m)];
    pmin = fmin(pmin, s);
    const double r = wr_rsqrt(s);
    di[J] = r;
    G[lt(J, J)] = s * r;
#pragma unroll
    for (int i = J + 1; i < 6; ++i) {
        double v = G[lt(i, J)];
#pragma unroll
        for (int m = 0; m < J; ++m) v -= G[lt(i, m)] * G[lt(J, m)];
        G[lt(i, J)] = v * r;
    }
}
// in-place Cholesky of a 6 x 6 lower triangle; di = reciprocal diagonal.  Returns the smallest pivot.
CMPC_HD double chol6(double* G, double* di) {
    double pmin = 1e300;
    chol6_col<0>(G, di, pmin); chol6_col<1>(G, di, pmin); chol6_col<2>(G, di, pmin);
    chol6_col<3>(G, di, pmin); chol6_col<4>(G, di, pmin); chol6_col<5>(G, di, pmin);
    return pmin;
}
template <int J>
CMPC_HD void fsub6_row(const double* L, const double* di, double* y) {
    double s = y[J];
#pragma unroll
    for (int m = 0; m < J; ++m) s -= L[lt(J, m)] * y[m];
    y[J] = s * di[J];
}
// y <- L^-1 y
CMPC_HD void fsub6(const double* L, const double* di, double* y) {
    fsub6_row<0>(L, di, y); fsub6_row<1>(L, di, y); fsub6_row<2>(L, di, y);
    fsub6_row<3>(L, di, y); fsub6_row<4>(L, di, y); fsub6_row<5>(L, di, y);
}
template <int J>
CMPC_HD void bsub6_row(const double* L, const double* di, double* y) {
    double s = y[J];
#pragma unroll
    for (int m = J + 1; m < 6; ++m) s -= L[lt(m, J)] * y[m];
    y[J] = s * di[J];
}
// y <- L^-T y
CMPC_HD void bsub6(const double* L, const double* di, double* y) {
    bsub6_row<5>(L, di, y); bsub6_row<4>(L, di, y); bsub6_row<3>(L, di, y);
    bsub6_row<2>(L, di, y); bsub6_row<1>(L, di, y); bsub6_row<0>(L, di, y);
}
template <int J>
CMPC_HD void lmul6_row(const double* L, const double* w, double* out) {
    double s = 0.0;
#pragma unroll
    for (int m = 0; m <= J; ++m) s += L[lt(J, m)] * w[m];
    out[J] = s;
}
// out = L w
CMPC_HD void lmul6(const double* L, const double* w, double* out) {
    lmul6_row<0>(L, w, out); lmul6_row<1>(L, w, out); lmul6_row<2>(L, w, out);
    lmul6_row<3>(L, w, out); lmul6_row<4>(L, w, out); lmul6_row<5>(L, w, out);
}
// column J of Nn = I + L' Lam L  (lower triangle)
template <int J>
CMPC_HD void nn_col(const double* lam, const double* L, double* Ln) {
    double tj[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        double s = 0.0;
#pragma unroll
        for (int m = J; m < 6; ++m) s += lam[r >= m ? lt(r, m) : lt(m, r)] * L[lt(m, J)];
        tj[r] = s;
    }
#pragma unroll
    for (int i = J; i < 6; ++i) {
        double s = (i == J) ? 1.0 : 0.0;
#pragma unroll
        for (int m = 5; m >= 0; --m) if (m >= i) s += L[lt(m >= i ? m : i, i)] * tj[m];
        Ln[lt(i, J)] = s;
    }
}

struct Env {                 // what one quad needs to know about its robot (uniform inside the quad)
    const Params* p;
    const Tab* tb;
    const Bat* bt;
    int b;                   // robot
    D2* gains;               // this thread's slots: gains[(k * GAIN_D2 + e) * gstride]
    size_t gstride;
    double dt, h;
};
CMPC_HD const double* in_xref(const Env& e) { return e.bt->x_ref + (size_t)e.b * 12 * e.bt->N; }
CMPC_HD const double* in_rfoot(const Env& e) { return e.bt->r_foot + (size_t)e.b * 12 * e.bt->N; }
CMPC_HD double* out_u(const Env& e) { return e.bt->u + (size_t)e.b * 12 * e.bt->N; }
CMPC_HD double* out_y(const Env& e) { return e.bt->y + (size_t)e.b * 28 * e.bt->N; }
CMPC_HD double* out_X(const Env& e) { return e.bt->X + (size_t)e.b * 12 * e.bt->N; }
CMPC_HD double* out_stats(const Env& e) { return e.bt->stats + (size_t)e.b * NSTAT; }
CMPC_HD int stance_at(const Env& e, int leg, int k) {
    return mask_bit(e.bt->mask ? e.bt->mask + (size_t)e.b * e.bt->W : nullptr, e.bt->N, leg, k);
}

// block a of  Bbar w
CMPC_HD void bbar_rows(int a, const double* cst, double dt, double h, const double* w, double* out) {
    const double cy = cst[0], sy = cst[1], minv = cst[2];
    const double cf = (a == 0) ? h * minv : (a == 1 ? h : (a == 2 ? dt * minv : dt));
    const bool tq = (a & 1) != 0;         // selects, not pointer arithmetic: w is a register array
    const double s0 = tq ? w[3] : w[0], s1 = tq ? w[4] : w[1], s2 = tq ? w[5] : w[2];
    const double c = (a == 1) ? cy : 1.0, sn = (a == 1) ? sy : 0.0;
    out[0] = cf * (c * s0 + sn * s1);
    out[1] = cf * (-sn * s0 + c * s1);
    out[2] = cf * s2;
}
// out (6) = Bbar' v  for a 12-vector v in shared memory
CMPC_HD void bbar_t(const double* cst, double dt, double h, const double* v, double* out) {
    const double cy = cst[0], sy = cst[1], minv = cst[2];
#pragma unroll
    for (int c = 0; c < 3; ++c) out[c] = minv * (h * v[c] + dt * v[6 + c]);
    out[3] = h * (cy * v[3] - sy * v[4]) + dt * v[9];
    out[4] = h * (sy * v[3] + cy * v[4]) + dt * v[10];
    out[5] = h * v[5] + dt * v[11];
}

// prefetch the lever arms of the four stages of group g (stages 4g .. 4g+3; N is a multiple of 4): thread q brings the
// three rows of foot q, 32 bytes = one sector each
CMPC_HD void ring_issue(int q, Sh* sh, const Env& e, int g) {
    const int N = e.bt->N;
    const double* rf = in_rfoot(e) + 4 * g;
    double* dst = sh->ring[g & 1];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        cp16(dst + (3 * q + c) * 4, rf + (size_t)(3 * q + c) * N);
        cp16(dst + (3 * q + c) * 4 + 2, rf + (size_t)(3 * q + c) * N + 2);
    }
}
CMPC_HD const double* ring_at(const Sh* sh, int k) { return sh->ring[(k >> 2) & 1] + (k & 3); }      // row r at [4 r]

CMPC_HD void robot_consts(const Env& e, double* cst) {
    DynCommon dc;
    dyn_common(dc, in_xref(e), e.bt->N, e.bt->I_world + (size_t)e.b * 9, e.bt->mass[e.b], e.bt->dt);
    cst[0] = dc.cy; cst[1] = dc.sy; cst[2] = dc.minv;
    for (int i = 0; i < 9; ++i) cst[3 + i] = dc.Iinv[i];
}

// ------------------------------------------------------------------------------------------------------------------
// robot set-up: constants, initial working set (from the contact table, or from the previous duals when warm)
// ------------------------------------------------------------------------------------------------------------------
CMPC_HD int init_robot(int qlane, bool valid, TS* ts, Sh* sh, const Env& e, int warm) {
    const int N = e.bt->N;
    unsigned char* codes = codes_of(sh);
    int nst = 0;
    WR_Q_BEGIN
    (void)t; (void)q;
    if (valid) {
        if (e.bt->cst) {             // formed ahead of the sweeps, all robots at once (no sincos / serial loads in this kernel)
#pragma unroll
            for (int c = 0; c < 3; ++c) sh->cst[3 * q + c] = e.bt->cst[(size_t)e.b * 12 + 3 * q + c];
        }
#if !defined(__CUDA_ARCH__)
        else if (q == 0) robot_consts(e, sh->cst);      // host emulation: formed here
#endif
        const double* yo = out_y(e);
        for (int k = 0; k < N; ++k) {
            const int st = stance_at(e, q, k);
            unsigned char c = SWING;
            if (st) {
                c = 0;
                if (warm) {
                    // warm = 1: the previous working set as it is (the reference re-uses its raw previous solution,
                    // centroidal_mpc.py:92-95,108-110); warm = 2: shifted by one stage (the cycle advanced by about one dt)
                    const int ks = (warm == 2 && k + 1 < N) ? k + 1 : k;
                    const double* yf = yo + 12 * N + 16 * ks + 4 * q;
                    const int az = yo[12 * ks + 3 * q + 2] < 0.0;
                    const int ax = yf[0] > 0.0 ? 1 : (yf[1] > 0.0 ? 2 : 0);
                    const int ay = yf[2] > 0.0 ? 1 : (yf[3] > 0.0 ? 2 : 0);
                    c = (unsigned char)(az | (ax << 1) | (ay << 3));
                }
            }
            codes[4 * k + q] = c;
            codes[4 * N + 4 * k + q] = 254;          // "previous" sets that match nothing
            codes[8 * N + 4 * k + q] = 254;
        }
    }
    WR_Q_END
    WR_SYNC();
    if (valid) for (int i = 0; i < 4 * N; ++i) nst += codes[i] != SWING;      // uniform in the quad
    return nst;
}

// ------------------------------------------------------------------------------------------------------------------
// backward sweep for the working set cur.  Returns the smallest pivot seen (<= 0: not positive definite).
//
// What bounds this kernel is the load/store data pipe of the SM (shared-memory wavefronts), not the FP64 pipe, so a
// stage moves as little as possible through shared memory: the per-foot terms of Lam_k are formed by every thread
// itself (no exchange), S = P Bbar is read once for the 6 x 6 work and once for the update of all three rows, P is
// parked once around the 6 x 6 work, and after the update only the rows the partner threads need are exchanged.
// ------------------------------------------------------------------------------------------------------------------
CMPC_HD double backward_sweep(int qlane, TS* ts, Sh* sh, const Env& e, const unsigned char* cur) {
    const int N = e.bt->N;
    const Params& p = *e.p;
    const Tab& tb = *e.tb;
    const double dt = e.dt, h = e.h;
    double pmin = 1e300;
    // terminal cost and the first prefetch
    WR_Q_BEGIN
    const double* xr = in_xref(e);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int c = 0; c < 12; ++c) t.P[i * 12 + c] = (c == 3 * q + i) ? tb.Q[3 * q + i] : 0.0;
        t.pv[i] = -tb.Q[3 * q + i] * xr[(size_t)(3 * q + i) * N + (N - 1)];
    }
    ring_issue(q, sh, e, (N - 1) >> 2);
    WR_Q_END
    for (int k = N - 1; k >= 0; --k) {
        // ---- phase 1: own rows of S = P Bbar, P g, p_a; P is parked
        WR_Q_BEGIN
        if ((k & 3) == 3 || k == N - 1) cp_commit_wait();             // the group of this stage has arrived (own copies)
        if (k > 0) {
            const double* xr = in_xref(e);
#pragma unroll
            for (int i = 0; i < 3; ++i) t.xr[i] = xr[(size_t)(3 * q + i) * N + (k - 1)];
        }
        const double cy = sh->cst[0], sy = sh->cst[1], minv = sh->cst[2];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double* Pi = t.P + i * 12;
            double* o = sh->E1 + q * 18 + i * 6;
            D2 v0, v1, v2;
            v0.x = minv * (h * Pi[0] + dt * Pi[6]);
            v0.y = minv * (h * Pi[1] + dt * Pi[7]);
            v1.x = minv * (h * Pi[2] + dt * Pi[8]);
            v1.y = h * (Pi[3] * cy - Pi[4] * sy) + dt * Pi[9];
            v2.x = h * (Pi[3] * sy + Pi[4] * cy) + dt * Pi[10];
            v2.y = h * Pi[5] + dt * Pi[11];
            *reinterpret_cast<D2*>(o) = v0; *reinterpret_cast<D2*>(o + 2) = v1; *reinterpret_cast<D2*>(o + 4) = v2;
            t.pg[i] = -9.81 * (h * Pi[2] + dt * Pi[8]);
            sh->E3[3 * q + i] = t.pv[i];
            if (k > 0) {
                double* pk = sh->PK + q * 36 + i * 12;
#pragma unroll
                for (int c = 0; c < 12; c += 2) { D2 v; v.x = Pi[c]; v.y = Pi[c + 1]; *reinterpret_cast<D2*>(pk + c) = v; }
            }
        }
        WR_Q_END
        WR_SYNC();
        // ---- phase 2: Lam_k / what_k from the four feet, 6 x 6 factorizations (every thread), vector part, gains, own rows of
        //      D = P - S Phi S' and of D A
        WR_Q_BEGIN
        if ((k & 3) == 3 && k >= 4) ring_issue(q, sh, e, (k >> 2) - 1);      // next group (every thread of the quad has passed the wait above)
        const double cy = sh->cst[0], sy = sh->cst[1], minv = sh->cst[2];
        const double* E1 = sh->E1;
        const double* own = E1 + q * 18;
        double lam[21], wh[6], bq[6];
        {
#pragma unroll
            for (int i = 0; i < 21; ++i) lam[i] = 0.0;
#pragma unroll
            for (int i = 0; i < 6; ++i) wh[i] = 0.0;
            const double* rg = ring_at(sh, k);
#pragma unroll 1          // rolled on purpose: the stage body is instruction-fetch bound (17 % of the warp samples wait for instructions)
            for (int j = 0; j < 4; ++j) {
                const FootP f = foot_proj(cur[4 * k + j], j, tb, p.mu, p.fz_min);
                const double r3[3] = {rg[(3 * j) * 4], rg[(3 * j + 1) * 4], rg[(3 * j + 2) * 4]};
                double W[9], Wz[3];
                foot_W(sh->cst + 3, r3, W);
#pragma unroll
                for (int m = 0; m < 3; ++m) Wz[m] = W[m * 3] * f.zx + W[m * 3 + 1] * f.zy + W[m * 3 + 2];
                const double sx = f.sg * f.zx, sy_ = f.sg * f.zy;
                lam[lt(0, 0)] += f.dx + sx * f.zx;
                lam[lt(1, 0)] += sx * f.zy;
                lam[lt(1, 1)] += f.dy + sy_ * f.zy;
                lam[lt(2, 0)] += sx;
                lam[lt(2, 1)] += sy_;
                lam[lt(2, 2)] += f.sg;
#pragma unroll
                for (int m = 0; m < 3; ++m) {
                    const double a0 = f.dx * W[m * 3], a1 = f.dy * W[m * 3 + 1], a2 = f.sg * Wz[m];
                    lam[lt(3 + m, 0)] += a0 + sx * Wz[m];
                    lam[lt(3 + m, 1)] += a1 + sy_ * Wz[m];
                    lam[lt(3 + m, 2)] += a2;
#pragma unroll
                    for (int n = 0; n <= m; ++n) lam[lt(3 + m, 3 + n)] += a0 * W[n * 3] + a1 * W[n * 3 + 1] + a2 * Wz[n];
                }
                wh[0] += f.fh * f.zx; wh[1] += f.fh * f.zy; wh[2] += f.fh;
#pragma unroll
                for (int m = 0; m < 3; ++m) wh[3 + m] += f.fh * Wz[m];
            }
        }
        {   // ghat = g + Bbar what;  bq = Bbar'q = S' ghat + Bbar' p
            double gh[12];
#pragma unroll
            for (int a = 0; a < 4; ++a) bbar_rows(a, sh->cst, dt, h, wh, gh + 3 * a);
            gh[2] += -9.81 * h;
            gh[8] += -9.81 * dt;
            bbar_t(sh->cst, dt, h, sh->E3, bq);
#pragma unroll
            for (int r = 0; r < 12; ++r) {
                const D2 a = ld2(E1 + r * 6), b = ld2(E1 + r * 6 + 2), c = ld2(E1 + r * 6 + 4);
                bq[0] += a.x * gh[r]; bq[1] += a.y * gh[r]; bq[2] += b.x * gh[r];
                bq[3] += b.y * gh[r]; bq[4] += c.x * gh[r]; bq[5] += c.y * gh[r];
            }
        }
        double L[21], di[6];
        {   // Gbar = Bbar' S, lower triangle
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c <= r; ++c) L[lt(r, c)] = minv * (h * E1[0 * 18 + r * 6 + c] + dt * E1[2 * 18 + r * 6 + c]);
#pragma unroll
            for (int c = 0; c < 6; c += 2) {
                const D2 x0 = ld2(E1 + 18 + c), x1 = ld2(E1 + 18 + 6 + c), x2 = ld2(E1 + 18 + 12 + c);
                const D2 w0 = ld2(E1 + 54 + c), w1 = ld2(E1 + 54 + 6 + c), w2 = ld2(E1 + 54 + 12 + c);
                const double r0a = h * (cy * x0.x - sy * x1.x) + dt * w0.x, r0b = h * (cy * x0.y - sy * x1.y) + dt * w0.y;
                const double r1a = h * (sy * x0.x + cy * x1.x) + dt * w1.x, r1b = h * (sy * x0.y + cy * x1.y) + dt * w1.y;
                const double r2a = h * x2.x + dt * w2.x, r2b = h * x2.y + dt * w2.y;
                if (c <= 3) L[lt(3, c)] = r0a;
                if (c + 1 <= 3) L[lt(3, c + 1 <= 3 ? c + 1 : 0)] = r0b;
                if (c <= 4) L[lt(4, c)] = r1a;
                if (c + 1 <= 4) L[lt(4, c + 1 <= 4 ? c + 1 : 0)] = r1b;
                L[lt(5, c)] = r2a;
                L[lt(5, c + 1)] = r2b;
            }
        }
        pmin = fmin(pmin, chol6(L, di));
        double Ln[21], dn[6];
        nn_col<0>(lam, L, Ln); nn_col<1>(lam, L, Ln); nn_col<2>(lam, L, Ln);        // Nn = I + L' Lam L
        nn_col<3>(lam, L, Ln); nn_col<4>(lam, L, Ln); nn_col<5>(lam, L, Ln);
        chol6(Ln, dn);
        D2* g = e.gains + WR_GQ + (size_t)k * GAIN_D2 * e.gstride;
        // vector part: kbar = Gam bq,  phiq = Phi bq;  dvec = what - phiq enters d_a = p_a + (P g)_a + S_a (what - Phi bq)
        double dvec[6];
        {
            double w[6], kb[6], phiq[6];
            fsub6(L, di, bq);
#pragma unroll
            for (int c = 0; c < 6; ++c) w[c] = bq[c];
            fsub6(Ln, dn, w);
            bsub6(Ln, dn, w);
            lmul6(L, w, kb);
#pragma unroll
            for (int c = 0; c < 6; ++c) phiq[c] = bq[c] - w[c];
            bsub6(L, di, phiq);
#pragma unroll
            for (int c = 0; c < 6; ++c) dvec[c] = wh[c] - phiq[c];
            D2 kk;
            kk.x = q == 0 ? kb[0] : (q == 1 ? kb[2] : (q == 2 ? kb[4] : 0.0));
            kk.y = q == 0 ? kb[1] : (q == 1 ? kb[3] : (q == 2 ? kb[5] : 0.0));
            g[9 * e.gstride] = kk;
        }
        // gains and z_i = Phi S_i', one row at a time
        double z[18];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            double y[6], w[6], kt[6];
            {
                const D2 a = ld2(own + i * 6), b = ld2(own + i * 6 + 2), c = ld2(own + i * 6 + 4);
                y[0] = a.x; y[1] = a.y; y[2] = b.x; y[3] = b.y; y[4] = c.x; y[5] = c.y;
            }
            if (k > 0) {      // d_a = p_a + (P g)_a + S_a (what - Phi bq)
                double d = t.pv[i] + t.pg[i];
#pragma unroll
                for (int c = 0; c < 6; ++c) d += y[c] * dvec[c];
                t.qa[i] = d;
                if (q < 2) sh->D4[q * 4 + i] = d;
            }
            fsub6(L, di, y);
#pragma unroll
            for (int c = 0; c < 6; ++c) w[c] = y[c];
            fsub6(Ln, dn, w);
            bsub6(Ln, dn, w);
            lmul6(L, w, kt);
#pragma unroll
            for (int c = 0; c < 3; ++c) { D2 v; v.x = kt[2 * c]; v.y = kt[2 * c + 1]; g[(3 * i + c) * e.gstride] = v; }
            if (k > 0) {
#pragma unroll
                for (int c = 0; c < 6; ++c) z[i * 6 + c] = y[c] - w[c];
                bsub6(L, di, z + i * 6);
            }
        }
        if (k > 0) {
            // the three rows of D = P - Z S' together (S is read once), then D A; the rows of p, rpy go to their partners
            double* pk = sh->PK + q * 36;
#pragma unroll
            for (int c = 0; c < 36; c += 2) { const D2 v = ld2(pk + c); t.P[c] = v.x; t.P[c + 1] = v.y; }
#pragma unroll
            for (int b = 0; b < 12; ++b) {
                const D2 a = ld2(E1 + b * 6), bb = ld2(E1 + b * 6 + 2), c = ld2(E1 + b * 6 + 4);
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const double* zi = z + i * 6;
                    t.P[i * 12 + b] -= zi[0] * a.x + zi[1] * a.y + zi[2] * bb.x + zi[3] * bb.y + zi[4] * c.x + zi[5] * c.y;
                }
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                double* Pi = t.P + i * 12;
                Pi[6] += dt * Pi[0]; Pi[7] += dt * Pi[1]; Pi[8] += dt * Pi[2];
                Pi[9] += dt * (cy * Pi[3] - sy * Pi[4]);
                Pi[10] += dt * (sy * Pi[3] + cy * Pi[4]);
                Pi[11] += dt * Pi[5];
            }
            if (q < 2) {
#pragma unroll
                for (int c = 0; c < 36; c += 2) { D2 v; v.x = t.P[c]; v.y = t.P[c + 1]; *reinterpret_cast<D2*>(pk + c) = v; }
            }
        }
        WR_Q_END
        if (k == 0) break;
        WR_SYNC();
        // ---- phase 3: P <- Q + A'(D A),  p <- -Q xref + A' d
        WR_Q_BEGIN
        const double q0 = tb.Q[3 * q], q1 = tb.Q[3 * q + 1], q2 = tb.Q[3 * q + 2];
        const double dtc = q >= 2 ? dt : 0.0;
        const double rc = q == 3 ? sh->cst[0] : 1.0, rs = q == 3 ? sh->cst[1] : 0.0;
        if (q >= 2) {
            const double* X = sh->PK + (q & 1) * 36;          // rows of the partner block (p for v, rpy for omega)
#pragma unroll
            for (int c = 0; c < 12; c += 2) {
                const D2 x0 = ld2(X + c), x1 = ld2(X + 12 + c), x2 = ld2(X + 24 + c);
                t.P[c] += dt * (rc * x0.x - rs * x1.x);
                t.P[c + 1] += dt * (rc * x0.y - rs * x1.y);
                t.P[12 + c] += dt * (rs * x0.x + rc * x1.x);
                t.P[12 + c + 1] += dt * (rs * x0.y + rc * x1.y);
                t.P[24 + c] += dt * x2.x;
                t.P[24 + c + 1] += dt * x2.y;
            }
        }
#pragma unroll
        for (int c = 0; c < 12; ++c) {       // selects: P is a register array, its indices must be compile-time constants
            t.P[c] += (c == 3 * q) ? q0 : 0.0;
            t.P[12 + c] += (c == 3 * q + 1) ? q1 : 0.0;
            t.P[24 + c] += (c == 3 * q + 2) ? q2 : 0.0;
        }
        const double* D = sh->D4 + (q & 1) * 4;
        const double d0 = D[0], d1 = D[1], d2 = D[2];
        t.pv[0] = t.qa[0] + dtc * (rc * d0 - rs * d1) - q0 * t.xr[0];
        t.pv[1] = t.qa[1] + dtc * (rs * d0 + rc * d1) - q1 * t.xr[1];
        t.pv[2] = t.qa[2] + dtc * d2 - q2 * t.xr[2];
        WR_Q_END
        WR_SYNC();            // the partners' rows and S are rewritten in the next stage
    }
    WR_SYNC();
    return pmin;
}

// ------------------------------------------------------------------------------------------------------------------
// forward sweep: states, forces, multipliers, next working set.  Writes forces / stance duals / states to the output
// arrays when `valid`.  Returns bit 0: the working set changed, bit 1: the new set is one of the last eight (a cycle;
// `it` = sweeps done before this one = valid entries of the history), bit 2: there is a single-exchange candidate,
// bits 8-15: the number of foot-steps whose working set changed, bits 16 and up: rows of the next working set.
// ------------------------------------------------------------------------------------------------------------------
// With `damp` the block update is one-sided on every other foot-step: rows whose multiplier is negative leave the working
// set only where k + foot + it is even (violated rows always join).  Dropping half of the candidates per sweep is what
// keeps the updates of heavily disturbed robots (40+ active rows) from oscillating; see policy_step.
CMPC_HD int forward_sweep(int qlane, bool valid, TS* ts, Sh* sh, const Env& e, const unsigned char* cur,
                          unsigned char* next, int it, bool damp) {
    const int N = e.bt->N;
    const Params& p = *e.p;
    const Tab& tb = *e.tb;
    const double dt = e.dt, h = e.h;
    const double tol = 1e-10;
    double* xs = sh->E1;
    double* mup = sh->E1 + 12;
    double* wp = sh->E1 + 36;
    WR_Q_BEGIN
    const double* x0 = e.bt->x0 + (size_t)e.b * 12;
#pragma unroll
    for (int i = 0; i < 3; ++i) t.pv[i] = x0[3 * q + i];
    t.chg = 0; t.nact = 0; t.hsh = 2166136261u;
    t.dworst = -tol; t.amost = 1e-9; t.didx = -1; t.aidx = -1;
    ring_issue(q, sh, e, 0);
    const D2* g = e.gains + WR_GQ;
#pragma unroll
    for (int i = 0; i < GAIN_D2; ++i) t.gk[i] = g[i * e.gstride];
    WR_Q_END
    for (int k = 0; k < N; ++k) {
        // ---- F1: publish x
        WR_Q_BEGIN
        if ((k & 3) == 0) cp_commit_wait();
#pragma unroll
        for (int i = 0; i < 3; ++i) xs[3 * q + i] = t.pv[i];
        WR_Q_END
        WR_SYNC();
        // ---- F2: A x (own block), partial wrench co-state; gains of the next stage
        WR_Q_BEGIN
        if ((k & 3) == 0 && k + 4 < N) ring_issue(q, sh, e, (k >> 2) + 1);
        const double dtf = q < 2 ? dt : 0.0;
        const double rc = q == 1 ? sh->cst[0] : 1.0, rs = q == 1 ? sh->cst[1] : 0.0;
        const double* X = xs + 3 * (q | 2);
        t.ax[0] = t.pv[0] + dtf * (rc * X[0] + rs * X[1]);
        t.ax[1] = t.pv[1] + dtf * (-rs * X[0] + rc * X[1]);
        t.ax[2] = t.pv[2] + dtf * X[2];
        const D2 kk = t.gk[9];
        double mp[6];
#pragma unroll
        for (int c = 0; c < 3; ++c) { mp[2 * c] = (q == c) ? kk.x : 0.0; mp[2 * c + 1] = (q == c) ? kk.y : 0.0; }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const D2 v = t.gk[3 * i + c];
                mp[2 * c] += v.x * t.ax[i];
                mp[2 * c + 1] += v.y * t.ax[i];
            }
#pragma unroll
        for (int c = 0; c < 6; ++c) mup[q * 6 + c] = mp[c];
        if (k + 1 < N) {
            const D2* g = e.gains + WR_GQ + (size_t)(k + 1) * GAIN_D2 * e.gstride;
#pragma unroll
            for (int i = 0; i < GAIN_D2; ++i) t.gk[i] = g[i * e.gstride];
        }
        WR_Q_END
        WR_SYNC();
        // ---- F3: foot q -- force, multipliers, next code, wrench contribution
        WR_Q_BEGIN
        double mu6[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) mu6[c] = mup[c] + mup[6 + c] + mup[12 + c] + mup[18 + c];
        const unsigned char code = cur[4 * k + q];
        const FootP f = foot_proj(code, q, tb, p.mu, p.fz_min);
        double W[9];
        {
            const double* rg = ring_at(sh, k) + 12 * q;
            const double r3[3] = {rg[0], rg[4], rg[8]};
            foot_W(sh->cst + 3, r3, W);
        }
        double tv[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) tv[c] = mu6[c] + W[c] * mu6[3] + W[3 + c] * mu6[4] + W[6 + c] * mu6[5];
        const double zt = f.zx * tv[0] + f.zy * tv[1] + tv[2];
        double fo[3];
        fo[0] = f.fh * f.zx - f.dx * tv[0] - f.sg * f.zx * zt;
        fo[1] = f.fh * f.zy - f.dy * tv[1] - f.sg * f.zy * zt;
        fo[2] = f.fh - f.sg * zt;
        unsigned char nc = SWING;
        double l5[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
        if (code != SWING) {
            const int az = code & 1, ax = (code >> 1) & 3, ay = (code >> 3) & 3;
            const double d0 = tb.R[3 * q] * fo[0] + tv[0], d1 = tb.R[3 * q + 1] * fo[1] + tv[1], d2 = tb.R[3 * q + 2] * fo[2] + tv[2];
            double lx = 0.0, ly = 0.0;
            if (ax == 1) { lx = -2.0 * d0; l5[1] = lx; } else if (ax == 2) { lx = 2.0 * d0; l5[2] = lx; }
            if (ay == 1) { ly = -2.0 * d1; l5[3] = ly; } else if (ay == 2) { ly = 2.0 * d1; l5[4] = ly; }
            if (az) l5[0] = 2.0 * d2 - p.mu * (lx + ly);
            const double s0 = l5[0] + (p.fz_min - fo[2]);
            const double s1 = l5[1] + (fo[0] - p.mu * fo[2]), s2 = l5[2] + (-fo[0] - p.mu * fo[2]);
            const double s3 = l5[3] + (fo[1] - p.mu * fo[2]), s4 = l5[4] + (-fo[1] - p.mu * fo[2]);
            const int nz = s0 > tol;
            const int nx = (s1 > tol && s1 >= s2) ? 1 : ((s2 > tol && s2 > s1) ? 2 : 0);
            const int ny = (s3 > tol && s3 >= s4) ? 1 : ((s4 > tol && s4 > s3) ? 2 : 0);
            const bool keep = damp && (((k + q + it) & 1) != 0);
            const int nz2 = (keep && az) ? 1 : nz, nx2 = (keep && ax != 0) ? ax : nx, ny2 = (keep && ay != 0) ? ay : ny;
            nc = (unsigned char)(nz2 | (nx2 << 1) | (ny2 << 3));
            // single-exchange candidates (solve_active_set_fast, second phase): the most negative multiplier of the
            // working set, the most violated row outside it (not the face opposite to an active one)
            const double vv[5] = {p.fz_min - fo[2], fo[0] - p.mu * fo[2], -fo[0] - p.mu * fo[2], fo[1] - p.mu * fo[2], -fo[1] - p.mu * fo[2]};
            const bool on[5] = {az != 0, ax == 1, ax == 2, ay == 1, ay == 2};
            const bool free5[5] = {true, ax == 0, ax == 0, ay == 0, ay == 0};
            const int base = 5 * (4 * k + q);
#pragma unroll
            for (int r5 = 0; r5 < 5; ++r5) {
                if (on[r5]) { if (l5[r5] < t.dworst) { t.dworst = l5[r5]; t.didx = base + r5; } }
                else if (free5[r5] && vv[r5] > t.amost) { t.amost = vv[r5]; t.aidx = base + r5; }
            }
        }
        next[4 * k + q] = nc;
        t.chg += (nc != code);
        t.hsh = (t.hsh ^ nc) * 16777619u;
        if (nc != SWING) t.nact += (nc & 1) + ((nc >> 1) & 3 ? 1 : 0) + ((nc >> 3) & 3 ? 1 : 0);
        if (valid) {
            double* uo = out_u(e) + 12 * k + 3 * q;
            double* yo = out_y(e);
            uo[0] = fo[0]; uo[1] = fo[1]; uo[2] = fo[2];
            if (code != SWING) yo[12 * k + 3 * q + 2] = -l5[0];
            double* yf = yo + 12 * N + 16 * k + 4 * q;
            { D2 a, b2; a.x = l5[1]; a.y = l5[2]; b2.x = l5[3]; b2.y = l5[4];          // 32-byte aligned: two 16-byte stores
              reinterpret_cast<D2*>(yf)[0] = a; reinterpret_cast<D2*>(yf)[1] = b2; }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            wp[q * 6 + c] = fo[c];
            wp[q * 6 + 3 + c] = W[c * 3] * fo[0] + W[c * 3 + 1] * fo[1] + W[c * 3 + 2] * fo[2];
        }
        WR_Q_END
        WR_SYNC();
        // ---- F4: x_{k+1}
        WR_Q_BEGIN
        double w6[6], bw[3];
#pragma unroll
        for (int c = 0; c < 6; ++c) w6[c] = wp[c] + wp[6 + c] + wp[12 + c] + wp[18 + c];
        bbar_rows(q, sh->cst, dt, h, w6, bw);
        t.pv[0] = t.ax[0] + bw[0];
        t.pv[1] = t.ax[1] + bw[1];
        t.pv[2] = t.ax[2] + bw[2] + (q == 0 ? -9.81 * h : (q == 2 ? -9.81 * dt : 0.0));
        if (valid) {
            double* Xo = out_X(e) + 12 * k + 3 * q;
            Xo[0] = t.pv[0]; Xo[1] = t.pv[1]; Xo[2] = t.pv[2];
        }
        WR_Q_END
    }
    WR_SYNC();
    int res = 0;
    double* cand = sh->E1 + 60;           // 4 x (value of the drop candidate, value of the add candidate)
    int* candi = reinterpret_cast<int*>(sh->E1 + 68);
    WR_Q_BEGIN
    sh->flag[q] = t.chg | (t.nact << 12);
    sh->hq[q] = t.hsh;
    cand[2 * q] = t.dworst; cand[2 * q + 1] = t.amost;
    candi[2 * q] = t.didx; candi[2 * q + 1] = t.aidx;
    WR_Q_END
    WR_SYNC();
    const int fsum = sh->flag[0] + sh->flag[1] + sh->flag[2] + sh->flag[3];
    const int nchg = fsum & 0xfff, nact = fsum >> 12;
    res = (nchg > 0 ? 1 : 0) | ((nchg > 255 ? 255 : nchg) << 8) | (nact << 16);
    // (a damped update depends on the parity of the sweep, so the same set at the other parity is not a repetition)
    const unsigned H = sh->hq[0] * 0x9E3779B1u + sh->hq[1] * 0x85EBCA77u + sh->hq[2] * 0xC2B2AE3Du + sh->hq[3] * 0x27D4EB2Fu +
                       ((damp && (it & 1)) ? 0x165667B1u : 0u);
    {
        const int nh = it < 8 ? it : 8;
        bool rep = false;
        for (int j = 0; j < nh; ++j) rep |= (sh->hist[j] == H);
        if (rep) res |= 2;
    }
    {   // the same reduction in every thread: worst drop candidate, else best add candidate
        double dw = 0.0, am = 0.0;
        int di_ = -1, ai_ = -1;
        for (int j = 0; j < 4; ++j) {
            if (candi[2 * j] >= 0 && (di_ < 0 || cand[2 * j] < dw)) { dw = cand[2 * j]; di_ = candi[2 * j]; }
            if (candi[2 * j + 1] >= 0 && (ai_ < 0 || cand[2 * j + 1] > am)) { am = cand[2 * j + 1]; ai_ = candi[2 * j + 1]; }
        }
        if (di_ >= 0 || ai_ >= 0) res |= 4;
        WR_SYNC();
        WR_Q_BEGIN
        (void)t;
        if (q == 0) { sh->flag[4] = di_; sh->flag[5] = di_ >= 0 ? -1 : ai_; sh->hist[it & 7] = H; }
        WR_Q_END
    }
    WR_SYNC();
    return res;
}

// Single exchange: next = cur with one row dropped (most negative multiplier) or, if there is none, one row added (most
// violated), as left in sh->flag[4..5] by the forward sweep.
CMPC_HD void single_step(int qlane, bool doit, TS* ts, Sh* sh, const unsigned char* cur, unsigned char* next, int N) {
    WR_Q_BEGIN
    (void)t;
    if (doit) for (int i = q; i < 4 * N; i += 4) next[i] = cur[i];
    WR_Q_END
    WR_SYNC();
    WR_Q_BEGIN
    (void)t;
    if (doit && q == 0) {
        const int drop = sh->flag[4], add = sh->flag[5];
        const int idx = drop >= 0 ? drop : add;
        if (idx >= 0) {
            const int pos = idx / 5, r5 = idx - 5 * pos;
            int c = next[pos];
            const int on = drop >= 0 ? 0 : 1;
            if (r5 == 0) c = on ? (c | 1) : (c & ~1);
            else if (r5 <= 2) c = (c & ~(3 << 1)) | ((on ? r5 : 0) << 1);
            else c = (c & ~(3 << 3)) | ((on ? r5 - 2 : 0) << 3);
            next[pos] = (unsigned char)c;
        }
    }
    WR_Q_END
    WR_SYNC();
}

// ------------------------------------------------------------------------------------------------------------------
// Finish of a robot whose working set has settled: co-states, stationarity and feasibility from first principles -- X,
// u, y as written by the forward sweep, independent of the factorizations -- and the remaining outputs in the
// reference's layouts.  The inputs of a stage are prefetched two slots deep into the storage of PK.  Returns 1 if the
// certificate holds (status / statistics written), else 0.
// ------------------------------------------------------------------------------------------------------------------
CMPC_HD void fin_issue(int q, bool valid, Sh* sh, int slot, const Env& e, int k) {
    if (!valid) return;
    const int N = e.bt->N;
    double* d = sh->PK + slot * 72 + q * 18;
    const double* xk = out_X(e) + 12 * k + 3 * q;
    const double* xr = in_xref(e);
    const double* rf = in_rfoot(e);
    const double* f = out_u(e) + 12 * k + 3 * q;
    const double* yo = out_y(e);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        cp8(d + c, xk + c);
        cp8(d + 3 + c, xr + (size_t)(3 * q + c) * N + k);
        cp8(d + 6 + c, rf + (size_t)(3 * q + c) * N + k);
        cp8(d + 9 + c, f + c);
    }
    cp8(d + 12, yo + 12 * k + 3 * q + 2);
#pragma unroll
    for (int c = 0; c < 4; ++c) cp8(d + 13 + c, yo + 12 * N + 16 * k + 4 * q + c);
}

CMPC_HD int finish_robot(int qlane, bool valid, TS* ts, Sh* sh, const Env& e, const unsigned char* cur, int warm, int nst, int sweeps) {
    const int N = e.bt->N;
    const Params& p = *e.p;
    const Tab& tb = *e.tb;
    const double dt = e.dt, h = e.h;
    double* nus = sh->E3;
    WR_Q_BEGIN
    t.red[0] = 0.0; t.red[1] = 0.0; t.red[2] = 0.0; t.red[3] = 0.0;       // rd, rp, objective part, active rows
    t.ax[0] = 0.0; t.ax[1] = 0.0; t.ax[2] = 0.0;                          // co-state block
    nus[3 * q] = 0.0; nus[3 * q + 1] = 0.0; nus[3 * q + 2] = 0.0;
    fin_issue(q, valid, sh, (N - 1) & 1, e, N - 1);
    WR_Q_END
    for (int k = N - 1; k >= 0; --k) {
        WR_Q_BEGIN
        cp_commit_wait();
        WR_Q_END
        WR_SYNC();
        WR_Q_BEGIN
        if (k > 0) fin_issue(q, valid, sh, (k - 1) & 1, e, k - 1);
        const double* d = sh->PK + (k & 1) * 72 + q * 18;
        const double dtc = (q >= 2 && k < N - 1) ? dt : 0.0;
        const double rc = q == 3 ? sh->cst[0] : 1.0, rs = q == 3 ? sh->cst[1] : 0.0;
        const double* X = nus + 3 * (q & 1);
        double s[3];
        s[0] = dtc * (rc * X[0] - rs * X[1]);
        s[1] = dtc * (rs * X[0] + rc * X[1]);
        s[2] = dtc * X[2];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double xr = d[3 + i];
            const double dd = d[i] - xr;
            s[i] += (k < N - 1 ? t.ax[i] : 0.0) - 2.0 * tb.Q[3 * q + i] * dd;
            t.red[2] += tb.Q[3 * q + i] * (dd * dd - xr * xr);
        }
        t.qa[0] = s[0]; t.qa[1] = s[1]; t.qa[2] = s[2];
        WR_Q_END
        WR_SYNC();            // everybody has read the co-state of step k+1
        WR_Q_BEGIN
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            t.ax[i] = t.qa[i];
            nus[3 * q + i] = t.qa[i];
            if (valid && e.bt->nu) e.bt->nu[(size_t)e.b * 12 * N + 12 * k + 3 * q + i] = t.qa[i];
        }
        WR_Q_END
        WR_SYNC();
        WR_Q_BEGIN
        if (valid) {
            const double* d = sh->PK + (k & 1) * 72 + q * 18;
            double bn[6];
            bbar_t(sh->cst, dt, h, nus, bn);
            double W[9], s3[3];
            foot_W(sh->cst + 3, d + 6, W);
#pragma unroll
            for (int c = 0; c < 3; ++c) s3[c] = bn[c] + W[c] * bn[3] + W[3 + c] * bn[4] + W[6 + c] * bn[5];
            double* yo = out_y(e);
            if (cur[4 * k + q] != SWING) {
                const double* f = d + 9;
                const double* yf = d + 13;
                const double l0 = -d[12];
                const double atl[3] = {yf[0] - yf[1], yf[2] - yf[3], -l0 - p.mu * (yf[0] + yf[1] + yf[2] + yf[3])};
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    t.red[0] = fmax(t.red[0], fabs(2.0 * tb.R[3 * q + c] * f[c] - s3[c] + atl[c]));
                    t.red[2] += tb.R[3 * q + c] * f[c] * f[c];
                }
                const double v = fmax(p.fz_min - f[2], fmax(fabs(f[0]), fabs(f[1])) - p.mu * f[2]);
                t.red[1] = fmax(t.red[1], v);
                t.red[3] += (l0 > 0.0) + (yf[0] > 0.0) + (yf[1] > 0.0) + (yf[2] > 0.0) + (yf[3] > 0.0);
                // multipliers must not be negative (they are > 0 on the working set by construction)
                t.red[0] = fmax(t.red[0], fmax(fmax(-l0, -yf[0]), fmax(fmax(-yf[1], -yf[2]), -yf[3])));
                yo[12 * k + 3 * q] = 0.0;
                yo[12 * k + 3 * q + 1] = 0.0;
            } else {
                yo[12 * k + 3 * q] = s3[0];
                yo[12 * k + 3 * q + 1] = s3[1];
                yo[12 * k + 3 * q + 2] = s3[2];
            }
        }
        WR_Q_END
    }
    WR_SYNC();
    double* R4 = sh->E1;
    WR_Q_BEGIN
#pragma unroll
    for (int i = 0; i < 4; ++i) R4[4 * q + i] = t.red[i];
    WR_Q_END
    WR_SYNC();
    const double rd = fmax(fmax(R4[0], R4[4]), fmax(R4[8], R4[12]));
    const double rp = fmax(fmax(R4[1], R4[5]), fmax(R4[9], R4[13]));
    const double obj = R4[2] + R4[6] + R4[10] + R4[14];
    const double na = R4[3] + R4[7] + R4[11] + R4[15];
    const int ok = (rd <= 1e-6) && (rp <= 1e-9);
    WR_Q_BEGIN
    (void)t;
    if (valid && ok && q == 0) {
        double* rho_p = e.bt->rho ? e.bt->rho + e.b : nullptr;
        const double rho = (warm && rho_p && *rho_p > 0.0) ? *rho_p : p.rho0;
        if (rho_p) *rho_p = rho;
        double* st = out_stats(e);
        e.bt->status[e.b] = ST_SOLVED;
        e.bt->iters[e.b] = 0;
        st[0] = fmax(rp, 0.0);
        st[1] = rd;
        st[2] = obj;
        st[3] = (double)(3 * nst);
        st[4] = na;
        st[5] = rho;
        st[6] = (double)(sweeps - 1);
        st[7] = (double)(sweeps > 1 ? (int)PATH_WRENCH : (int)PATH_RICCATI);
    }
    WR_Q_END
    WR_SYNC();
    return ok;
}

// Device route: the sweep kernel only marks a settled robot; wrench_certificate_kernel (cmpc.cu, certify_group below) forms
// the certificate and the remaining outputs for all settled robots of the batch at full memory parallelism.
CMPC_HD void mark_pending(int qlane, bool conv, const Env& e, int nst, int sweeps) {
    if (conv && qlane == 0) {
        e.bt->status[e.b] = ST_PENDING;
        double* st = out_stats(e);
        st[3] = (double)(3 * nst);
        st[6] = (double)(sweeps - 1);
    }
}

constexpr int kSingleMax = 8;        // single exchanges after the block budget before the robot is handed to the condensed kernel
constexpr int kBreakMax = 6;         // one-off single exchanges (cycle breakers) before the robot stays in single-exchange mode
constexpr int kHandOffIter = 6, kHandOffRows = 12;
constexpr int kDampAfter = 8;        // block updates are one-sided on every other foot-step from this sweep on
constexpr int kBlockExtra = 12;      // block updates: pdas_max_iter + kBlockExtra sweeps (damped updates of disturbed robots take up to ~20)

// Working-set policy after a sweep (the same in the device kernel and in solve_robot).  Block updates by the primal-dual
// rule settle the nominal robots in 1-6 sweeps; on heavily disturbed robots (40+ active rows) the plain rule oscillates,
// so from sweep kDampAfter on the updates are damped (forward_sweep, `damp`) -- all sampled disturbed robots then settle in
// 9-20 sweeps instead of not at all.  A working set that comes back (hash history, any period up to 8) is moved off the
// cycle by ONE single exchange (most negative multiplier out, else most violated row in); after kBreakMax such breaks, or
// when the block budget is spent, single exchanges only.
// fl: flags of forward_sweep.  Returns 1 = converged, 0 = go on, 2 = hand over to the condensed kernel; `single` tells which
// update to apply.
struct Policy { int nbreak; bool latched; int damp_from; };
CMPC_HD Policy policy_init() { return Policy{0, false, kDampAfter}; }
CMPC_HD bool policy_damp(const Policy& pl, int it) { return it >= pl.damp_from; }
CMPC_HD int policy_step(Policy& pl, int fl, int it, int max_it, bool handoff, bool& single) {
    single = pl.latched;
    const bool damped = policy_damp(pl, it);
    if (pl.latched || damped) { if (!(fl & 4)) return 1; }      // nothing to exchange: the Karush-Kuhn-Tucker conditions hold
    else if (!(fl & 1)) return 1;
    if (pl.latched) return 0;
    // a lightly constrained robot that is still moving rows after kHandOffIter sweeps (0.06 % of the nominal workload: cycles
    // of period 3-5) costs the condensed kernel ~60 us but would keep its warp alive for several more sweeps at the end of
    // the batch: hand it over.  Heavily constrained robots stay (the condensed kernel needs milliseconds for those).
    // Only where the condensed kernel keeps its factor in shared memory (N <= 16): at N = 32 / 48 it costs milliseconds too.
    if (handoff && it + 1 >= kHandOffIter && (fl >> 16) <= kHandOffRows) return 2;
    const bool budget = it + 1 >= max_it + kBlockExtra;
    if (((fl & 1) && (fl & 2)) || budget) {
        if (!damped && !budget) { pl.damp_from = it + 1; return 0; }      // first answer to a cycle: damp the updates from now on
        single = true;
        if (++pl.nbreak > kBreakMax || budget) pl.latched = true;
        if (!(fl & 4)) return 1;
    }
    return 0;
}

// One robot from set-up to outputs (host emulation; the device kernel in cmpc.cu interleaves these steps over the
// eight robots of a warp).  Returns 1 if finished here, 0 if the robot goes on to the condensed kernel.
CMPC_HD int solve_robot(int qlane, TS* ts, Sh* sh, const Env& e, int nfmax, int warm, int* sweeps_out) {
    const int N = e.bt->N;
    unsigned char* codes = codes_of(sh);
    const int nst = init_robot(qlane, true, ts, sh, e, warm);
    if (sweeps_out) *sweeps_out = 0;
    if (nst == 0 || nst > nfmax || (N & 3)) return 0;     // lever arms are fetched four stages at a time
    const int max_it = e.p->pdas_max_iter;
    Policy pl = policy_init();
    for (int it = 0; it < max_it + kBlockExtra + kSingleMax; ++it) {
        const unsigned char* cur = codes + (size_t)(it % 3) * 4 * N;
        unsigned char* next = codes + (size_t)((it + 1) % 3) * 4 * N;
        const double pmin = backward_sweep(qlane, ts, sh, e, cur);
        if (!(pmin > 0.0)) return 0;
        const int fl = forward_sweep(qlane, true, ts, sh, e, cur, next, it, policy_damp(pl, it));
        if (sweeps_out) *sweeps_out = it + 1;
        bool single;
        const int ps = policy_step(pl, fl, it, max_it, N <= 16, single);
        if (ps == 1) return finish_robot(qlane, true, ts, sh, e, cur, warm, nst, it + 1);
        if (ps == 2) return 0;
        single_step(qlane, single, ts, sh, cur, next, N);
    }
    return 0;
}

#if defined(__CUDACC__)
// ------------------------------------------------------------------------------------------------------------------
// Certificate of one settled robot by SIXTEEN lanes (lane l owns the stages l, l + 16, l + 32): co-states, stationarity and
// feasibility from first principles -- X, u, y as the last forward sweep wrote them, independent of the factorizations --
// and the remaining outputs in the reference's layouts; the same quantities as finish_robot (the host emulation's
// version).  A = I + dt E with E^2 = 0, so the co-state recursion nu_k = A'nu_{k+1} - 2 Q (x_{k+1} - xref_k) has the closed
// form  nu_k[p, rpy] = -2 S_k,  nu_k[v, omega] = -2 S_k + dt Rz' sum_{j > k} nu_j[p, rpy]  with suffix sums S over the
// stages: two suffix scans over the sixteen lanes instead of a sequential sweep, and every global access is coalesced
// across the lanes.  Returns (to every lane) 1 if the certificate holds.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double grp_suffix_sum(double v, unsigned gmask, int gl) {       // inclusive, from lane gl up to 15
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
        const double t = __shfl_down_sync(gmask, v, o, 16);
        if (gl + o < 16) v += t;
    }
    return v;
}
__device__ __forceinline__ double grp_max(double v, unsigned gmask) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(gmask, v, o, 16));
    return v;
}
__device__ __forceinline__ double grp_sum(double v, unsigned gmask) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o, 16);
    return v;
}
__device__ inline int certify_group(int gl, unsigned gmask, const Params& p, const Bat& bt, int b, int warm) {
    const int N = bt.N;
    const double dt = bt.dt, h = dt * dt / 2.0;
    const double* __restrict__ xr = bt.x_ref + (size_t)b * 12 * N;
    const double* __restrict__ rf = bt.r_foot + (size_t)b * 12 * N;
    const double* __restrict__ Xo = bt.X + (size_t)b * 12 * N;
    const double* __restrict__ uo = bt.u + (size_t)b * 12 * N;
    double* yo = bt.y + (size_t)b * 28 * N;
    const uint64_t* mk = bt.mask ? bt.mask + (size_t)b * bt.W : nullptr;
    // constants of the robot (every lane; the loads are broadcasts)
    double cst[12];
    if (bt.cst) {
#pragma unroll
        for (int i = 0; i < 12; ++i) cst[i] = bt.cst[(size_t)b * 12 + i];
    } else {
        double sum = 0.0;
        for (int k = gl; k < N; k += 16) sum += xr[(size_t)5 * N + k];
        const double yaw = grp_sum(sum, gmask) / (double)N;
        DynCommon dc;
        dyn_common(dc, xr, 0, bt.I_world + (size_t)b * 9, bt.mass[b], dt);        // N = 0: yaw is formed above
        cst[0] = cos(yaw); cst[1] = sin(yaw); cst[2] = dc.minv;
#pragma unroll
        for (int i = 0; i < 9; ++i) cst[3 + i] = dc.Iinv[i];
    }
    const double cy = cst[0], sy = cst[1];
    double rd = 0.0, rp = 0.0, obj = 0.0, na = 0.0;
    double carry0[12], carry1[6];          // suffix sums of the chunks above this one: S (12) and sum of nu[p, rpy] (6)
#pragma unroll
    for (int i = 0; i < 12; ++i) carry0[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) carry1[i] = 0.0;
    const int nchunk = (N + 15) >> 4;
    for (int ch = nchunk - 1; ch >= 0; --ch) {
        const int k = ch * 16 + gl;
        const bool act = k < N;
        const int kk = act ? k : N - 1;
        double S[12], nu[12];
        // everything the stage reads, up front: one round of memory latency per chunk
        double xr12[12], x12[12], f12[12], r12[12], yb4[4], yf16[16];
#pragma unroll
        for (int i = 0; i < 12; ++i) { xr12[i] = xr[(size_t)i * N + kk]; r12[i] = rf[(size_t)i * N + kk]; }
        {
            const double2* xp = reinterpret_cast<const double2*>(Xo + 12 * kk);
            const double2* up = reinterpret_cast<const double2*>(uo + 12 * kk);
            const double2* yp = reinterpret_cast<const double2*>(yo + 12 * N + 16 * kk);
#pragma unroll
            for (int i = 0; i < 6; ++i) { const double2 a = xp[i], c2 = up[i]; x12[2 * i] = a.x; x12[2 * i + 1] = a.y; f12[2 * i] = c2.x; f12[2 * i + 1] = c2.y; }
#pragma unroll
            for (int i = 0; i < 8; ++i) { const double2 a = yp[i]; yf16[2 * i] = a.x; yf16[2 * i + 1] = a.y; }
#pragma unroll
            for (int j = 0; j < 4; ++j) yb4[j] = yo[12 * kk + 3 * j + 2];
        }
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            const double r = xr12[i];
            const double dd = x12[i] - r;
            const double qe = act ? p.Q[i] * dd : 0.0;
            obj += act ? p.Q[i] * (dd * dd - r * r) : 0.0;
            S[i] = grp_suffix_sum(qe, gmask, gl) + carry0[i];
        }
        // nu[p, rpy] = -2 S;  T = sum_{j > k} nu_j[p, rpy] = (inclusive suffix sum of nu) - nu_k
        double T[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            nu[i] = -2.0 * S[i];
            T[i] = grp_suffix_sum(act ? nu[i] : 0.0, gmask, gl) + carry1[i] - (act ? nu[i] : 0.0);
        }
        nu[6] = -2.0 * S[6] + dt * T[0];
        nu[7] = -2.0 * S[7] + dt * T[1];
        nu[8] = -2.0 * S[8] + dt * T[2];
        nu[9] = -2.0 * S[9] + dt * (cy * T[3] - sy * T[4]);
        nu[10] = -2.0 * S[10] + dt * (sy * T[3] + cy * T[4]);
        nu[11] = -2.0 * S[11] + dt * T[5];
        // carries for the chunk below: totals of this chunk (lane 0 holds the inclusive sums over the whole chunk)
#pragma unroll
        for (int i = 0; i < 12; ++i) carry0[i] = __shfl_sync(gmask, S[i], 0, 16);
#pragma unroll
        for (int i = 0; i < 6; ++i) carry1[i] = __shfl_sync(gmask, T[i] + nu[i], 0, 16);
        if (act) {
            if (bt.nu) {
                double2* no = reinterpret_cast<double2*>(bt.nu + (size_t)b * 12 * N + 12 * k);
#pragma unroll
                for (int i = 0; i < 6; ++i) no[i] = make_double2(nu[2 * i], nu[2 * i + 1]);
            }
            double bn[6];
            bbar_t(cst, dt, h, nu, bn);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double r3[3] = {r12[3 * j], r12[3 * j + 1], r12[3 * j + 2]};
                double W[9], s3[3];
                foot_W(cst + 3, r3, W);
#pragma unroll
                for (int c = 0; c < 3; ++c) s3[c] = bn[c] + W[c] * bn[3] + W[3 + c] * bn[4] + W[6 + c] * bn[5];
                if (mask_bit(mk, N, j, k)) {
                    const double* f = f12 + 3 * j;
                    const double l0 = -yb4[j];
                    const double y0 = yf16[4 * j], y1 = yf16[4 * j + 1], y2 = yf16[4 * j + 2], y3 = yf16[4 * j + 3];
                    const double atl[3] = {y0 - y1, y2 - y3, -l0 - p.mu * (y0 + y1 + y2 + y3)};
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        rd = fmax(rd, fabs(2.0 * p.R[3 * j + c] * f[c] - s3[c] + atl[c]));
                        obj += p.R[3 * j + c] * f[c] * f[c];
                    }
                    rp = fmax(rp, fmax(p.fz_min - f[2], fmax(fabs(f[0]), fabs(f[1])) - p.mu * f[2]));
                    na += (l0 > 0.0) + (y0 > 0.0) + (y1 > 0.0) + (y2 > 0.0) + (y3 > 0.0);
                    rd = fmax(rd, fmax(fmax(-l0, -y0), fmax(fmax(-y1, -y2), -y3)));      // multipliers must not be negative
                    yo[12 * k + 3 * j] = 0.0;
                    yo[12 * k + 3 * j + 1] = 0.0;
                } else {
                    yo[12 * k + 3 * j] = s3[0];
                    yo[12 * k + 3 * j + 1] = s3[1];
                    yo[12 * k + 3 * j + 2] = s3[2];
                }
            }
        }
    }
    rd = grp_max(rd, gmask); rp = grp_max(rp, gmask); obj = grp_sum(obj, gmask); na = grp_sum(na, gmask);
    const int ok = (rd <= 1e-6) && (rp <= 1e-9);
    if (ok && gl == 0) {
        double* rho_p = bt.rho ? bt.rho + b : nullptr;
        const double rho = (warm && rho_p && *rho_p > 0.0) ? *rho_p : p.rho0;
        if (rho_p) *rho_p = rho;
        double* st = bt.stats + (size_t)b * NSTAT;
        const int sweeps = (int)st[6] + 1;
        bt.status[b] = ST_SOLVED;
        bt.iters[b] = 0;
        st[0] = fmax(rp, 0.0);
        st[1] = rd;
        st[2] = obj;
        st[4] = na;
        st[5] = rho;
        st[7] = (double)(sweeps > 1 ? (int)PATH_WRENCH : (int)PATH_RICCATI);
    }
    return ok;
}
#endif

CMPC_HD void fill_tab(Tab& tb, const Params& p, int i) {
    // entry i of the tables (i < 12: Q, R, Rinv; i < 16: sig)
    if (i < 12) { tb.Q[i] = p.Q[i]; tb.R[i] = p.R[i]; tb.Rinv[i] = 1.0 / p.R[i]; }
    if (i < 16) {
        const int leg = i >> 2, tx = (i >> 1) & 1, ty = i & 1;
        tb.sig[i] = 1.0 / (p.R[3 * leg + 2] + (tx ? p.mu * p.mu * p.R[3 * leg] : 0.0) + (ty ? p.mu * p.mu * p.R[3 * leg + 1] : 0.0));
    }
}

}  // namespace wr
}  // namespace cmpc
