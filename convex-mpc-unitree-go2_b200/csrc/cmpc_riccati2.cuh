// cmpc_riccati2.cuh -- Riccati pre-pass, device version 2: TWO robots per warp (one per 16-lane half), lane i of a
// half owns state row i (and, in its second role, input row i of the stage), the 12 x 12 cost-to-go row lives in
// registers, shared memory is only the exchange medium for rows that other lanes need as broadcast operands.
// Same recursion, same outputs and same accept / reject rule as ric::riccati_one (cmpc_riccati.cuh), which stays
// the reference implementation for the host emulation; the GPU tests compare the two paths.
//
// Per stage k (m = 3 x stance legs, B = B_d[k] stance columns, all loops unrolled, branches uniform per half):
//   PB_i   = P_i B                      own row x structured B columns (U, W of the stance legs: broadcast loads)
//   T_a    = column a of PB             (role switch through shared memory; T = B'P by symmetry of P)
//   G_a    = T_a B + R e_a,  S_a = T_a A,  bq_a = B_a'q
//   G = L L',  W = inv(L)               row per lane, 16-wide shuffles, rsqrt chain (as diag_factor of cmpc_fast.cuh)
//   Y_j    = W_j S,  K_a = (W'Y)_a,  kff = W'W bq
//   P_i   <- Q e_i + (A'(P A))_i - sum_a S_ai K_a ,   p_i <- -Q_i xref_i + (A'q)_i - sum_a S_ai kff_a
#pragma once
#include "cmpc_riccati.cuh"

#if defined(__CUDACC__)
namespace cmpc {
namespace ric2 {

constexpr int LD = 14;
constexpr int MAT = 12 * LD;
#ifndef CMPC_RIC_SYNC_EVERY
#define CMPC_RIC_SYNC_EVERY 8     // lock-step granularity: one CTA barrier every this many stages (measured: 1 -> 10.4 ms,
                                  // 2 -> 10.0, 4 -> 9.7, 8 / 16 / once per sweep -> 9.4 ms per 65 536 robots)
#endif

struct WsH {          // shared memory of one robot (half-warp)
    double* RF;       // 12N lever arms, index (leg*3 + a)*N + k
    double* XR;       // 12N reference, index i*12 + r ; later the staged box multipliers
    double* X;        // 12N rolled-out states
    double* NU;       // 12N co-states
    double* U;        // 12N forces, output layout 12k + 3 leg + comp
    double* M0;       // PB rows -> Y rows -> N rows
    double* M1;       // W (row j, column c) -> K rows
    double* M2;       // S rows
    double* UW;       // 4 x 18
    double* v0;       // 16: q / x_k / nu_k
    double* v1;       // 16: bq -> yv -> kff / u_k
    double* dynv;     // 16: cy, sy, Iinv[9], minv
    int* vstart;      // N+1
};

__device__ __forceinline__ size_t carve_half(WsH& w, unsigned char* base, int N) {
    double* p = reinterpret_cast<double*>(base);
    auto take = [&](size_t n) { double* r = p; p += (n + 1) & ~(size_t)1; return r; };
    w.RF = take((size_t)12 * N); w.XR = take((size_t)12 * N); w.X = take((size_t)12 * N);
    w.NU = take((size_t)12 * N); w.U = take((size_t)12 * N);
    w.M0 = take(MAT); w.M1 = take(MAT); w.M2 = take(MAT);
    w.UW = take(72); w.v0 = take(16); w.v1 = take(16); w.dynv = take(16);
    w.vstart = reinterpret_cast<int*>(p);
    return (size_t)(reinterpret_cast<unsigned char*>(w.vstart + ((N + 1 + 3) & ~3)) - base);
}
inline size_t half_bytes(int N) {
    const size_t d = (size_t)5 * 12 * N + 3 * MAT + 72 + 16 * 3;
    return (d * 8 + (size_t)((N + 1 + 3) & ~3) * 4 + 15) & ~(size_t)15;
}

__device__ __forceinline__ double shfl16(unsigned hmask, double v, int src) {
    const int lo = __shfl_sync(hmask, __double2loint(v), src, 16);
    const int hi = __shfl_sync(hmask, __double2hiint(v), src, 16);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double max16(unsigned hmask, double v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        const int lo = __shfl_xor_sync(hmask, __double2loint(v), o, 16);
        const int hi = __shfl_xor_sync(hmask, __double2hiint(v), o, 16);
        v = fmax(v, __hiloint2double(hi, lo));
    }
    return v;
}
__device__ __forceinline__ double sum16(unsigned hmask, double v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        const int lo = __shfl_xor_sync(hmask, __double2loint(v), o, 16);
        const int hi = __shfl_xor_sync(hmask, __double2hiint(v), o, 16);
        v += __hiloint2double(hi, lo);
    }
    return v;
}

struct Dyn { double cy, sy, minv, dt, h; };

// reciprocal square root of a normal positive double: MUFU.RSQ64H seed + one third-order correction (the CUDA
// library's sequence without its special-case branch; a non-positive pivot is caught by the caller)
__device__ __forceinline__ double rsqrt_fast(double sv) {
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(sv));
    const double t = y0 * y0;
    const double e = fma(-t, sv, 1.0);
    const double pl = fma(e, 0.375, 0.5);
    const double u = y0 * e;
    return fma(pl, u, y0);
}

// stance bit (leg, k) from the robot's mask words held in registers (W <= 3 words, bit leg*N + k)
__device__ __forceinline__ int mbit(unsigned long long m0, unsigned long long m1, unsigned long long m2, int N, int leg, int k) {
    const int b = leg * N + k;
    const unsigned long long wv = (b < 64) ? m0 : (b < 128 ? m1 : m2);
    return (int)((wv >> (b & 63)) & 1ull);
}

// out[a] = v . B[:, a]  for the compact stance columns a < m (unrolled; U, W of the stance legs are broadcast loads)
__device__ __forceinline__ void row_times_B(const double v[12], double out[12], int m, const int legs[4], const double* UW,
                                            const Dyn& d) {
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        if (3 * s < m) {
            const double* uw = UW + 18 * legs[s];
            double u9[9], w9[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) { u9[q] = uw[q]; w9[q] = uw[9 + q]; }
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) {
                double a = d.h * d.minv * v[cc] + d.dt * d.minv * v[6 + cc];
                double b = 0.0;
#pragma unroll
                for (int r = 0; r < 3; ++r) { a += d.h * (v[3 + r] * u9[r * 3 + cc]); b += v[9 + r] * w9[r * 3 + cc]; }
                out[3 * s + cc] = a + d.dt * b;
            }
        } else {
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) out[3 * s + cc] = 0.0;
        }
    }
}

// U, W of the four legs at step k -> UW (lanes 0..3), followed by a half-warp barrier
__device__ __forceinline__ void leg_mats(unsigned hmask, int hl, const WsH& w, int N, int k, const Dyn& d) {
    if (hl < 4) {
        const double r0 = w.RF[(hl * 3 + 0) * N + k], r1 = w.RF[(hl * 3 + 1) * N + k], r2 = w.RF[(hl * 3 + 2) * N + k];
        const double sk[9] = {0.0, -r2, r1, r2, 0.0, -r0, -r1, r0, 0.0};
        const double* Ii = w.dynv + 2;
        double Wm[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) Wm[i * 3 + j] = Ii[i * 3] * sk[j] + Ii[i * 3 + 1] * sk[3 + j] + Ii[i * 3 + 2] * sk[6 + j];
        double* o = w.UW + 18 * hl;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            o[j] = d.cy * Wm[j] + d.sy * Wm[3 + j];
            o[3 + j] = -d.sy * Wm[j] + d.cy * Wm[3 + j];
            o[6 + j] = Wm[6 + j];
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) o[9 + i] = Wm[i];
    }
    __syncwarp(hmask);
}

// One backward stage.  MT > 0: the number of stance variables is the compile-time constant MT (6 and 12 cover the
// trot), every guard below folds away and the stage is straight-line code; MT == 0: run-time m.
template <int MT>
__device__ __forceinline__ void back_stage(unsigned hmask, int hl, int i, bool act, const Params& p, WsH& w, const Dyn& d,
                                           int N, int k, int m_rt, const int legs[4], double (&Pr)[12], double& pvi,
                                           double& pmin, double* Kst, double* kst, int ra, int rb, double ca, double cb,
                                           double Qi, double gz2, double gz8) {
    const int m = MT ? MT : m_rt;
        const int voff = w.vstart[k];
        const double qi = Pr[2] * gz2 + Pr[8] * gz8 + pvi;          // q = P g + p
        double Sa[12], Ka[12];
        double kffa = 0.0;
#pragma unroll
        for (int c = 0; c < 12; ++c) { Sa[c] = 0.0; Ka[c] = 0.0; }
        if (m > 0) {
            leg_mats(hmask, hl, w, N, k, d);
            {   // PB row -> M0 ; q -> v0
                double PB[12];
                row_times_B(Pr, PB, m, legs, w.UW, d);
                if (act) {
#pragma unroll
                    for (int c = 0; c < 12; c += 2) *reinterpret_cast<double2*>(w.M0 + i * LD + c) = make_double2(PB[c], PB[c + 1]);
                    w.v0[i] = qi;
                }
            }
            __syncwarp(hmask);
            // second role: input row a = i
            double Ta[12], Ga[12];
#pragma unroll
            for (int c = 0; c < 12; ++c) Ta[c] = w.M0[c * LD + i];
            row_times_B(Ta, Ga, m, legs, w.UW, d);
            {
                const int sl = i / 3, cc = i - 3 * sl;
                const int lg = sl == 0 ? legs[0] : (sl == 1 ? legs[1] : (sl == 2 ? legs[2] : legs[3]));
                const double Rv = p.R[3 * lg + cc];
                // S_a = T_a A ; bq_a = B_a' q ; unit rows beyond m keep the factorisation well defined
#pragma unroll
                for (int c = 0; c < 12; ++c) Sa[c] = Ta[c];
#pragma unroll
                for (int c = 0; c < 3; ++c) Sa[6 + c] += d.dt * Ta[c];
                Sa[9] += d.dt * (d.cy * Ta[3] - d.sy * Ta[4]);
                Sa[10] += d.dt * (d.sy * Ta[3] + d.cy * Ta[4]);
                Sa[11] += d.dt * Ta[5];
                const double* uw = w.UW + 18 * lg;
                double bq = d.h * d.minv * w.v0[cc] + d.dt * d.minv * w.v0[6 + cc];
#pragma unroll
                for (int r = 0; r < 3; ++r) bq += d.h * uw[r * 3 + cc] * w.v0[3 + r] + d.dt * uw[9 + r * 3 + cc] * w.v0[9 + r];
                const bool row = act && i < m;
#pragma unroll
                for (int c = 0; c < 12; ++c) {
                    if (c == i) Ga[c] += Rv;
                    if (!row) { Ga[c] = (c == i) ? 1.0 : 0.0; Sa[c] = 0.0; }
                }
                if (act) w.v1[i] = row ? bq : 0.0;
            }
            // Cholesky of G with inverse: lane a holds row a of G, ends with column a of W = inv(L)
            double Wc[12];
            {
                double sw[12];
#pragma unroll
                for (int c = 0; c < 12; ++c) { sw[c] = 0.0; Wc[c] = 0.0; }
                double piv = shfl16(hmask, Ga[0], 0);
#pragma unroll
                for (int j = 0; j < 12; ++j) {
                    if (j < m) {
                        if (!(piv > 0.0)) pmin = -1.0;
                        const double dj = rsqrt_fast(piv);
                        const double l = Ga[j] * dj;
                        if (j + 1 < 12) piv = shfl16(hmask, Ga[(j + 1) % 12] - l * l, (j + 1) % 12);   // next pivot goes out first
                        const double wj = (j == hl) ? dj : -dj * sw[j];
                        Wc[j] = wj;
#pragma unroll
                        for (int c = j + 1; c < 12; ++c) {
                            if (c < m) {
                                const double lk = shfl16(hmask, l, c);
                                Ga[c] -= l * lk;
                                sw[c] += lk * wj;
                            }
                        }
                    }
                }
            }
            // W (row j, column c) -> M1 ; S rows -> M2
            if (act) {
#pragma unroll
                for (int j = 0; j < 12; ++j) w.M1[j * LD + i] = (j >= i && j < m && i < m) ? Wc[j] : 0.0;
#pragma unroll
                for (int c = 0; c < 12; c += 2) *reinterpret_cast<double2*>(w.M2 + i * LD + c) = make_double2(Sa[c], Sa[c + 1]);
            }
            __syncwarp(hmask);
            // Y_j = W_j S ; yv_j = W_j bq     (row j = i)
            double Yj[12];
            double yv = 0.0;
#pragma unroll
            for (int c = 0; c < 12; ++c) Yj[c] = 0.0;
#pragma unroll 2
            for (int b = 0; b < m; ++b) {               // rolled on purpose: the kernel is instruction-fetch bound
                const double wv = w.M1[i * LD + b];
                yv += wv * w.v1[b];
#pragma unroll
                for (int c = 0; c < 12; c += 2) {
                    const double2 s2 = *reinterpret_cast<const double2*>(w.M2 + b * LD + c);
                    Yj[c] += wv * s2.x; Yj[c + 1] += wv * s2.y;
                }
            }
            __syncwarp(hmask);                 // everybody has read W rows and bq
            if (act) {
#pragma unroll
                for (int c = 0; c < 12; c += 2) *reinterpret_cast<double2*>(w.M0 + i * LD + c) = make_double2(Yj[c], Yj[c + 1]);
                w.v1[i] = yv;
            }
            __syncwarp(hmask);
            // K_a = sum_j W_ja Y_j ; kff_a = sum_j W_ja yv_j     (own column of W in registers)
#pragma unroll 2
            for (int j = 0; j < m; ++j) {               // W_ja from shared memory (M1 still holds W; zero above the diagonal)
                const double wv = w.M1[j * LD + i];
                kffa += wv * w.v1[j];
#pragma unroll
                for (int c = 0; c < 12; c += 2) {
                    const double2 y2 = *reinterpret_cast<const double2*>(w.M0 + j * LD + c);
                    Ka[c] += wv * y2.x; Ka[c + 1] += wv * y2.y;
                }
            }
            if (!(act && i < m)) {
                kffa = 0.0;
#pragma unroll
                for (int c = 0; c < 12; ++c) Ka[c] = 0.0;
            }
            __syncwarp(hmask);                 // everybody has read Y rows and yv
            if (act && i < m) {
                double* kr = Kst + (size_t)(voff + i) * 12;
#pragma unroll
                for (int c = 0; c < 12; c += 2) *reinterpret_cast<double2*>(kr + c) = make_double2(Ka[c], Ka[c + 1]);
                kst[voff + i] = kffa;
            }
        } else {
            if (act) w.v0[i] = qi;
        }
        if (k == 0) return;
        // N = P A (own row) -> M0 ; K rows -> M1 ; kff -> v1
        {
            double Ni[12];
#pragma unroll
            for (int c = 0; c < 12; ++c) Ni[c] = Pr[c];
#pragma unroll
            for (int c = 0; c < 3; ++c) Ni[6 + c] += d.dt * Pr[c];
            Ni[9] += d.dt * (d.cy * Pr[3] - d.sy * Pr[4]);
            Ni[10] += d.dt * (d.sy * Pr[3] + d.cy * Pr[4]);
            Ni[11] += d.dt * Pr[5];
            if (act) {
#pragma unroll
                for (int c = 0; c < 12; c += 2) {
                    *reinterpret_cast<double2*>(w.M0 + i * LD + c) = make_double2(Ni[c], Ni[c + 1]);
                    *reinterpret_cast<double2*>(w.M1 + i * LD + c) = make_double2(Ka[c], Ka[c + 1]);
                }
                w.v1[i] = kffa;
            }
            __syncwarp(hmask);
            // P_i <- Q e_i + N_i + ca N_ra + cb N_rb - sum_a S_ai K_a ;  p_i likewise
            double pn = -Qi * w.XR[(k - 1) * 12 + i] + qi + ca * w.v0[ra] + cb * w.v0[rb];
#pragma unroll
            for (int c = 0; c < 12; c += 2) {
                const double2 na = *reinterpret_cast<const double2*>(w.M0 + ra * LD + c);
                const double2 nb = *reinterpret_cast<const double2*>(w.M0 + rb * LD + c);
                Pr[c] = Ni[c] + ca * na.x + cb * nb.x + ((c == i) ? Qi : 0.0);
                Pr[c + 1] = Ni[c + 1] + ca * na.y + cb * nb.y + ((c + 1 == i) ? Qi : 0.0);
            }
#pragma unroll 2
            for (int a = 0; a < m; ++a) {
                const double s = w.M2[a * LD + i];
                pn -= s * w.v1[a];
#pragma unroll
                for (int c = 0; c < 12; c += 2) {
                    const double2 k2 = *reinterpret_cast<const double2*>(w.M1 + a * LD + c);
                    Pr[c] -= s * k2.x; Pr[c + 1] -= s * k2.y;
                }
            }
            pvi = pn;
            __syncwarp(hmask);                 // M0, M1, M2, v0, v1 are free for the next stage
        }
    }

// One robot on one half-warp.  Returns 1 if finished here (outputs written), 0 if it goes on to the condensed path.
// LS (lock-step): the halves of a CTA meet at a CTA barrier before every backward stage and forward step, so that
// they run the same stretch of code at the same time and share fetched instruction lines (the sweep is
// instruction-fetch bound).  Every half then has to reach every barrier: early exits become `alive = false`, and
// halves without a robot (valid == false) walk through the loops idle.  Returns -1 for those.
template <bool LS>
__device__ __forceinline__ int riccati_half(unsigned hmask, int hl, bool valid, const Params& p, const QpIn& in, QpOut& o,
                                            WsH& w, int nfmax, int warm, double* gains) {
    const int N = in.N;
    const int W = (4 * N + 63) >> 6;
    bool alive = valid;
    const unsigned long long mk0 = (alive && in.mask) ? in.mask[0] : ~0ull, mk1 = (alive && in.mask && W > 1) ? in.mask[1] : ~0ull,
                             mk2 = (alive && in.mask && W > 2) ? in.mask[2] : ~0ull;
    const int i = hl < 12 ? hl : 11;            // lanes 12..15 shadow row 11 (reads stay in bounds, writes are masked)
    const bool act = hl < 12;
    // ---- set-up
    if (alive && hl == 0) {
        DynCommon dc;
        dyn_common(dc, in.x_ref, N, in.I_world, in.mass, in.dt);
        w.dynv[0] = dc.cy; w.dynv[1] = dc.sy;
        for (int q = 0; q < 9; ++q) w.dynv[2 + q] = dc.Iinv[q];
        w.dynv[11] = dc.minv;
        int nf = 0;
        for (int k = 0; k < N; ++k) {
            w.vstart[k] = 3 * nf;
            for (int leg = 0; leg < 4; ++leg) nf += mbit(mk0, mk1, mk2, N, leg, k);
        }
        w.vstart[N] = 3 * nf;
    }
    for (int idx = hl; alive && idx < 12 * N; idx += 16) {
        const int r = idx / N, c = idx - r * N;
        w.XR[c * 12 + r] = in.x_ref[idx];
        w.RF[idx] = in.r_foot[idx];
        w.U[idx] = 0.0;
    }
    __syncwarp(hmask);
    const int n = alive ? w.vstart[N] : 0;
    if (n > 3 * nfmax || n == 0) { if (!LS) return 0; alive = false; }
    Dyn d;
    d.cy = w.dynv[0]; d.sy = w.dynv[1]; d.minv = w.dynv[11]; d.dt = in.dt; d.h = in.dt * in.dt / 2.0;
    const double gz2 = -9.81 * d.h, gz8 = -9.81 * d.dt;
    double* Kst = gains;
    double* kst = gains + (size_t)3 * nfmax * 12;
    // A^T coupling of row i:  (A'v)_i = v_i + ca v_ra + cb v_rb
    int ra = 0, rb = 0;
    double ca = 0.0, cb = 0.0;
    if (i >= 6 && i < 9) { ra = i - 6; ca = d.dt; }
    else if (i == 9) { ra = 3; ca = d.dt * d.cy; rb = 4; cb = -d.dt * d.sy; }
    else if (i == 10) { ra = 3; ca = d.dt * d.sy; rb = 4; cb = d.dt * d.cy; }
    else if (i == 11) { ra = 5; ca = d.dt; }
    const double Qi = p.Q[i];

    // ---- backward sweep
    double Pr[12];
#pragma unroll
    for (int c = 0; c < 12; ++c) Pr[c] = (c == i) ? Qi : 0.0;
    double pvi = -Qi * w.XR[(N - 1) * 12 + i];
    double pmin = 1.0;
    for (int k = N - 1; k >= 0; --k) {
        if (LS && ((N - 1 - k) % CMPC_RIC_SYNC_EVERY) == 0) __syncthreads();
        if (!alive) continue;
        int legs[4] = {0, 0, 0, 0};
        int cnt = 0;
#pragma unroll
        for (int leg = 0; leg < 4; ++leg)
            if (mbit(mk0, mk1, mk2, N, leg, k)) {
                if (cnt == 0) legs[0] = leg; else if (cnt == 1) legs[1] = leg; else if (cnt == 2) legs[2] = leg; else legs[3] = leg;
                ++cnt;
            }
        const int m = 3 * cnt;
        if (cnt == 2) back_stage<6>(hmask, hl, i, act, p, w, d, N, k, m, legs, Pr, pvi, pmin, Kst, kst, ra, rb, ca, cb, Qi, gz2, gz8);
        else if (cnt == 4) back_stage<12>(hmask, hl, i, act, p, w, d, N, k, m, legs, Pr, pvi, pmin, Kst, kst, ra, rb, ca, cb, Qi, gz2, gz8);
        else back_stage<0>(hmask, hl, i, act, p, w, d, N, k, m, legs, Pr, pvi, pmin, Kst, kst, ra, rb, ca, cb, Qi, gz2, gz8);
    }
    pmin = -max16(hmask, -pmin);
    if (!(pmin > 0.0)) { if (!LS) return 0; alive = false; }
    __syncwarp(hmask);

    // ---- forward sweep
    double xi = alive ? in.x0[i] : 0.0;
    for (int k = 0; k < N; ++k) {
        if (LS && (k % CMPC_RIC_SYNC_EVERY) == 0) __syncthreads();
        if (!alive) continue;
        int legs[4] = {0, 0, 0, 0};
        int cnt = 0;
#pragma unroll
        for (int leg = 0; leg < 4; ++leg)
            if (mbit(mk0, mk1, mk2, N, leg, k)) {
                if (cnt == 0) legs[0] = leg; else if (cnt == 1) legs[1] = leg; else if (cnt == 2) legs[2] = leg; else legs[3] = leg;
                ++cnt;
            }
        const int m = 3 * cnt;
        const int voff = w.vstart[k];
        const bool row = act && i < m;
        double kr[12];
        double kf = 0.0;
#pragma unroll
        for (int c = 0; c < 12; ++c) kr[c] = 0.0;
        if (row) {                              // own gain row (global scratch, L2) -- issued before the barrier
            const double* g = Kst + (size_t)(voff + i) * 12;
#pragma unroll
            for (int c = 0; c < 12; c += 2) { const double2 t2 = *reinterpret_cast<const double2*>(g + c); kr[c] = t2.x; kr[c + 1] = t2.y; }
            kf = kst[voff + i];
        }
        if (act) w.v0[i] = xi;
        if (m > 0) leg_mats(hmask, hl, w, N, k, d); else __syncwarp(hmask);
        double ua = kf;
#pragma unroll
        for (int c = 0; c < 12; ++c) ua += kr[c] * w.v0[c];
        ua = -ua;
        const int sl = i / 3, cc = i - 3 * sl;
        const int lg = sl == 0 ? legs[0] : (sl == 1 ? legs[1] : (sl == 2 ? legs[2] : legs[3]));
        if (row) { w.v1[i] = ua; w.U[12 * k + 3 * lg + cc] = ua; }
        else if (act) w.v1[i] = 0.0;
        __syncwarp(hmask);
        // x_{k+1,i} = (A x)_i + g_i + sum_a B_ia u_a
        double xn = xi;
        if (i < 3) xn += d.dt * w.v0[6 + i];
        else if (i == 3) xn += d.dt * (d.cy * w.v0[9] + d.sy * w.v0[10]);
        else if (i == 4) xn += d.dt * (-d.sy * w.v0[9] + d.cy * w.v0[10]);
        else if (i == 5) xn += d.dt * w.v0[11];
        if (i == 2) xn += gz2;
        if (i == 8) xn += gz8;
        {
            const int blk = i / 3, r = i - 3 * blk;       // 0: p, 1: rpy, 2: v, 3: omega
            const double lin = (blk == 0) ? d.h * d.minv : (blk == 2 ? d.dt * d.minv : 0.0);
            const double ang = (blk == 1) ? d.h : (blk == 3 ? d.dt : 0.0);
            const int off = (blk == 3) ? 9 : 0;           // omega rows use W, rpy rows use U
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                if (3 * s < m) {
                    const double* uw = w.UW + 18 * legs[s] + off + 3 * r;
                    const double u0 = w.v1[3 * s], u1 = w.v1[3 * s + 1], u2 = w.v1[3 * s + 2];
                    const double ur = (r == 0) ? u0 : (r == 1 ? u1 : u2);
                    xn += lin * ur + ang * (uw[0] * u0 + uw[1] * u1 + uw[2] * u2);
                }
            }
        }
        if (act) w.X[k * 12 + i] = xn;
        xi = xn;
        __syncwarp(hmask);
    }

    if (!alive) return valid ? 0 : -1;

    // ---- feasibility of the unconstrained minimiser
    double mv = -1e300;
    for (int e = hl; e < 4 * N; e += 16) {
        const int k = e >> 2, leg = e & 3;
        if (!mbit(mk0, mk1, mk2, N, leg, k)) continue;
        const double* f = w.U + 12 * k + 3 * leg;
        mv = fmax(mv, fmax(p.fz_min - f[2], fmax(fabs(f[0]), fabs(f[1])) - p.mu * f[2]));
    }
    mv = max16(hmask, mv);
    if (!(mv <= 1e-9)) return 0;

    // ---- co-states, stationarity, objective (first principles)
    double part = 0.0, rd = 0.0;
    for (int idx = hl; idx < 12 * N; idx += 16) {
        const int r = idx % 12;
        const double xr = w.XR[idx], dd = w.X[idx] - xr;
        part += p.Q[r] * (dd * dd - xr * xr);
    }
    double nui = 0.0;
    for (int k = N - 1; k >= 0; --k) {
        double s = -2.0 * Qi * (w.X[k * 12 + i] - w.XR[k * 12 + i]);
        if (k < N - 1) s += nui + ca * w.v0[ra] + cb * w.v0[rb];      // A' nu_{k+1}
        __syncwarp(hmask);
        nui = s;
        if (act) { w.v0[i] = s; w.NU[k * 12 + i] = s; }
        __syncwarp(hmask);
    }
    __syncwarp(hmask);
    double* ybox = w.XR;                       // the reference is no longer needed: stage the box multipliers there
    for (int e = hl; e < 4 * N; e += 16) {
        const int k = e >> 2, leg = e & 3;
        const int st = mbit(mk0, mk1, mk2, N, leg, k);
        const double r0 = w.RF[(leg * 3 + 0) * N + k], r1 = w.RF[(leg * 3 + 1) * N + k], r2 = w.RF[(leg * 3 + 2) * N + k];
        const double sk[9] = {0.0, -r2, r1, r2, 0.0, -r0, -r1, r0, 0.0};
        const double* Ii = w.dynv + 2;
        double Wm[9], Um[9];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int j = 0; j < 3; ++j) Wm[a * 3 + j] = Ii[a * 3] * sk[j] + Ii[a * 3 + 1] * sk[3 + j] + Ii[a * 3 + 2] * sk[6 + j];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            Um[j] = d.cy * Wm[j] + d.sy * Wm[3 + j];
            Um[3 + j] = -d.sy * Wm[j] + d.cy * Wm[3 + j];
            Um[6 + j] = Wm[6 + j];
        }
        const double* nu = w.NU + k * 12;
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
            double s = d.h * d.minv * nu[cc] + d.dt * d.minv * nu[6 + cc];
#pragma unroll
            for (int q = 0; q < 3; ++q) s += d.h * Um[q * 3 + cc] * nu[3 + q] + d.dt * Wm[q * 3 + cc] * nu[9 + q];
            const int idx = 12 * k + 3 * leg + cc;
            if (st) {
                const double Rv = p.R[3 * leg + cc], f = w.U[idx];
                rd = fmax(rd, fabs(2.0 * Rv * f - s));
                part += Rv * f * f;
                ybox[idx] = (cc == 2) ? -0.0 : 0.0;
            } else {
                ybox[idx] = s;
            }
        }
    }
    rd = max16(hmask, rd);
    const double obj = sum16(hmask, part);
    if (!(rd <= 1e-6)) return 0;
    __syncwarp(hmask);

    // ---- outputs
    for (int idx = hl; idx < 12 * N; idx += 16) { o.u[idx] = w.U[idx]; o.y[idx] = ybox[idx]; }
    for (int idx = hl; idx < 16 * N; idx += 16) o.y[12 * N + idx] = 0.0;
    if (o.X) { for (int idx = hl; idx < 12 * N; idx += 16) o.X[idx] = w.X[idx]; }
    if (o.nu) { for (int idx = hl; idx < 12 * N; idx += 16) o.nu[idx] = w.NU[idx]; }
    if (hl == 0) {
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

}  // namespace ric2
}  // namespace cmpc
#endif
