// cmpc_emul.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the kernel source (csrc/cmpc_core.cuh) for the host with a ONE-THREAD CTA, so that the
// CPU test-suite can check the *kernel logic* (indexing, recursions, active-set / ADMM control
// flow) against the oracle without a GPU.  It is not linked into libcmpc.so, it is not importable
// from the package, and nothing in the product path can reach it.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../convex-mpc-unitree-go2_b200/csrc/cmpc_core.cuh"
#include "../../convex-mpc-unitree-go2_b200/csrc/cmpc_fast.cuh"
#include "../../convex-mpc-unitree-go2_b200/csrc/cmpc_riccati.cuh"
#include "../../convex-mpc-unitree-go2_b200/csrc/cmpc_traj.cuh"
#include "../../convex-mpc-unitree-go2_b200/csrc/cmpc_wrench.cuh"

using namespace cmpc;

extern "C" {

void emul_default_params(Params* p) {
    const double Q[12] = {1, 1, 50, 10, 20, 1, 2, 2, 1, 1, 1, 1};
    for (int i = 0; i < 12; ++i) { p->Q[i] = Q[i]; p->R[i] = 1e-5; }
    p->mu = 0.8; p->fz_min = 10.0; p->eps_abs = 1e-4; p->eps_rel = 1e-4; p->max_iter = 1000;
    p->rho0 = 1e-4; p->sigma = 1e-6; p->alpha = 1.6; p->mode = 1; p->polish = 0;
    p->check_termination = 10; p->adaptive_rho_interval = 25; p->pdas_max_iter = 16;
}

int emul_params_size() { return (int)sizeof(Params); }

void emul_contact_table(int B, int N, const double* t0, double dt, double gait_hz, double duty,
                        const double* off, uint64_t* mask) {
    const int W = (4 * N + 63) / 64;
    const double period = 1 / gait_hz;
    for (int b = 0; b < B; ++b) {
        for (int w = 0; w < W; ++w) mask[(size_t)b * W + w] = 0;
        for (int leg = 0; leg < 4; ++leg)
            for (int k = 0; k < N; ++k)
                if (stance_bit(t0[b], dt, k, period, off[leg], duty)) {
                    const int bit = leg * N + k;
                    mask[(size_t)b * W + (bit >> 6)] |= (1ull << (bit & 63));
                }
    }
}

static QpIn make_in(int b, int N, const double* Ad, const double* Bd, const double* gd, const double* x0,
                    const double* x_ref, const double* r_foot, const double* I_world, const double* mass,
                    double dt, const uint64_t* mask) {
    const int W = (4 * N + 63) / 64;
    QpIn in;
    in.Ad = Ad ? Ad + (size_t)b * 144 : nullptr;
    in.Bd = Bd ? Bd + (size_t)b * N * 144 : nullptr;
    in.gd = gd ? gd + (size_t)b * 12 : nullptr;
    in.x0 = x0 + (size_t)b * 12;
    in.x_ref = x_ref + (size_t)b * 12 * N;
    in.r_foot = r_foot ? r_foot + (size_t)b * 12 * N : nullptr;
    in.I_world = I_world ? I_world + (size_t)b * 9 : nullptr;
    in.mass = mass ? mass[b] : 1.0;
    in.dt = dt;
    in.mask = mask ? mask + (size_t)b * W : nullptr;
    in.N = N;
    return in;
}

int emul_build(const Params* p, int B, int N, const double* Ad, const double* Bd, const double* gd,
               const double* x0, const double* x_ref, const double* r_foot, const double* I_world,
               const double* mass, double dt, double* H, double* g) {
    const int nfmax = 4 * N, n = 12 * N;
    Ws w;
    std::vector<unsigned char> buf(ws_carve(w, reinterpret_cast<unsigned char*>(4096), N, nfmax, nullptr) + 64);
    ws_carve(w, buf.data(), N, nfmax, nullptr);
    Cta c{0, 1, 0};
    for (int b = 0; b < B; ++b) {
        QpIn in = make_in(b, N, Ad, Bd, gd, x0, x_ref, r_foot, I_world, mass, dt, nullptr);
        build_dense_one(c, *p, in, w, nfmax, H + (size_t)b * n * n, g + (size_t)b * n);
    }
    return 0;
}

int emul_solve(const Params* p, int B, int N, int nfmax, const double* Ad, const double* Bd, const double* gd,
               const double* x0, const double* x_ref, const double* r_foot, const double* I_world,
               const double* mass, double dt, const uint64_t* mask, int warm, double* u, double* y, double* rho,
               double* X, double* nu, int32_t* status, int32_t* iters, double* stats) {
    Ws w;
    std::vector<unsigned char> buf(ws_carve(w, reinterpret_cast<unsigned char*>(4096), N, nfmax, nullptr) + 64);
    ws_carve(w, buf.data(), N, nfmax, nullptr);
    Cta c{0, 1, 0};
    for (int b = 0; b < B; ++b) {
        QpIn in = make_in(b, N, Ad, Bd, gd, x0, x_ref, r_foot, I_world, mass, dt, mask);
        QpOut o;
        o.u = u + (size_t)b * 12 * N;
        o.y = y + (size_t)b * 28 * N;
        o.rho = rho ? rho + b : nullptr;
        o.X = X ? X + (size_t)b * 12 * N : nullptr;
        o.nu = nu ? nu + (size_t)b * 12 * N : nullptr;
        o.status = status + b;
        o.iters = iters + b;
        o.stats = stats + (size_t)b * NSTAT;
        solve_one(c, *p, in, o, w, nfmax, warm);
    }
    return 0;
}

// v2 (raw-input fast path): same outputs, block-packed factor, closed-form build
int emul_solve_fast(const Params* p, int B, int N, int nfmax, const double* x0, const double* x_ref,
                    const double* r_foot, const double* I_world, const double* mass, double dt,
                    const uint64_t* mask, int warm, double* u, double* y, double* rho, double* X, double* nu,
                    int32_t* status, int32_t* iters, double* stats) {
    fast::WsF w;
    std::vector<unsigned char> buf(fast::ws_carve_fast(w, reinterpret_cast<unsigned char*>(4096), N, nfmax, nullptr) + 64);
    fast::ws_carve_fast(w, buf.data(), N, nfmax, nullptr);
    fast::Cx c = fast::make_cx(0, 1);
    fast::init_tables(c, w);
    for (int b = 0; b < B; ++b) {
        QpIn in = make_in(b, N, nullptr, nullptr, nullptr, x0, x_ref, r_foot, I_world, mass, dt, mask);
        QpOut o;
        o.u = u + (size_t)b * 12 * N;
        o.y = y + (size_t)b * 28 * N;
        o.rho = rho ? rho + b : nullptr;
        o.X = X ? X + (size_t)b * 12 * N : nullptr;
        o.nu = nu ? nu + (size_t)b * 12 * N : nullptr;
        o.status = status + b;
        o.iters = iters + b;
        o.stats = stats + (size_t)b * NSTAT;
        fast::solve_one_fast(c, *p, in, o, w, nfmax, warm);
    }
    return 0;
}

// Riccati pre-pass (csrc/cmpc_riccati.cuh) with a one-thread "warp": done[b] = 1 where the robot was finished
// by the pre-pass (outputs written), 0 where it would go on to the condensed path (outputs untouched).
int emul_riccati(const Params* p, int B, int N, int nfmax, const double* x0, const double* x_ref,
                 const double* r_foot, const double* I_world, const double* mass, double dt,
                 const uint64_t* mask, int warm, double* u, double* y, double* rho, double* X, double* nu,
                 int32_t* status, int32_t* iters, double* stats, int32_t* done) {
    ric::WsR w;
    std::vector<unsigned char> buf(ric::ws_carve_ric(w, reinterpret_cast<unsigned char*>(4096), N) + 64);
    ric::ws_carve_ric(w, buf.data(), N);
    std::vector<double> gains(ric::gain_doubles(nfmax));
    Cta c{0, 1, 1};
    for (int b = 0; b < B; ++b) {
        QpIn in = make_in(b, N, nullptr, nullptr, nullptr, x0, x_ref, r_foot, I_world, mass, dt, mask);
        QpOut o;
        o.u = u + (size_t)b * 12 * N;
        o.y = y + (size_t)b * 28 * N;
        o.rho = rho ? rho + b : nullptr;
        o.X = X ? X + (size_t)b * 12 * N : nullptr;
        o.nu = nu ? nu + (size_t)b * 12 * N : nullptr;
        o.status = status + b;
        o.iters = iters + b;
        o.stats = stats + (size_t)b * NSTAT;
        done[b] = ric::riccati_one(c, *p, in, o, w, nfmax, warm, gains.data());
    }
    return 0;
}

// ComTraj.generate_traj (csrc/cmpc_traj.cuh), one (robot, leg) at a time
int emul_generate_traj(int N, int B, const double* x0, const double* R_wb, const double* lever, const double* cmd,
                       const double* t0, double dt, double gait_hz, double duty, const double* off, const double* hip,
                       const double* pos_des_in, double* pos_des_out, double* x_ref, double* r_foot) {
    const double period = 1 / gait_hz;
    for (int b = 0; b < B; ++b) {
        double pin[3] = {pos_des_in[3 * b], pos_des_in[3 * b + 1], pos_des_in[3 * b + 2]};
        for (int leg = 0; leg < 4; ++leg)
            traj::generate_leg(N, leg, x0 + (size_t)b * 12, R_wb + (size_t)b * 9, lever + (size_t)b * 12 + 3 * leg,
                               cmd + (size_t)b * 4, t0[b], dt, period, duty, off[leg], hip + 3 * leg, pin,
                               pos_des_out + (size_t)b * 3, x_ref + (size_t)b * 12 * N, r_foot + ((size_t)b * 4 + leg) * 3 * N);
    }
    return 0;
}

int emul_srb_step(int N, int B, const double* x, const double* u, const double* x_ref, const double* r_foot,
                  const double* I_world, const double* mass, double T, const double* I_body, const double* so,
                  double* x_out, double* R_wb, double* I_out, double* lever) {
    for (int b = 0; b < B; ++b)
        traj::srb_step_one(N, x + (size_t)b * 12, u + (size_t)b * 12 * N, x_ref + (size_t)b * 12 * N, r_foot + (size_t)b * 12 * N,
                           I_world + (size_t)b * 9, mass[b], T, I_body, so, x_out + (size_t)b * 12, R_wb + (size_t)b * 9,
                           I_out + (size_t)b * 9, lever + (size_t)b * 12);
    return 0;
}

// Wrench-space projected Riccati + PDAS (csrc/cmpc_wrench.cuh): the four threads of a quad run one after the other
// between synchronisation points.  done[b] = 1 where the robot was finished here, sweeps[b] = Riccati sweeps used.
int emul_wrench(const Params* p, int B, int N, int nfmax, const double* x0, const double* x_ref, const double* r_foot,
                const double* I_world, const double* mass, double dt, const uint64_t* mask, int warm, double* u, double* y,
                double* rho, double* X, double* nu, int32_t* status, int32_t* iters, double* stats, int32_t* done,
                int32_t* sweeps) {
    std::vector<unsigned char> buf(wr::robot_bytes(N) + 64);
    wr::Sh* sh = reinterpret_cast<wr::Sh*>((reinterpret_cast<uintptr_t>(buf.data()) + 15) & ~(uintptr_t)15);
    std::vector<wr::D2> gains((size_t)N * wr::GAIN_D2 * 4);
    wr::Tab tb;
    for (int i = 0; i < 16; ++i) wr::fill_tab(tb, *p, i);
    wr::TS ts[4];
    std::vector<double> xs((size_t)B * 12 * N);
    wr::Bat bt{x0, x_ref, r_foot, I_world, mass, mask, u, y, rho, X ? X : xs.data(), nu, stats, status, iters, dt, N, (4 * N + 63) / 64};
    for (int b = 0; b < B; ++b) {
        wr::Env e;
        e.p = p; e.tb = &tb; e.bt = &bt; e.b = b;
        e.gains = gains.data();      // thread q uses gains[... * 4 + q]
        e.gstride = 4;
        e.dt = dt; e.h = dt * dt / 2.0;
        int sw = 0;
        done[b] = wr::solve_robot(0, ts, sh, e, nfmax, warm, &sw);
        sweeps[b] = sw;
    }
    return 0;
}

int emul_leg_jacobian(int B, const double* q, const double* R_wb, const double* link, double* J, double* p_body) {
    for (int b = 0; b < B; ++b)
        for (int leg = 0; leg < 4; ++leg)
            traj::leg_jacobian(q + (size_t)b * 12 + 3 * leg, R_wb + (size_t)b * 9, (leg & 1) ? -1.0 : 1.0, link[0], link[1], link[2],
                               J + ((size_t)b * 4 + leg) * 9, p_body ? p_body + ((size_t)b * 4 + leg) * 3 : nullptr);
    return 0;
}

size_t emul_ws_bytes_fast(int N, int nfmax) {
    fast::WsF w;
    return fast::ws_carve_fast(w, reinterpret_cast<unsigned char*>(4096), N, nfmax, nullptr);
}

size_t emul_ws_bytes(int N, int nfmax) {
    Ws w;
    return ws_carve(w, reinterpret_cast<unsigned char*>(4096), N, nfmax, nullptr);
}

}  // extern "C"
