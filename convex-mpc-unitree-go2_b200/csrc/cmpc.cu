// cmpc.cu -- CUDA kernels (sm_100a) and the C-ABI of include/cmpc.h.
//
// Kernels:
//   contact_table_kernel   gait.py:26-37                      one thread per robot
//   dynamics_kernel        com_trajectory.py:221-286          one thread per (robot, step)
//   build_kernel           centroidal_mpc.py:235-303 (condensed)   one CTA per robot (diagnostic)
//   solve_kernel           centroidal_mpc.py:69-120           one CTA per robot, everything in smem
//   solve_fast_kernel      the same from raw inputs (closed-form build, block-packed DMMA Cholesky)
//   riccati_kernel         nominal pre-pass of the fast path: one warp per robot (cmpc_riccati.cuh)
//   fp64_peak_kernel / smem_peak_kernel    roofline denominators measured on the box
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/cmpc.h"
#include "cmpc_core.cuh"
#include "cmpc_fast.cuh"
#include "cmpc_riccati.cuh"
#include "cmpc_riccati2.cuh"
#include "cmpc_traj.cuh"
#include "cmpc_wrench.cuh"

using namespace cmpc;

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};

int fail(const std::string& m) { g_err = m; return -1; }

#define CU_TRY(x)                                                                         \
    do {                                                                                  \
        cudaError_t e_ = (x);                                                             \
        if (e_ != cudaSuccess)                                                            \
            return fail(std::string(#x) + ": " + cudaGetErrorString(e_));                 \
    } while (0)

constexpr int kThreads = 256;

struct BatchIn {
    const double* Ad; const double* Bd; const double* gd;
    const double* x0; const double* x_ref; const double* r_foot; const double* I_world; const double* mass;
    double dt;
    const uint64_t* mask;
    int N, W;
};

struct BatchOut {
    double* u; double* y; double* rho; double* X; double* nu;
    int32_t* status; int32_t* iters; double* stats;
};

__device__ __forceinline__ QpIn qp_in(const BatchIn& bi, int b) {
    QpIn in;
    const int N = bi.N;
    in.Ad = bi.Ad ? bi.Ad + (size_t)b * 144 : nullptr;
    in.Bd = bi.Bd ? bi.Bd + (size_t)b * N * 144 : nullptr;
    in.gd = bi.gd ? bi.gd + (size_t)b * 12 : nullptr;
    in.x0 = bi.x0 + (size_t)b * 12;
    in.x_ref = bi.x_ref + (size_t)b * 12 * N;
    in.r_foot = bi.r_foot ? bi.r_foot + (size_t)b * 12 * N : nullptr;
    in.I_world = bi.I_world ? bi.I_world + (size_t)b * 9 : nullptr;
    in.mass = bi.mass ? bi.mass[b] : 1.0;
    in.dt = bi.dt;
    in.mask = bi.mask ? bi.mask + (size_t)b * bi.W : nullptr;
    in.N = N;
    return in;
}

__device__ __forceinline__ QpOut qp_out(const BatchOut& bo, int b, int N) {
    QpOut o;
    o.u = bo.u + (size_t)b * 12 * N;
    o.y = bo.y + (size_t)b * 28 * N;
    o.rho = bo.rho ? bo.rho + b : nullptr;
    o.X = bo.X ? bo.X + (size_t)b * 12 * N : nullptr;
    o.nu = bo.nu ? bo.nu + (size_t)b * 12 * N : nullptr;
    o.status = bo.status + b;
    o.iters = bo.iters + b;
    o.stats = bo.stats + (size_t)b * NSTAT;
    return o;
}

// ---------------------------------------------------------------------------------------------
__global__ void contact_table_kernel(int B, int N, int W, const double* __restrict__ t0, double dt, double period,
                                     double duty, double o0, double o1, double o2, double o3,
                                     uint64_t* __restrict__ mask) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double off[4] = {o0, o1, o2, o3};
    const double t = t0[b];
    uint64_t words[3] = {0, 0, 0};
    for (int leg = 0; leg < 4; ++leg)
        for (int k = 0; k < N; ++k) {
            const int bit = leg * N + k;
            if (stance_bit(t, dt, k, period, off[leg], duty)) words[bit >> 6] |= (1ull << (bit & 63));
        }
    for (int wv = 0; wv < W; ++wv) mask[(size_t)b * W + wv] = words[wv];
}

__global__ void pack_contact_kernel(int B, int N, int W, const int32_t* __restrict__ table, uint64_t* __restrict__ mask) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    uint64_t words[3] = {0, 0, 0};
    for (int bit = 0; bit < 4 * N; ++bit)
        if (table[(size_t)b * 4 * N + bit] != 0) words[bit >> 6] |= (1ull << (bit & 63));
    for (int wv = 0; wv < W; ++wv) mask[(size_t)b * W + wv] = words[wv];
}

__global__ void dynamics_kernel(int B, int N, const double* __restrict__ x_ref, const double* __restrict__ r_foot,
                                const double* __restrict__ I_world, const double* __restrict__ mass, double dt,
                                double* __restrict__ Ad, double* __restrict__ Bd, double* __restrict__ gd) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * N) return;
    const int b = idx / N, k = idx - b * N;
    DynCommon d;
    dyn_common(d, x_ref + (size_t)b * 12 * N, N, I_world + (size_t)b * 9, mass[b], dt);
    if (k == 0) {
        for (int i = 0; i < 144; ++i) Ad[(size_t)b * 144 + i] = dyn_Ad(d, i / 12, i % 12);
        double g[12];
        dyn_gd(d, g);
        for (int i = 0; i < 12; ++i) gd[(size_t)b * 12 + i] = g[i];
    }
    double* out = Bd + ((size_t)b * N + k) * 144;
    for (int leg = 0; leg < 4; ++leg) {
        double r[3];
        for (int a = 0; a < 3; ++a) r[a] = r_foot[(size_t)b * 12 * N + (size_t)(leg * 3 + a) * N + k];
        for (int comp = 0; comp < 3; ++comp) {
            double col[12];
            dyn_Bd_col(d, r, comp, col);
            for (int a = 0; a < 12; ++a) out[a * 12 + 3 * leg + comp] = col[a];
        }
    }
}

__global__ void __launch_bounds__(kThreads)
build_kernel(Params p, BatchIn bi, int B, int nfmax, double* __restrict__ H, double* __restrict__ g,
             double* hp_scratch, size_t hp_stride) {
    extern __shared__ __align__(16) unsigned char smem[];
    Ws w;
    ws_carve(w, smem, bi.N, nfmax, hp_scratch ? hp_scratch + (size_t)blockIdx.x * hp_stride : nullptr);
    Cta c;
    c.tid = threadIdx.x; c.nt = blockDim.x; c.warp = 0;
    const int n = 12 * bi.N;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        QpIn in = qp_in(bi, b);
        build_dense_one(c, p, in, w, nfmax, H + (size_t)b * n * n, g + (size_t)b * n);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kThreads)
solve_kernel(Params p, BatchIn bi, BatchOut bo, int B, int nfmax, int warm, double* hp_scratch, size_t hp_stride) {
    extern __shared__ __align__(16) unsigned char smem[];
    Ws w;
    ws_carve(w, smem, bi.N, nfmax, hp_scratch ? hp_scratch + (size_t)blockIdx.x * hp_stride : nullptr);
    Cta c;
    c.tid = threadIdx.x; c.nt = blockDim.x; c.warp = 0;
    const int N = bi.N;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        QpIn in = qp_in(bi, b);
        QpOut o;
        o.u = bo.u + (size_t)b * 12 * N;
        o.y = bo.y + (size_t)b * 28 * N;
        o.rho = bo.rho ? bo.rho + b : nullptr;
        o.X = bo.X ? bo.X + (size_t)b * 12 * N : nullptr;
        o.nu = bo.nu ? bo.nu + (size_t)b * 12 * N : nullptr;
        o.status = bo.status + b;
        o.iters = bo.iters + b;
        o.stats = bo.stats + (size_t)b * NSTAT;
        solve_one(c, p, in, o, w, nfmax, warm);
        __syncthreads();
    }
}

// v2 fast path (raw inputs): persistent-style grid-stride loop, one CTA per robot at a time,
// two CTAs resident per SM when the workspace allows it.
template <bool kExternal>
__global__ void __launch_bounds__(kThreads, 2)
solve_fast_kernel(Params p, BatchIn bi, BatchOut bo, int B, int nfmax, int warm, double* hb_scratch, size_t hb_stride,
                  const int* __restrict__ worklist, const int* __restrict__ wl_count, double* yg_scratch, size_t yg_stride) {
    extern __shared__ __align__(16) unsigned char smem[];
    fast::WsF w;
    if (kExternal) fast::ws_carve_fast<2>(w, smem, bi.N, nfmax, hb_scratch + (size_t)blockIdx.x * hb_stride);
    else fast::ws_carve_fast<1>(w, smem, bi.N, nfmax, nullptr);
    if (yg_scratch) w.Yg = yg_scratch + (size_t)blockIdx.x * yg_stride;      // rows of large working sets (L2-resident)
    const fast::Cx c = fast::make_cx(threadIdx.x, blockDim.x);
    fast::init_tables(c, w);
    PHASE_KERNEL_BEGIN();
    const int N = bi.N;
    // with a work-list (robots the Riccati pre-pass could not finish) the loop runs over its entries
    const int count = worklist ? *wl_count : B;
    if (worklist) {
        // work-list entries differ a lot in cost (active-set iterations): CTAs take the next entry from a cursor
        // (wl_count[1], zeroed with the counter) instead of a fixed stride
        __shared__ int next_it;
        for (;;) {
            if (threadIdx.x == 0) next_it = atomicAdd(const_cast<int*>(wl_count) + 1, 1);
            __syncthreads();
            const int it = next_it;
            __syncthreads();
            if (it >= count) break;
            const int b = worklist[it];
            QpIn in = qp_in(bi, b);
            QpOut o = qp_out(bo, b, N);
            fast::solve_one_fast(c, p, in, o, w, nfmax, warm);
            __syncthreads();
        }
    } else {
        for (int it = blockIdx.x; it < count; it += gridDim.x) {
            QpIn in = qp_in(bi, it);
            QpOut o = qp_out(bo, it, N);
            fast::solve_one_fast(c, p, in, o, w, nfmax, warm);
            __syncthreads();
        }
    }
    PHASE_KERNEL_END();
}

// Nominal pre-pass of the fast path: one warp per robot (cmpc_riccati.cuh); robots it cannot finish are
// appended to the work-list of the condensed kernel above.
constexpr int kRicThreads = 128;
__global__ void __launch_bounds__(kRicThreads, 3)
riccati_kernel(Params p, BatchIn bi, BatchOut bo, int B, int nfmax, int warm, double* __restrict__ gains,
               size_t gain_stride, int* __restrict__ worklist, int* __restrict__ wl_count, size_t smem_per_warp) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    ric::WsR w;
    ric::ws_carve_ric(w, smem + (size_t)wid * smem_per_warp, bi.N);
    Cta c;
    c.tid = lane; c.nt = 32; c.warp = 1;
    const int gw = blockIdx.x * wpb + wid, nw = gridDim.x * wpb;
    double* g = gains + (size_t)gw * gain_stride;
    const int N = bi.N;
    for (int b = gw; b < B; b += nw) {
        QpIn in = qp_in(bi, b);
        QpOut o = qp_out(bo, b, N);
        const int done = ric::riccati_one(c, p, in, o, w, nfmax, warm, g);
        if (!done && lane == 0) worklist[atomicAdd(wl_count, 1)] = b;
        __syncwarp();
    }
}

// Pre-pass, version 2 (cmpc_riccati2.cuh): two robots per warp, register-resident rows.
constexpr int kRic2Threads = 64;
__global__ void __launch_bounds__(kRic2Threads)
riccati2_kernel(Params p, BatchIn bi, BatchOut bo, int B, int nfmax, int warm, double* __restrict__ gains,
                size_t gain_stride, int* __restrict__ worklist, int* __restrict__ wl_count, size_t smem_per_half) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, half = lane >> 4, hl = lane & 15;
    const unsigned hmask = 0xFFFFu << (lane & 16);
    const int hpb = (blockDim.x >> 5) * 2;                       // robots in flight per block
    const int slot = wid * 2 + half;
    ric2::WsH w;
    ric2::carve_half(w, smem + (size_t)slot * smem_per_half, bi.N);
    const int gh = blockIdx.x * hpb + slot, nh = gridDim.x * hpb;
    double* g = gains + (size_t)gh * gain_stride;
    const int N = bi.N;
    for (int b = gh; b < B; b += nh) {
        QpIn in = qp_in(bi, b);
        QpOut o = qp_out(bo, b, N);
        const int done = ric2::riccati_half<false>(hmask, hl, true, p, in, o, w, nfmax, warm, g);
        if (!done && hl == 0) worklist[atomicAdd(wl_count, 1)] = b;
        __syncwarp(hmask);
    }
}

// The same sweep with the sixteen robots of a 256-thread CTA in lock-step (one CTA barrier per stage).
constexpr int kRic2LsThreads = 256;
__global__ void __launch_bounds__(kRic2LsThreads, 1)
riccati2_lockstep_kernel(Params p, BatchIn bi, BatchOut bo, int B, int nfmax, int warm, double* __restrict__ gains,
                         size_t gain_stride, int* __restrict__ worklist, int* __restrict__ wl_count, size_t smem_per_half) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, half = lane >> 4, hl = lane & 15;
    const unsigned hmask = 0xFFFFu << (lane & 16);
    const int hpb = (blockDim.x >> 5) * 2;
    const int slot = wid * 2 + half;
    ric2::WsH w;
    ric2::carve_half(w, smem + (size_t)slot * smem_per_half, bi.N);
    const int gh = blockIdx.x * hpb + slot, nh = gridDim.x * hpb;
    double* g = gains + (size_t)gh * gain_stride;
    const int N = bi.N;
    const int rounds = (B + nh - 1) / nh;
    for (int it = 0; it < rounds; ++it) {
        const int b = gh + it * nh;
        const bool valid = b < B;
        QpIn in = qp_in(bi, valid ? b : 0);
        QpOut o = qp_out(bo, valid ? b : 0, N);
        const int done = ric2::riccati_half<true>(hmask, hl, valid, p, in, o, w, nfmax, warm, g);
        if (done == 0 && hl == 0) worklist[atomicAdd(wl_count, 1)] = b;
        __syncwarp(hmask);
    }
}

// Wrench-space projected Riccati + primal-dual active set (cmpc_wrench.cuh): four threads per robot, eight robots per
// warp, persistent CTAs that take robots from a device cursor.  A quad keeps its robot until the working set has
// settled (every sweep has the same length, so the eight robots of a warp stay in step whatever their iteration
// counts are) and then takes the next one.  Robots it cannot finish (cycling working sets, budget) are appended to
// the work-list of the condensed kernel.  ctl: [0] work-list length, [1] cursor of the condensed kernel, [2] robot cursor.
constexpr int kWrThreads = 128;
__global__ void __launch_bounds__(kWrThreads, 2)
wrench_pdas_kernel(Params p, wr::Bat bt, int B, int nfmax, int warm, wr::D2* __restrict__ gains,
                   int* __restrict__ worklist, int* __restrict__ ctl, size_t robot_bytes) {
    extern __shared__ __align__(16) unsigned char smem[];
    wr::Tab* tb = reinterpret_cast<wr::Tab*>(smem);
    if (threadIdx.x < 16) wr::fill_tab(*tb, p, threadIdx.x);
    __syncthreads();
    const int lane = threadIdx.x & 31, qlane = threadIdx.x & 3, rslot = threadIdx.x >> 2;
    const size_t tab_bytes = (sizeof(wr::Tab) + 15) & ~(size_t)15;
    wr::Sh* sh = reinterpret_cast<wr::Sh*>(smem + tab_bytes + (size_t)rslot * robot_bytes);
    unsigned char* codes = wr::codes_of(sh);
    const int N = bt.N;
    wr::TS ts[1];
    wr::Env e;
    e.p = &p; e.tb = tb; e.bt = &bt; e.b = 0;
    e.gains = gains + (size_t)blockIdx.x * N * wr::GAIN_D2 * blockDim.x + threadIdx.x;
    e.gstride = blockDim.x;
    e.dt = bt.dt; e.h = bt.dt * bt.dt / 2.0;
    const int max_it = p.pdas_max_iter;
    int b = -1, it = 0, nst = 0;
    bool exhausted = false;
    wr::Policy pl = wr::policy_init();
    for (;;) {
        const bool need = (b < 0) && !exhausted;
        int nb = -1;
        if (need && qlane == 0) nb = atomicAdd(ctl + 2, 1);
        nb = __shfl_sync(0xffffffffu, nb, lane & ~3);
        bool fresh = false;
        if (need) {
            if (nb < B) { b = nb; it = 0; pl = wr::policy_init(); fresh = true; e.b = b; }
            else exhausted = true;
        }
        if (!__any_sync(0xffffffffu, b >= 0)) break;
        const int n2 = wr::init_robot(qlane, fresh, ts, sh, e, warm);
        if (fresh) nst = n2;
        bool fail = fresh && (nst == 0 || nst > nfmax);          // nothing to optimise / over the stance bound: condensed kernel reports
        const bool valid = (b >= 0) && !fail;
        const unsigned char* cur = codes + (size_t)(it % 3) * 4 * N;
        unsigned char* next = codes + (size_t)((it + 1) % 3) * 4 * N;
        const double pmin = wr::backward_sweep(qlane, ts, sh, e, cur);
        const int fl = wr::forward_sweep(qlane, valid, ts, sh, e, cur, next, it, wr::policy_damp(pl, it));
        if (valid && !(pmin > 0.0)) fail = true;
        bool single = false;
        const int ps = wr::policy_step(pl, fl, it, max_it, N <= 16, single);       // working-set policy (block updates, damping, cycle breakers)
        const bool conv = valid && !fail && ps == 1;
        if (valid && !conv && (ps == 2 || it + 1 >= max_it + wr::kBlockExtra + wr::kSingleMax)) fail = true;
        wr::mark_pending(qlane, conv, e, nst, it + 1);      // the certificate kernel takes it from here
        wr::single_step(qlane, single && valid && !conv && !fail, ts, sh, cur, next, N);
        if (fail && qlane == 0) worklist[atomicAdd(ctl, 1)] = b;
        if ((conv && !fail) || fail) b = -1; else if (valid) ++it;
    }
}

// Certificates and remaining outputs of the robots the sweep kernel settled (cmpc_wrench.cuh, certify_group): sixteen lanes
// per robot, nothing carried over from the sweeps -- everything is recomputed from X, u, y.  Robots whose certificate does not
// hold join the work-list of the condensed kernel.
__global__ void __launch_bounds__(128, 3)
wrench_certificate_kernel(Params p, wr::Bat bt, int B, int warm, int* __restrict__ worklist, int* __restrict__ ctl) {
    const int lane = threadIdx.x & 31, gl = lane & 15;
    const unsigned gmask = 0xFFFFu << (lane & 16);
    const int per = blockDim.x >> 4;
    for (int base = blockIdx.x * per; base < B; base += gridDim.x * per) {
        const int b = base + (threadIdx.x >> 4);
        if (b >= B || bt.status[b] != wr::ST_PENDING) continue;          // uniform inside the group of sixteen
        const int ok = wr::certify_group(gl, gmask, p, bt, b, warm);
        if (!ok && gl == 0) worklist[atomicAdd(ctl, 1)] = b;
    }
}

// Constants of every robot ahead of the sweeps (cy, sy, 1/m, Iinv: what wr::robot_consts forms): one thread per robot, so that
// the sweep kernel -- four threads per robot, bound by instruction fetch -- carries neither sincos nor N serial loads.
__global__ void robot_consts_kernel(int B, int N, const double* __restrict__ x_ref, const double* __restrict__ I_world,
                                    const double* __restrict__ mass, double dt, double* __restrict__ cst) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    DynCommon dc;
    dyn_common(dc, x_ref + (size_t)b * 12 * N, N, I_world + (size_t)b * 9, mass[b], dt);
    double* o = cst + (size_t)b * 12;
    o[0] = dc.cy; o[1] = dc.sy; o[2] = dc.minv;
    for (int i = 0; i < 9; ++i) o[3 + i] = dc.Iinv[i];
}

// U_opt[:, 0] of every robot into a compact (B,12) buffer (cmpc_cycle_host, first_step_only): a strided device-to-host copy of
// 96-byte rows is far slower than this kernel plus one contiguous copy.
__global__ void first_step_kernel(int B, int N, const double* __restrict__ u, double* __restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 12 * B) return;
    const int b = e / 12, c = e - 12 * b;
    out[e] = u[(size_t)b * 12 * N + c];
}

// A head start for the condensed kernel (see cmpc_solve): one thread that lets `cycles` clock ticks pass.
__global__ void head_start_kernel(long long cycles) {
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) { }
}

// Batched ComTraj.generate_traj (cmpc_traj.cuh): one thread per (robot, leg).
__global__ void generate_traj_kernel(int B, int N, const double* __restrict__ x0, const double* __restrict__ R_wb,
                                     const double* __restrict__ lever, const double* __restrict__ cmd,
                                     const double* __restrict__ t0, double dt, double period, double duty,
                                     double o0, double o1, double o2, double o3, double h0, double h1, double h2,
                                     double h3, double h4, double h5, double h6, double h7, double h8, double h9,
                                     double h10, double h11, const double* pos_des_in, double* pos_des_out,
                                     double* __restrict__ x_ref, double* __restrict__ r_foot) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 4 * B) return;
    const int b = e >> 2, leg = e & 3;
    const double off = leg == 0 ? o0 : (leg == 1 ? o1 : (leg == 2 ? o2 : o3));
    const double hipv[12] = {h0, h1, h2, h3, h4, h5, h6, h7, h8, h9, h10, h11};
    double pin[3];
    for (int a = 0; a < 3; ++a) pin[a] = pos_des_in[(size_t)b * 3 + a];
    // every thread of a robot has read pos_des_in before lane (leg 0) may overwrite it in place:
    // the four threads of a robot sit in one warp
    __syncwarp();
    traj::generate_leg(N, leg, x0 + (size_t)b * 12, R_wb + (size_t)b * 9, lever + (size_t)b * 12 + 3 * leg,
                       cmd + (size_t)b * 4, t0[b], dt, period, duty, off, hipv + 3 * leg, pin,
                       pos_des_out ? pos_des_out + (size_t)b * 3 : nullptr, x_ref + (size_t)b * 12 * N,
                       r_foot + ((size_t)b * 4 + leg) * 3 * N);
}

// SRB closed-loop step (cmpc_traj.cuh): one thread per robot.
__global__ void srb_step_kernel(int B, int N, const double* __restrict__ x, const double* __restrict__ u,
                                const double* __restrict__ x_ref, const double* __restrict__ r_foot,
                                const double* __restrict__ I_world, const double* __restrict__ mass, double T,
                                double ib0, double ib1, double ib2, double s0, double s1, double s2, double s3,
                                double s4, double s5, double s6, double s7, double s8, double s9, double s10,
                                double s11, double* x_out, double* R_wb_out, double* I_out, double* lever_out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double ib[3] = {ib0, ib1, ib2};
    const double so[12] = {s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11};
    double xin[12], Iw[9];
    for (int i = 0; i < 12; ++i) xin[i] = x[(size_t)b * 12 + i];       // x_out / I_out may alias the inputs
    for (int i = 0; i < 9; ++i) Iw[i] = I_world[(size_t)b * 9 + i];
    traj::srb_step_one(N, xin, u + (size_t)b * 12 * N, x_ref + (size_t)b * 12 * N, r_foot + (size_t)b * 12 * N, Iw,
                       mass[b], T, ib, so, x_out + (size_t)b * 12, R_wb_out + (size_t)b * 9, I_out + (size_t)b * 9,
                       lever_out + (size_t)b * 12);
}

// Stance torque mapping (cmpc_traj.cuh): one thread per (robot, leg).
__global__ void stance_torque_kernel(int B, int N, const double* __restrict__ J, const double* __restrict__ u,
                                     const double* __restrict__ t_now, double period, double duty, double o0, double o1,
                                     double o2, double o3, double tau_max, double* __restrict__ tau, int32_t* __restrict__ mask_now) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 4 * B) return;
    const int b = e >> 2, leg = e & 3;
    const double off = leg == 0 ? o0 : (leg == 1 ? o1 : (leg == 2 ? o2 : o3));
    const int st = traj::stance_bit_now(t_now[b], 0.0, 0, period, off, duty);
    traj::stance_torque_leg(J + ((size_t)b * 4 + leg) * 9, u + (size_t)b * 12 * N + 3 * leg, st, tau_max,
                            tau + (size_t)b * 12 + 3 * leg);
    if (mask_now) mask_now[e] = st;
}

// Analytic foot Jacobians (cmpc_traj.cuh): one thread per (robot, leg); legs FL FR RL RR, left legs have side = +1.
__global__ void leg_jacobian_kernel(int B, const double* __restrict__ q, const double* __restrict__ R_wb, double l1, double l2,
                                    double l3, double* __restrict__ J, double* __restrict__ p_body) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 4 * B) return;
    const int b = e >> 2, leg = e & 3;
    traj::leg_jacobian(q + (size_t)b * 12 + 3 * leg, R_wb + (size_t)b * 9, (leg & 1) ? -1.0 : 1.0, l1, l2, l3,
                       J + (size_t)e * 9, p_body ? p_body + (size_t)e * 3 : nullptr);
}

// ---------------------------------------------------------------------------------------------
// Roofline denominators: dependent-free DFMA streams and shared-memory 8-byte reads.
// ---------------------------------------------------------------------------------------------
__global__ void fp64_peak_kernel(double* out, int iters) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, s = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, s); a1 = fma(a1, m, s); a2 = fma(a2, m, s); a3 = fma(a3, m, s);
        a4 = fma(a4, m, s); a5 = fma(a5, m, s); a6 = fma(a6, m, s); a7 = fma(a7, m, s);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void smem_peak_kernel(double* out, int iters) {
    extern __shared__ __align__(16) unsigned char smem[];
    double2* s = reinterpret_cast<double2*>(smem);
    const int nvec = 4096;   // 64 KB
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) s[i] = make_double2(i, 1.0);
    __syncthreads();
    double acc = 0.0;
    int idx = threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const double2 v = s[(idx + u * blockDim.x) & (nvec - 1)];
            acc += v.x + v.y;
        }
        idx = (idx + 8 * blockDim.x) & (nvec - 1);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// FP64 tensor-core (DMMA m8n8k4) stream: 8 independent accumulator pairs per warp.
__global__ void dmma_peak_kernel(double* out, int iters) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = 1.0 + i; }
    const double a = 1.0000001, b = 0.9999999;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// dependent DMMA chain: latency of one m8n8k4 in cycles (single warp)
__global__ void dmma_latency_kernel(double* out, long long* cycles, int iters) {
    double c0 = threadIdx.x, c1 = 1.0;
    const double a = 1.0000001, b = 0.9999999;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    const long long t1 = clock64();
    out[threadIdx.x] = c0 + c1;
    if (threadIdx.x == 0) *cycles = t1 - t0;
}

// single-warp dependent-issue latencies (cycles): 0 DFMA, 1 rsqrt(double), 2 double shuffle, 3 LDS.64
__global__ void latency_kernel(double* out, long long* cycles, int iters) {
    __shared__ double sm[64];
    double v = 1.0 + threadIdx.x * 1e-9;
    sm[threadIdx.x] = (double)((threadIdx.x * 7 + 1) & 31);
    sm[threadIdx.x + 32] = 0.0;
    __syncwarp();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) v = fma(v, 1.0000001, 1e-9);
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) v = rsqrt(v) + 1.5;
    t1 = clock64();
    if (threadIdx.x == 0) cycles[1] = t1 - t0;     // includes one DADD
    t0 = clock64();
    for (int i = 0; i < iters; ++i) v = __shfl_sync(0xffffffffu, v, (threadIdx.x + 1) & 31);
    t1 = clock64();
    if (threadIdx.x == 0) cycles[2] = t1 - t0;
    int idx = threadIdx.x;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) idx = (int)sm[idx];
    t1 = clock64();
    if (threadIdx.x == 0) cycles[3] = t1 - t0;     // LDS.64 + F2I
    out[threadIdx.x] = v + idx;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// Handle
// ---------------------------------------------------------------------------------------------
struct cmpc_handle {
    int N = 16, W = 1, max_batch = 0, device = 0;
    int nfmax = 64;        // largest number of stance foot-steps any robot may have
    int force_generic = 0; // diagnostics: route raw inputs through the generic kernel too
    int sm_count = 148;
    size_t smem_optin = 0, smem_per_sm = 0;
    Params p;
    int prepass = 4;       // pre-pass ahead of the condensed kernel (active-set mode, raw inputs): 0 off, 1-3 Riccati sweeps of round 1, 4 wrench-space PDAS
    int prepass_min_batch = 2048;   // CMPC_PREPASS_MIN_BATCH overrides (diagnostics)
    // Workspace.  Everything a solve needs beyond its arguments lives in one of four SLOTS (work-list, counters, gain
    // scratch, L2-resident scratch of the condensed kernel); consecutive calls rotate over the slots, so up to four
    // solves of one handle may be in flight on different streams.  Slots are sized by cmpc_reserve (or by the first
    // solve, for max_batch robots); nothing is allocated afterwards.
    struct Slot {
        int* worklist = nullptr; int* ctl = nullptr;          // ctl: [0] work-list length, [1] cursor of the condensed kernel, [2] robot cursor of the wrench kernel
        double* gains = nullptr; size_t gain_bytes = 0;       // Riccati gains (pre-pass 1-3) or wrench gains (pre-pass 4)
        double* yg = nullptr;                                 // rows of large working sets, per CTA of the condensed kernel
        double* hp = nullptr;                                 // block-packed factor when it does not fit shared memory
        double* xs = nullptr;                                 // predicted states (B,12N) when the caller does not ask for them
        double* cst = nullptr;                                // (B,12) robot constants of route 4
        // route 4: the condensed kernel works on the hand-overs of the sweep kernel on `aux` while the certificate kernel runs;
        // robots whose certificate fails go to a second list (worklist + wl_cap, ctl + 4), served after the join
        size_t wl_cap = 0;
        cudaStream_t aux = nullptr;
        cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    };
    Slot slot[4];
    std::atomic<unsigned> slot_next{0};
    // per-kernel timing of cmpc_solve (cmpc_set_profile): events before the pre-pass, between the two kernels and after
    int profile = 0;
    cudaEvent_t pev[4] = {nullptr, nullptr, nullptr, nullptr};      // before | after the sweep kernel | after the pre-pass | after the condensed kernel
    int pev_valid = 0;
    int reserved_batch = 0;         // robots the slots are sized for (0 = not reserved yet)
    int reserved_nfmax = 0;
    size_t yg_stride = 0, hp_stride = 0, hp_stride_generic = 0;
    size_t reserved_bytes = 0;
    // cudaFuncSetAttribute is issued only when the dynamic shared-memory size of a kernel changes
    size_t attr_fast[2] = {0, 0}, attr_generic = 0, attr_build = 0, attr_ric[3] = {0, 0, 0}, attr_wrench = 0, attr_cert = 0;
    // device-resident state for the *_host entry
    struct HostPath {
        bool ready = false;
        double *x0 = nullptr, *x_ref = nullptr, *r_foot = nullptr, *I_world = nullptr, *mass = nullptr, *t0 = nullptr;
        uint64_t* mask = nullptr;
        double *u = nullptr, *y = nullptr, *rho = nullptr, *stats = nullptr;
        int32_t *status = nullptr, *iters = nullptr;
        double *R_wb = nullptr, *lever = nullptr, *cmd = nullptr, *pos_des = nullptr, *u0 = nullptr;      // cmpc_cycle_host only
        cudaStream_t s[2] = {nullptr, nullptr};
    } hp;
};

namespace {

void default_params(Params& p) {
    const double Q[12] = {1, 1, 50, 10, 20, 1, 2, 2, 1, 1, 1, 1};   // centroidal_mpc.py:12
    for (int i = 0; i < 12; ++i) { p.Q[i] = Q[i]; p.R[i] = 1e-5; }  // centroidal_mpc.py:13
    p.mu = 0.8;            // centroidal_mpc.py:15
    p.fz_min = 10.0;       // centroidal_mpc.py:127
    p.eps_abs = 1e-4;      // centroidal_mpc.py:25-26 (OPTS)
    p.eps_rel = 1e-4;
    p.max_iter = 1000;     // :27
    p.rho0 = 1e-4;         // OSQP's 0.1 is 3 orders off for this problem
    p.sigma = 1e-6;
    p.alpha = 1.6;
    p.mode = CMPC_MODE_ACTIVE_SET;
    p.polish = 0;
    p.check_termination = 10;       // :31
    p.adaptive_rho_interval = 25;   // :32
    p.pdas_max_iter = 16;
}

size_t smem_needed(int N, int nfmax, bool hp_external) {
    Ws w;
    unsigned char* base = reinterpret_cast<unsigned char*>(static_cast<uintptr_t>(1 << 20));
    static double dummy;
    return ws_carve(w, base, N, nfmax, hp_external ? &dummy : nullptr);
}

// cudaFuncSetAttribute only when the size changes (the call is not free and is not capturable)
int set_smem(const void* kernel, size_t need, size_t* cached) {
    if (*cached == need) return 0;
    CU_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    *cached = need;
    return 0;
}

size_t smem_needed_fast(int N, int nfmax, bool hb_external) {
    fast::WsF w;
    unsigned char* base = reinterpret_cast<unsigned char*>(static_cast<uintptr_t>(1 << 20));
    static double dummy;
    return fast::ws_carve_fast(w, base, N, nfmax, hb_external ? &dummy : nullptr);
}

// per-CTA strides (doubles) of the global scratch the kernels may need, 0 if they do not need it
size_t hp_stride_generic(const cmpc_handle* h, int nfmax) {
    if (smem_needed(h->N, nfmax, false) <= h->smem_optin) return 0;
    const size_t nmax = 3 * (size_t)nfmax;
    return ((nmax + 1) * (nmax + 2) / 2 + 1) & ~(size_t)1;
}
size_t hp_stride_fast(const cmpc_handle* h, int nfmax) {
    if (smem_needed_fast(h->N, nfmax, false) <= h->smem_optin) return 0;
    const size_t nblk = (3 * (size_t)nfmax + 7) / 8;
    return nblk * (nblk + 1) / 2 * 64;
}
size_t yg_stride_of(const cmpc_handle* h, int nfmax) {
    const size_t npad = ((3 * (size_t)nfmax + 7) / 8) * 8;
    return (size_t)fast::kcap_fast(nfmax, h->N) * npad;
}

// wrench kernel: shared memory per CTA and the grid for B robots
size_t wrench_smem(const cmpc_handle* h) {
    return ((sizeof(wr::Tab) + 15) & ~(size_t)15) + (size_t)(kWrThreads / 4) * wr::robot_bytes(h->N);
}
int wrench_grid(const cmpc_handle* h, int B) {
    int per_sm = (int)(h->smem_per_sm / (wrench_smem(h) + 1024));
    if (per_sm > 2) per_sm = 2;                      // 255 registers x 128 threads x 2
    if (per_sm < 1) per_sm = 1;
    if (const char* e = getenv("CMPC_WRENCH_CTAS_PER_SM")) { const int v = atoi(e); if (v >= 1 && v < per_sm) per_sm = v; }   // diagnostics
    const int want = (B + kWrThreads / 4 - 1) / (kWrThreads / 4), cap = h->sm_count * per_sm;
    return want < cap ? want : cap;
}

struct SlotSizes { size_t worklist, ctl, gains, yg, hp, xs, total; };
SlotSizes slot_sizes(const cmpc_handle* h, int B) {
    SlotSizes z;
    const int ctas = h->sm_count * 2;
    z.worklist = (size_t)2 * B * sizeof(int);
    z.ctl = 8 * sizeof(int);
    const size_t g_ric = ric::gain_doubles(h->nfmax) * (size_t)h->sm_count * 32 * sizeof(double);    // up to 32 robots in flight per SM
    const size_t g_wr = (size_t)h->sm_count * 2 * h->N * wr::GAIN_D2 * kWrThreads * sizeof(wr::D2);
    z.gains = g_ric > g_wr ? g_ric : g_wr;
    z.yg = yg_stride_of(h, h->nfmax) * ctas * sizeof(double);
    size_t hs = hp_stride_fast(h, h->nfmax);
    const size_t hg = hp_stride_generic(h, h->nfmax), hb = hp_stride_generic(h, 4 * h->N);
    if (hg > hs) hs = hg;
    if (hb > hs) hs = hb;
    z.hp = hs * ctas * sizeof(double);
    z.xs = (size_t)B * 12 * h->N * sizeof(double);
    z.total = 4 * (z.worklist + z.ctl + z.gains + z.yg + z.hp + z.xs + (size_t)B * 12 * sizeof(double));
    return z;
}

void free_slots(cmpc_handle* h) {
    for (auto& q : h->slot) {
        void* ptrs[] = {q.worklist, q.ctl, q.gains, q.yg, q.hp, q.xs, q.cst};
        for (void* p : ptrs) if (p) cudaFree(p);
        if (q.aux) cudaStreamDestroy(q.aux);
        if (q.ev_fork) cudaEventDestroy(q.ev_fork);
        if (q.ev_join) cudaEventDestroy(q.ev_join);
        q = cmpc_handle::Slot();
    }
    h->reserved_batch = 0;
    h->reserved_bytes = 0;
}

// (Re)size the four slots for batches of up to B robots with the current stance bound.  Synchronises the device when
// it has to free; not to be called while a graph is being captured.
int reserve_slots(cmpc_handle* h, int B) {
    if (B <= h->reserved_batch && h->nfmax == h->reserved_nfmax) return 0;
    if (h->reserved_batch) { CU_TRY(cudaDeviceSynchronize()); free_slots(h); }
    const SlotSizes z = slot_sizes(h, B);
    for (auto& q : h->slot) {
        CU_TRY(cudaMalloc(&q.worklist, z.worklist));
        CU_TRY(cudaMalloc(&q.ctl, z.ctl));
        CU_TRY(cudaMalloc(&q.gains, z.gains));
        q.gain_bytes = z.gains;
        if (z.yg) CU_TRY(cudaMalloc(&q.yg, z.yg));
        if (z.hp) CU_TRY(cudaMalloc(&q.hp, z.hp));
        CU_TRY(cudaMalloc(&q.xs, z.xs));
        CU_TRY(cudaMalloc(&q.cst, (size_t)B * 12 * sizeof(double)));
        q.wl_cap = (size_t)B;
        CU_TRY(cudaStreamCreateWithFlags(&q.aux, cudaStreamNonBlocking));
        CU_TRY(cudaEventCreateWithFlags(&q.ev_fork, cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&q.ev_join, cudaEventDisableTiming));
    }
    h->reserved_batch = B;
    h->reserved_nfmax = h->nfmax;
    h->reserved_bytes = z.total;
    return 0;
}

int ensure_reserved(cmpc_handle* h, int B) {
    if (B <= h->reserved_batch && h->nfmax == h->reserved_nfmax) return 0;
    return reserve_slots(h, B > h->max_batch ? B : h->max_batch);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

const char* cmpc_last_error(void) { return g_err.c_str(); }
const char* cmpc_version(void) { return "cmpc-b200 0.1.0 (sm_100a)"; }
long long cmpc_launch_count(void) { return g_launches.load(); }

int cmpc_create(int N, int max_batch, int device, cmpc_handle** out) {
    if (!out) return fail("out is null");
    if (N < 1 || N > 48) return fail("horizon N must be in [1, 48]");
    if (max_batch < 1) return fail("max_batch must be >= 1");
    CU_TRY(cudaSetDevice(device));
    int sms = 0, optin = 0, per_sm = 0;
    CU_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    CU_TRY(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    CU_TRY(cudaDeviceGetAttribute(&per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device));
    cmpc_handle* h = new cmpc_handle();      // nothing below can fail: no handle is ever leaked
    h->N = N; h->W = (4 * N + 63) / 64; h->max_batch = max_batch; h->device = device;
    h->nfmax = 4 * N;
    default_params(h->p);
    h->sm_count = sms;
    h->smem_optin = (size_t)optin;
    h->smem_per_sm = (size_t)per_sm;
    if (const char* e = getenv("CMPC_PREPASS_MIN_BATCH")) { const int mb = atoi(e); if (mb >= 1) h->prepass_min_batch = mb; }
    if (const char* e = getenv("CMPC_PDAS_MAX_ITER")) { const int v = atoi(e); if (v >= 1) h->p.pdas_max_iter = v; }   // tuning
    *out = h;
    return 0;
}

int cmpc_destroy(cmpc_handle* h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    free_slots(h);
    auto& q = h->hp;
    void* ptrs[] = {q.x0, q.x_ref, q.r_foot, q.I_world, q.mass, q.t0, q.mask, q.u, q.y, q.rho, q.stats, q.status, q.iters,
                    q.R_wb, q.lever, q.cmd, q.pos_des, q.u0};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (auto s : q.s) if (s) cudaStreamDestroy(s);
    for (auto e : h->pev) if (e) cudaEventDestroy(e);
    delete h;
    return 0;
}

int cmpc_set_params(cmpc_handle* h, const double Q[12], const double R[12], double mu, double fz_min,
                    double eps_abs, double eps_rel, int max_iter, double rho0, double sigma, double alpha,
                    int mode, int polish, int check_termination, int adaptive_rho_interval) {
    if (!h) return fail("null handle");
    if (Q) for (int i = 0; i < 12; ++i) h->p.Q[i] = Q[i];
    if (R) for (int i = 0; i < 12; ++i) h->p.R[i] = R[i];
    for (int i = 0; i < 12; ++i) if (!(h->p.R[i] > 0.0) || !(h->p.Q[i] >= 0.0)) return fail("need Q >= 0 and R > 0");
    if (!(mu > 0.0) || !(eps_abs > 0.0) || !(eps_rel >= 0.0) || max_iter < 1 || !(rho0 > 0.0) || !(sigma > 0.0) ||
        !(alpha > 0.0 && alpha < 2.0) || check_termination < 1)
        return fail("invalid solver parameter");
    if (mode != CMPC_MODE_ADMM && mode != CMPC_MODE_ACTIVE_SET) return fail("unknown mode");
    h->p.mu = mu; h->p.fz_min = fz_min; h->p.eps_abs = eps_abs; h->p.eps_rel = eps_rel; h->p.max_iter = max_iter;
    h->p.rho0 = rho0; h->p.sigma = sigma; h->p.alpha = alpha; h->p.mode = mode; h->p.polish = polish;
    h->p.check_termination = check_termination; h->p.adaptive_rho_interval = adaptive_rho_interval;
    return 0;
}

int cmpc_set_max_stance(cmpc_handle* h, int nfmax) {
    if (!h) return fail("null handle");
    if (nfmax < 4) nfmax = 4;
    if (nfmax > 4 * h->N) nfmax = 4 * h->N;
    h->nfmax = nfmax;
    return 0;
}

int cmpc_set_profile(cmpc_handle* h, int on) {
    if (!h) return fail("null handle");
    CU_TRY(cudaSetDevice(h->device));
    if (on && !h->pev[0]) for (auto& e : h->pev) CU_TRY(cudaEventCreate(&e));
    h->profile = on ? 1 : 0;
    h->pev_valid = 0;
    return 0;
}

int cmpc_last_kernel_ms(cmpc_handle* h, double* prepass_ms, double* condensed_ms) {
    if (!h || !prepass_ms || !condensed_ms) return fail("null argument");
    if (!h->profile || !h->pev_valid) return fail("no profiled solve on this handle (cmpc_set_profile, then cmpc_solve with raw inputs)");
    CU_TRY(cudaSetDevice(h->device));
    CU_TRY(cudaEventSynchronize(h->pev[3]));
    float a = 0.f, b = 0.f;
    CU_TRY(cudaEventElapsedTime(&a, h->pev[0], h->pev[2]));
    CU_TRY(cudaEventElapsedTime(&b, h->pev[2], h->pev[3]));
    *prepass_ms = a; *condensed_ms = b;
    return 0;
}

int cmpc_last_kernel_ms3(cmpc_handle* h, double* sweep_ms, double* certificate_ms, double* condensed_ms) {
    if (!h || !sweep_ms || !certificate_ms || !condensed_ms) return fail("null argument");
    if (!h->profile || !h->pev_valid) return fail("no profiled solve on this handle (cmpc_set_profile, then cmpc_solve with raw inputs)");
    CU_TRY(cudaSetDevice(h->device));
    CU_TRY(cudaEventSynchronize(h->pev[3]));
    float a = 0.f, b = 0.f, c = 0.f;
    CU_TRY(cudaEventElapsedTime(&a, h->pev[0], h->pev[1]));
    CU_TRY(cudaEventElapsedTime(&b, h->pev[1], h->pev[2]));
    CU_TRY(cudaEventElapsedTime(&c, h->pev[2], h->pev[3]));
    *sweep_ms = a; *certificate_ms = b; *condensed_ms = c;
    return 0;
}

int cmpc_set_prepass(cmpc_handle* h, int on) {
    if (!h) return fail("null handle");
    h->prepass = on < 0 ? 0 : (on > 4 ? 4 : on);      // 0 off, 1-3 Riccati sweeps of round 1 (nominal robots only), 4 wrench-space PDAS (default)
    return 0;
}

int cmpc_set_generic(cmpc_handle* h, int on) {
    if (!h) return fail("null handle");
    h->force_generic = on ? 1 : 0;
    return 0;
}

int cmpc_contact_table(cmpc_handle* h, int B, const double* t0, double dt, double gait_hz, double duty,
                       const double phase_offset[4], uint64_t* mask_out, void* stream) {
    if (!h || !t0 || !mask_out || !phase_offset) return fail("null argument");
    if (B < 0) return fail("negative batch");
    if (B == 0) return 0;
    CU_TRY(cudaSetDevice(h->device));
    const double period = 1 / gait_hz;   // gait.py:17
    const int tpb = 128;
    contact_table_kernel<<<(B + tpb - 1) / tpb, tpb, 0, (cudaStream_t)stream>>>(
        B, h->N, h->W, t0, dt, period, duty, phase_offset[0], phase_offset[1], phase_offset[2], phase_offset[3], mask_out);
    ++g_launches;
    CU_TRY(cudaGetLastError());
    return 0;
}

int cmpc_generate_traj(int device, int N, int B, const double* x0, const double* R_world_to_body, const double* foot_lever,
                       const double* cmd, const double* t0, double dt, double gait_hz, double duty,
                       const double phase_offset[4], const double hip_offset[12], const double* pos_des_in,
                       double* pos_des_out, double* x_ref, double* r_foot, void* stream) {
    if (!x0 || !R_world_to_body || !foot_lever || !cmd || !t0 || !phase_offset || !hip_offset || !pos_des_in || !x_ref || !r_foot)
        return fail("null argument");
    if (N < 1 || N > 48) return fail("horizon N must be in [1, 48]");
    if (B < 0) return fail("negative batch");
    if (B == 0) return 0;
    CU_TRY(cudaSetDevice(device));
    const double period = 1 / gait_hz;   // gait.py:17
    const int tpb = 128, total = 4 * B;
    const double* h = hip_offset;
    generate_traj_kernel<<<(total + tpb - 1) / tpb, tpb, 0, (cudaStream_t)stream>>>(
        B, N, x0, R_world_to_body, foot_lever, cmd, t0, dt, period, duty, phase_offset[0], phase_offset[1], phase_offset[2],
        phase_offset[3], h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8], h[9], h[10], h[11], pos_des_in, pos_des_out,
        x_ref, r_foot);
    ++g_launches;
    CU_TRY(cudaGetLastError());
    return 0;
}

int cmpc_srb_step(int device, int N, int B, const double* x, const double* u, const double* x_ref, const double* r_foot,
                  const double* I_world, const double* mass, double T, const double I_body[3], const double stance_offset[12],
                  double* x_out, double* R_world_to_body_out, double* I_world_out, double* foot_lever_out, void* stream) {
    if (!x || !u || !x_ref || !r_foot || !I_world || !mass || !I_body || !stance_offset || !x_out || !R_world_to_body_out ||
        !I_world_out || !foot_lever_out)
        return fail("null argument");
    if (N < 1 || N > 48) return fail("horizon N must be in [1, 48]");
    if (B < 0) return fail("negative batch");
    if (B == 0) return 0;
    CU_TRY(cudaSetDevice(device));
    const int tpb = 128;
    const double* s = stance_offset;
    srb_step_kernel<<<(B + tpb - 1) / tpb, tpb, 0, (cudaStream_t)stream>>>(
        B, N, x, u, x_ref, r_foot, I_world, mass, T, I_body[0], I_body[1], I_body[2], s[0], s[1], s[2], s[3], s[4], s[5], s[6],
        s[7], s[8], s[9], s[10], s[11], x_out, R_world_to_body_out, I_world_out, foot_lever_out);
    ++g_launches;
    CU_TRY(cudaGetLastError());
    return 0;
}

int cmpc_stance_torque(int device, int N, int B, const double* J_foot_world, const double* u, const double* time_now,
                       double gait_hz, double duty, const double phase_offset[4], double tau_max, double* tau,
                       int32_t* mask_now, void* stream) {
    if (!J_foot_world || !u || !time_now || !phase_offset || !tau) return fail("null argument");
    if (N < 1 || N > 48) return fail("horizon N must be in [1, 48]");
    if (B < 0) return fail("negative batch");
    if (B == 0) return 0;
    CU_TRY(cudaSetDevice(device));
    const double period = 1 / gait_hz;   // gait.py:17
    const int tpb = 128, total = 4 * B;
    stance_torque_kernel<<<(total + tpb - 1) / tpb, tpb, 0, (cudaStream_t)stream>>>(
        B, N, J_foot_world, u, time_now, period, duty, phase_offset[0], phase_offset[1], phase_offset[2], phase_offset[3],
        tau_max, tau, mask_now);
    ++g_launches;
    CU_TRY(cudaGetLastError());
    return 0;
}

int cmpc_leg_jacobian(int device, int B, const double* q_joint, const double* R_world_to_body, const double link[3],
                      double* J_foot_world, double* foot_pos_body, void* stream) {
    if (!q_joint || !R_world_to_body || !link || !J_foot_world) return fail("null argument");
    if (B < 0) return fail("negative batch");
    if (B == 0) return 0;
    CU_TRY(cudaSetDevice(device));
    const int tpb = 128, total = 4 * B;
    leg_jacobian_kernel<<<(total + tpb - 1) / tpb, tpb, 0, (cudaStream_t)stream>>>(B, q_joint, R_world_to_body, link[0], link[1], link[2],
                                                                                 J_foot_world, foot_pos_body);
    ++g_launches;
    CU_TRY(cudaGetLastError());
    return 0;
}

int cmpc_pack_contact(cmpc_handle* h, int B, const int32_t* table, uint64_t* mask_out, void* stream) {
    if (!h || !table || !mask_out) return fail("null argument");
    if (B < 0) return fail("negative batch");
    if (B == 0) return 0;
    CU_TRY(cudaSetDevice(h->device));
    const int tpb = 128;
    pack_contact_kernel<<<(B + tpb - 1) / tpb, tpb, 0, (cudaStream_t)stream>>>(B, h->N, h->W, table, mask_out);
    ++g_launches;
    CU_TRY(cudaGetLastError());
    return 0;
}

int cmpc_dynamics(cmpc_handle* h, int B, const double* x_ref, const double* r_foot, const double* I_world,
                  const double* mass, double dt, double* Ad, double* Bd, double* gd, void* stream) {
    if (!h || !x_ref || !r_foot || !I_world || !mass || !Ad || !Bd || !gd) return fail("null argument");
    if (B < 0) return fail("negative batch");
    if (B == 0) return 0;
    CU_TRY(cudaSetDevice(h->device));
    const int tpb = 128, total = B * h->N;
    dynamics_kernel<<<(total + tpb - 1) / tpb, tpb, 0, (cudaStream_t)stream>>>(B, h->N, x_ref, r_foot, I_world, mass, dt, Ad, Bd, gd);
    ++g_launches;
    CU_TRY(cudaGetLastError());
    return 0;
}

static int check_inputs(const double* Ad, const double* Bd, const double* gd, const double* x0, const double* x_ref,
                        const double* r_foot, const double* I_world, const double* mass) {
    if (!x0 || !x_ref) return fail("x0 and x_ref are required");
    const bool have_ab = Ad && Bd && gd;
    const bool have_raw = r_foot && I_world && mass;
    if (!have_ab && !have_raw) return fail("need either (Ad,Bd,gd) or (r_foot,I_world,mass)");
    if ((Ad || Bd || gd) && !have_ab) return fail("Ad, Bd and gd must be given together");
    return 0;
}

int cmpc_build(cmpc_handle* h, int B, const double* Ad, const double* Bd, const double* gd, const double* x0,
               const double* x_ref, const double* r_foot, const double* I_world, const double* mass, double dt,
               double* H, double* g, void* stream) {
    if (!h || !H || !g) return fail("null argument");
    if (check_inputs(Ad, Bd, gd, x0, x_ref, r_foot, I_world, mass)) return -1;
    if (B < 0) return fail("negative batch");
    if (B == 0) return 0;
    CU_TRY(cudaSetDevice(h->device));
    if (ensure_reserved(h, B)) return -1;
    BatchIn bi{Ad, Bd, gd, x0, x_ref, r_foot, I_world, mass, dt, nullptr, h->N, h->W};
    const size_t stride = hp_stride_generic(h, 4 * h->N);
    const size_t smem = smem_needed(h->N, 4 * h->N, stride != 0);
    if (smem > h->smem_optin) return fail("workspace does not fit shared memory even with the matrix in global memory");
    if (set_smem((const void*)build_kernel, smem, &h->attr_build)) return -1;
    auto& sl = h->slot[h->slot_next.fetch_add(1) & 3u];
    const int cap = h->sm_count * 2;
    const int grid = stride ? (B < cap ? B : cap) : B;
    build_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(h->p, bi, B, 4 * h->N, H, g, stride ? sl.hp : nullptr, stride);
    ++g_launches;
    CU_TRY(cudaGetLastError());
    return 0;
}

int cmpc_workspace_bytes(cmpc_handle* h, int B, size_t* bytes) {
    if (!h || !bytes) return fail("null argument");
    if (B < 1) return fail("batch must be >= 1");
    *bytes = slot_sizes(h, B).total;
    return 0;
}

int cmpc_reserve(cmpc_handle* h, int B) {
    if (!h) return fail("null handle");
    if (B < 1) return fail("batch must be >= 1");
    CU_TRY(cudaSetDevice(h->device));
    if (B > h->max_batch) h->max_batch = B;
    return reserve_slots(h, B);
}

int cmpc_solve(cmpc_handle* h, int B, const double* Ad, const double* Bd, const double* gd, const double* x0,
               const double* x_ref, const double* r_foot, const double* I_world, const double* mass, double dt,
               const uint64_t* mask, int warm, double* u, double* y, double* rho, double* X, double* nu,
               int32_t* status, int32_t* iters, double* stats, void* stream) {
    if (!h || !u || !y || !status || !iters || !stats) return fail("null argument");
    if (check_inputs(Ad, Bd, gd, x0, x_ref, r_foot, I_world, mass)) return -1;
    if (B < 0) return fail("negative batch");
    if (B == 0) return 0;
    CU_TRY(cudaSetDevice(h->device));
    if (ensure_reserved(h, B)) return -1;         // allocates only before cmpc_reserve / the first solve
    cudaStream_t st = (cudaStream_t)stream;
    BatchIn bi{Ad, Bd, gd, x0, x_ref, r_foot, I_world, mass, dt, mask, h->N, h->W};
    BatchOut bo{u, y, rho, X, nu, status, iters, stats};
    // every in-flight solve owns one slot: work-list, counters, gains and the L2-resident scratch of the condensed
    // kernel are never shared between solves enqueued on different streams
    auto& sl = h->slot[h->slot_next.fetch_add(1) & 3u];
    if (!Ad && !h->force_generic) {
        // raw inputs: closed-form dynamics on the device
        const size_t hstride = hp_stride_fast(h, h->nfmax);
        const size_t smem = smem_needed_fast(h->N, h->nfmax, hstride != 0);
        if (smem > h->smem_optin) return fail("workspace does not fit shared memory even with the matrix in global memory");
        int per_sm = 2;
        if (!hstride) {
            per_sm = (int)(h->smem_per_sm / (smem + 1024));      // 1 KB per CTA is reserved by the system
            if (per_sm < 1) per_sm = 1;
            if (per_sm > 2) per_sm = 2;                           // register file: 256 threads x 128 registers x 2
            if (const char* e = getenv("CMPC_CTAS_PER_SM")) { const int v = atoi(e); if (v >= 1 && v < per_sm) per_sm = v; }   // diagnostics
        }
        if (hstride) { if (set_smem((const void*)solve_fast_kernel<true>, smem, &h->attr_fast[1])) return -1; }
        else if (set_smem((const void*)solve_fast_kernel<false>, smem, &h->attr_fast[0])) return -1;
        int grid_f = h->sm_count * per_sm;
        if (B < grid_f) grid_f = B;
        const int* wl = nullptr;
        const int* wlc = nullptr;
        bool forked = false;
        int cert_grid = 0;
        wr::Bat cert_bt{};
        // small batches are latency-bound (fewer robots than CTA slots x a few rounds): the extra kernel in front only
        // adds to the latency there (batch 1: 45 -> 99 us), so the pre-pass starts at prepass_min_batch robots
        if (h->profile) { for (int i = 0; i < 3; ++i) CU_TRY(cudaEventRecord(h->pev[i], st)); h->pev_valid = 0; }
        if (h->prepass && h->p.mode == CMPC_MODE_ACTIVE_SET && B >= h->prepass_min_batch) {
            CU_TRY(cudaMemsetAsync(sl.ctl, 0, 8 * sizeof(int), st));
            bool launched = false;
            if (h->prepass == 4) {
                // wrench-space projected Riccati + PDAS: finishes nominal and constrained robots alike; what it
                // cannot finish (cycling working sets) goes to the condensed kernel through the work-list
                const size_t smem_w = wrench_smem(h);
                const int grid_w = wrench_grid(h, B);
                if ((h->N & 3) == 0 && smem_w <= h->smem_optin && (size_t)grid_w * h->N * wr::GAIN_D2 * kWrThreads * sizeof(wr::D2) <= sl.gain_bytes) {
                    if (set_smem((const void*)wrench_pdas_kernel, smem_w, &h->attr_wrench)) return -1;
                    robot_consts_kernel<<<(B + 127) / 128, 128, 0, st>>>(B, h->N, x_ref, I_world, mass, dt, sl.cst);
                    ++g_launches;
                    wr::Bat bt{x0, x_ref, r_foot, I_world, mass, mask, u, y, rho, X ? X : sl.xs, nu, stats, status, iters, dt, h->N, h->W, sl.cst};
                    wrench_pdas_kernel<<<grid_w, kWrThreads, smem_w, st>>>(h->p, bt, B, h->nfmax, warm, reinterpret_cast<wr::D2*>(sl.gains),
                                                                         sl.worklist, sl.ctl, wr::robot_bytes(h->N));
                    ++g_launches;
                    CU_TRY(cudaGetLastError());
                    if (h->profile) CU_TRY(cudaEventRecord(h->pev[1], st));
                    const int want_c = (B + 7) / 8, cap_c = h->sm_count * 128;
                    cert_grid = want_c < cap_c ? want_c : cap_c;
                    cert_bt = bt;
                    forked = true;
                    launched = true;
                }
            }
            if (!launched) {
                // Riccati pre-pass of round 1: finishes the robots without an active constraint, lists the others
                ric::WsR wr_;
                const size_t per_warp = (ric::ws_carve_ric(wr_, reinterpret_cast<unsigned char*>(static_cast<uintptr_t>(1 << 20)), h->N) + 15) & ~(size_t)15;
                const int wpb = kRicThreads / 32;
                const size_t smem_r = per_warp * wpb;
                const size_t gd_ = ric::gain_doubles(h->nfmax);
                const size_t per_half = ric2::half_bytes(h->N);
                const int hpb = (kRic2Threads / 32) * 2;
                const size_t smem_2 = per_half * hpb;
                int per_sm2 = (int)(h->smem_per_sm / (smem_2 + 1024));
                if (per_sm2 > 8) per_sm2 = 8;
                // lock-step CTA: as many robots as fit shared memory (pairs: two per warp), at most sixteen
                int hpl = (int)(h->smem_optin / per_half) & ~1;
                if (hpl > (kRic2LsThreads / 32) * 2) hpl = (kRic2LsThreads / 32) * 2;
                const size_t smem_ls = per_half * (size_t)hpl;
                const int mode = h->prepass == 4 ? 3 : h->prepass;
                if (mode == 3 && hpl >= 4) {
                    const int want3 = (B + hpl - 1) / hpl;
                    if (set_smem((const void*)riccati2_lockstep_kernel, smem_ls, &h->attr_ric[2])) return -1;
                    riccati2_lockstep_kernel<<<want3 < h->sm_count ? want3 : h->sm_count, 16 * hpl, smem_ls, st>>>(
                        h->p, bi, bo, B, h->nfmax, warm, sl.gains, gd_, sl.worklist, sl.ctl, per_half);
                    launched = true;
                } else if (mode >= 2 && per_sm2 >= 1 && gd_ * (size_t)h->sm_count * per_sm2 * hpb * sizeof(double) <= sl.gain_bytes) {
                    const int want2 = (B + hpb - 1) / hpb, cap2 = h->sm_count * per_sm2;
                    if (set_smem((const void*)riccati2_kernel, smem_2, &h->attr_ric[1])) return -1;
                    riccati2_kernel<<<want2 < cap2 ? want2 : cap2, kRic2Threads, smem_2, st>>>(
                        h->p, bi, bo, B, h->nfmax, warm, sl.gains, gd_, sl.worklist, sl.ctl, per_half);
                    launched = true;
                } else if (smem_r <= h->smem_optin) {
                    int per_sm_r = (int)(h->smem_per_sm / (smem_r + 1024));
                    if (per_sm_r > 3) per_sm_r = 3;
                    if (per_sm_r < 1) per_sm_r = 1;
                    const int want = (B + wpb - 1) / wpb, cap_r = h->sm_count * per_sm_r;
                    if (set_smem((const void*)riccati_kernel, smem_r, &h->attr_ric[0])) return -1;
                    riccati_kernel<<<want < cap_r ? want : cap_r, kRicThreads, smem_r, st>>>(h->p, bi, bo, B, h->nfmax, warm, sl.gains, gd_,
                                                                                          sl.worklist, sl.ctl, per_warp);
                    launched = true;
                }
            }
            if (launched) {
                ++g_launches;
                CU_TRY(cudaGetLastError());
                wl = sl.worklist; wlc = sl.ctl;
                const int cap_f = h->sm_count * 2;
                grid_f = B < cap_f ? B : cap_f;
                if (h->profile) { if (h->prepass != 4) CU_TRY(cudaEventRecord(h->pev[1], st)); CU_TRY(cudaEventRecord(h->pev[2], st)); }
            }
        }
        auto condensed = [&](cudaStream_t s_, const int* list, const int* cnt) {
            if (hstride) solve_fast_kernel<true><<<grid_f, kThreads, smem, s_>>>(h->p, bi, bo, B, h->nfmax, warm, sl.hp, hstride, list, cnt, sl.yg, yg_stride_of(h, h->nfmax));
            else solve_fast_kernel<false><<<grid_f, kThreads, smem, s_>>>(h->p, bi, bo, B, h->nfmax, warm, nullptr, 0, list, cnt, sl.yg, yg_stride_of(h, h->nfmax));
            ++g_launches;
        };
        if (forked) {
            // sweep kernel -> { condensed kernel on the hand-overs, on st  ||  certificates on aux } -> condensed kernel on the
            // robots whose certificate did not hold (second list; normally empty: one CTA round of ~8 us).  The condensed kernel
            // stays on the caller's stream so that its few big CTAs (32 K registers, 106 KB shared memory each) are placed before
            // the 8 192 small CTAs of the certificate kernel arrive: placed second they starve until that kernel has drained.
            CU_TRY(cudaEventRecord(sl.ev_fork, st));
            CU_TRY(cudaStreamWaitEvent(sl.aux, sl.ev_fork, 0));
            condensed(st, wl, wlc);
            CU_TRY(cudaGetLastError());
            // ... and the block scheduler places whatever becomes ready first: the certificate kernel waits ~15 us behind the
            // fork so that the condensed kernel's CTAs are resident when its own arrive (without it the order is a coin toss:
            // 2.56 instead of 2.40 ms per 65 536 robots)
            head_start_kernel<<<1, 1, 0, sl.aux>>>(30000);
            if (h->profile) CU_TRY(cudaEventRecord(h->pev[1], sl.aux));
            wrench_certificate_kernel<<<cert_grid, 128, 0, sl.aux>>>(h->p, cert_bt, B, warm, sl.worklist + sl.wl_cap, sl.ctl + 4);
            CU_TRY(cudaGetLastError());          // (counted with the pre-pass above)
            if (h->profile) CU_TRY(cudaEventRecord(h->pev[2], sl.aux));
            CU_TRY(cudaEventRecord(sl.ev_join, sl.aux));
            CU_TRY(cudaStreamWaitEvent(st, sl.ev_join, 0));
            condensed(st, sl.worklist + sl.wl_cap, sl.ctl + 4);
        } else {
            condensed(st, wl, wlc);
        }
        CU_TRY(cudaGetLastError());
        if (h->profile) { CU_TRY(cudaEventRecord(h->pev[3], st)); h->pev_valid = 1; }
        return 0;
    }
    const size_t stride = hp_stride_generic(h, h->nfmax);
    const size_t smem = smem_needed(h->N, h->nfmax, stride != 0);
    if (smem > h->smem_optin) return fail("workspace does not fit shared memory even with the matrix in global memory");
    if (set_smem((const void*)solve_kernel, smem, &h->attr_generic)) return -1;
    const int cap = h->sm_count * 2;
    const int grid = stride ? (B < cap ? B : cap) : B;
    solve_kernel<<<grid, kThreads, smem, st>>>(h->p, bi, bo, B, h->nfmax, warm, stride ? sl.hp : nullptr, stride);
    ++g_launches;
    CU_TRY(cudaGetLastError());
    return 0;
}

// device-resident buffers of the *_host entries, sized for max_batch robots on first use
static int host_path_ready(cmpc_handle* h) {
    const int N = h->N;
    auto& q = h->hp;
    if (!q.ready) {
        const size_t mb = (size_t)h->max_batch;
        CU_TRY(cudaMalloc(&q.x0, mb * 12 * sizeof(double)));
        CU_TRY(cudaMalloc(&q.x_ref, mb * 12 * N * sizeof(double)));
        CU_TRY(cudaMalloc(&q.r_foot, mb * 12 * N * sizeof(double)));
        CU_TRY(cudaMalloc(&q.I_world, mb * 9 * sizeof(double)));
        CU_TRY(cudaMalloc(&q.mass, mb * sizeof(double)));
        CU_TRY(cudaMalloc(&q.t0, mb * sizeof(double)));
        CU_TRY(cudaMalloc(&q.mask, mb * h->W * sizeof(uint64_t)));
        CU_TRY(cudaMalloc(&q.u, mb * 12 * N * sizeof(double)));
        CU_TRY(cudaMalloc(&q.y, mb * 28 * N * sizeof(double)));
        CU_TRY(cudaMalloc(&q.rho, mb * sizeof(double)));
        CU_TRY(cudaMalloc(&q.stats, mb * CMPC_NSTAT * sizeof(double)));
        CU_TRY(cudaMalloc(&q.status, mb * sizeof(int32_t)));
        CU_TRY(cudaMalloc(&q.iters, mb * sizeof(int32_t)));
        CU_TRY(cudaMemset(q.rho, 0, mb * sizeof(double)));
        CU_TRY(cudaMemset(q.u, 0, mb * 12 * N * sizeof(double)));
        CU_TRY(cudaMemset(q.y, 0, mb * 28 * N * sizeof(double)));
        for (auto& s : q.s) CU_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        CU_TRY(cudaMalloc(&q.R_wb, mb * 9 * sizeof(double)));
        CU_TRY(cudaMalloc(&q.lever, mb * 12 * sizeof(double)));
        CU_TRY(cudaMalloc(&q.cmd, mb * 4 * sizeof(double)));
        CU_TRY(cudaMalloc(&q.pos_des, mb * 3 * sizeof(double)));
        CU_TRY(cudaMalloc(&q.u0, mb * 12 * sizeof(double)));
        q.ready = true;
    }
    return 0;
}

int cmpc_solve_host(cmpc_handle* h, int B, const double* x0, const double* x_ref, const double* r_foot,
                    const double* I_world, const double* mass, const double* t0, double dt, double gait_hz,
                    double duty, const double phase_offset[4], int warm, double* u, int32_t* status, int32_t* iters) {
    if (!h || !x0 || !x_ref || !r_foot || !I_world || !mass || !t0 || !phase_offset || !u || !status || !iters)
        return fail("null argument");
    if (B < 0 || B > h->max_batch) return fail("batch exceeds max_batch of the handle");
    if (B == 0) return 0;
    CU_TRY(cudaSetDevice(h->device));
    const int N = h->N;
    auto& q = h->hp;
    if (host_path_ready(h)) return -1;
    // chunks sized so that copies of one chunk hide behind the solve of the other
    // measured at 65 536 robots with the round-2 kernels (tools/host_path_probe.py): one chunk 9.3 ms, 32768 -> 7.4, 16384 -> 7.7,
    // 8192 -> 9.5, 4096 -> 13.5 (a chunk's kernel cannot end before its slowest robot has, so small chunks cost tails)
    int chunk = B <= 8192 ? B : (B < 65536 ? (B + 1) / 2 : 32768);
    if (const char* e = getenv("CMPC_HOST_CHUNK")) { const int v = atoi(e); if (v >= 256) chunk = v < B ? v : B; }   // tuning
    // the first copy in and the last copy out cannot hide behind anything: ramp the chunk size up from 4 096
    // (4096, 8192, then `chunk`) and finish with a 4 096-robot chunk, so that only small transfers are exposed
    const bool ramp = (B >= 8 * chunk) && !getenv("CMPC_HOST_CHUNK");
    // tuning: an explicit chunk schedule, e.g. CMPC_HOST_CHUNKS=16384,32768,16384 (the last entry repeats)
    int sched[16], nsched = 0;
    if (const char* e = getenv("CMPC_HOST_CHUNKS")) {
        const char* q_ = e;
        while (*q_ && nsched < 16) { const int v = atoi(q_); if (v >= 256) sched[nsched++] = v; while (*q_ && *q_ != ',') ++q_; if (*q_ == ',') ++q_; }
    } else if (B >= 65536 && !getenv("CMPC_HOST_CHUNK")) {
        // a small first chunk lets the kernels start early, a small last one shortens the copy out that nothing can hide
        // (measured at 65 536 robots, tools/host_path_probe.py: 8192,24576,24576,8192 -> 6.4 ms; two halves 7.0; four quarters 7.8)
        sched[0] = B / 8; sched[1] = 3 * (B / 8); sched[2] = 3 * (B / 8); sched[3] = B - sched[0] - sched[1] - sched[2]; nsched = 4;
    }
    int ci = 0, nb = 0, rc = 0;
    for (int lo = 0; lo < B && !rc; lo += nb, ++ci) {
        int want = chunk;
        if (nsched) want = sched[ci < nsched ? ci : nsched - 1];
        else if (ramp) {
            if (ci == 0) want = 4096;
            else if (ci == 1) want = 8192;
            const int rem = B - lo;
            if (rem > 4096 && rem - want < 4096) want = rem - 4096;      // leave a small last chunk
        }
        nb = (B - lo) < want ? (B - lo) : want;
        cudaStream_t s = q.s[ci & 1];
        const size_t o = (size_t)lo;
        auto up = [&](double* d, const double* src, size_t w) {
            return cudaMemcpyAsync(d + o * w, src + o * w, (size_t)nb * w * sizeof(double), cudaMemcpyHostToDevice, s);
        };
        if (up(q.x0, x0, 12) || up(q.x_ref, x_ref, (size_t)12 * N) || up(q.r_foot, r_foot, (size_t)12 * N) || up(q.I_world, I_world, 9) ||
            up(q.mass, mass, 1) || up(q.t0, t0, 1)) { rc = fail("host-to-device copy failed"); break; }
        if (cmpc_contact_table(h, nb, q.t0 + o, dt, gait_hz, duty, phase_offset, q.mask + o * h->W, s) ||
            cmpc_solve(h, nb, nullptr, nullptr, nullptr, q.x0 + o * 12, q.x_ref + o * 12 * N, q.r_foot + o * 12 * N,
                       q.I_world + o * 9, q.mass + o, dt, q.mask + o * h->W, warm, q.u + o * 12 * N, q.y + o * 28 * N,
                       q.rho + o, nullptr, nullptr, q.status + o, q.iters + o, q.stats + o * CMPC_NSTAT, s)) { rc = -1; break; }
        cudaError_t e = cudaMemcpyAsync(u + o * 12 * N, q.u + o * 12 * N, (size_t)nb * 12 * N * sizeof(double), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(status + o, q.status + o, (size_t)nb * sizeof(int32_t), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(iters + o, q.iters + o, (size_t)nb * sizeof(int32_t), cudaMemcpyDeviceToHost, s);
        if (e != cudaSuccess) { rc = fail(cudaGetErrorString(e)); break; }
    }
    // copies from / to the caller's buffers may still be in flight: never return before both streams have drained
    const cudaError_t e0 = cudaStreamSynchronize(q.s[0]), e1 = cudaStreamSynchronize(q.s[1]);
    if (rc) return rc;
    if (e0 != cudaSuccess || e1 != cudaSuccess) return fail(cudaGetErrorString(e0 != cudaSuccess ? e0 : e1));
    return 0;
}

int cmpc_cycle_host(cmpc_handle* h, int B, const double* x0, const double* R_world_to_body, const double* foot_lever,
                    const double* cmd, const double* t0, double* pos_des, const double* I_world, const double* mass,
                    double dt, double gait_hz, double duty, const double phase_offset[4], const double hip_offset[12],
                    int warm, int first_step_only, double* u, int32_t* status, int32_t* iters) {
    if (!h || !x0 || !R_world_to_body || !foot_lever || !cmd || !t0 || !pos_des || !I_world || !mass || !phase_offset ||
        !hip_offset || !u || !status || !iters)
        return fail("null argument");
    if (B < 0 || B > h->max_batch) return fail("batch exceeds max_batch of the handle");
    if (B == 0) return 0;
    CU_TRY(cudaSetDevice(h->device));
    if (host_path_ready(h)) return -1;
    const int N = h->N;
    auto& q = h->hp;
    // 408 bytes in and 96 (first step) or 96 N bytes out per robot: two halves on two streams are enough to overlap
    // what little is copied with the kernels
    // (full-horizon forces out: 1.5 KB per robot, worth overlapping; first step only: one chunk, one kernel tail)
    const int nchunk = (B >= 8192 && !first_step_only) ? 2 : 1;
    const int per = (B + nchunk - 1) / nchunk;
    int rc = 0;
    for (int ci = 0; ci < nchunk && !rc; ++ci) {
        const int lo = ci * per, nb = (B - lo) < per ? (B - lo) : per;
        if (nb <= 0) break;
        cudaStream_t s = q.s[ci & 1];
        const size_t o = (size_t)lo;
        auto up = [&](double* d, const double* src, size_t w) {
            return cudaMemcpyAsync(d + o * w, src + o * w, (size_t)nb * w * sizeof(double), cudaMemcpyHostToDevice, s);
        };
        if (up(q.x0, x0, 12) || up(q.R_wb, R_world_to_body, 9) || up(q.lever, foot_lever, 12) || up(q.cmd, cmd, 4) ||
            up(q.t0, t0, 1) || up(q.pos_des, pos_des, 3) || up(q.I_world, I_world, 9) || up(q.mass, mass, 1)) { rc = fail("host-to-device copy failed"); break; }
        if (cmpc_generate_traj(h->device, N, nb, q.x0 + o * 12, q.R_wb + o * 9, q.lever + o * 12, q.cmd + o * 4, q.t0 + o, dt,
                               gait_hz, duty, phase_offset, hip_offset, q.pos_des + o * 3, q.pos_des + o * 3,
                               q.x_ref + o * 12 * N, q.r_foot + o * 12 * N, s) ||
            cmpc_contact_table(h, nb, q.t0 + o, dt, gait_hz, duty, phase_offset, q.mask + o * h->W, s) ||
            cmpc_solve(h, nb, nullptr, nullptr, nullptr, q.x0 + o * 12, q.x_ref + o * 12 * N, q.r_foot + o * 12 * N,
                       q.I_world + o * 9, q.mass + o, dt, q.mask + o * h->W, warm, q.u + o * 12 * N, q.y + o * 28 * N,
                       q.rho + o, nullptr, nullptr, q.status + o, q.iters + o, q.stats + o * CMPC_NSTAT, s)) { rc = -1; break; }
        cudaError_t e;
        if (first_step_only) {     // U_opt[:, 0], the only column the consumer applies (test_MPC.py:196)
            first_step_kernel<<<(12 * nb + 255) / 256, 256, 0, s>>>(nb, N, q.u + o * 12 * N, q.u0 + o * 12);
            ++g_launches;
            e = cudaMemcpyAsync(u + o * 12, q.u0 + o * 12, (size_t)nb * 12 * sizeof(double), cudaMemcpyDeviceToHost, s);
        }
        else
            e = cudaMemcpyAsync(u + o * 12 * N, q.u + o * 12 * N, (size_t)nb * 12 * N * sizeof(double), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(pos_des + o * 3, q.pos_des + o * 3, (size_t)nb * 3 * sizeof(double), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(status + o, q.status + o, (size_t)nb * sizeof(int32_t), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(iters + o, q.iters + o, (size_t)nb * sizeof(int32_t), cudaMemcpyDeviceToHost, s);
        if (e != cudaSuccess) { rc = fail(cudaGetErrorString(e)); break; }
    }
    // copies from / to the caller's buffers may still be in flight: never return before both streams have drained
    const cudaError_t e0 = cudaStreamSynchronize(q.s[0]), e1 = cudaStreamSynchronize(q.s[1]);
    if (rc) return rc;
    if (e0 != cudaSuccess || e1 != cudaSuccess) return fail(cudaGetErrorString(e0 != cudaSuccess ? e0 : e1));
    return 0;
}

int cmpc_host_stats(cmpc_handle* h, int B, double* stats_host) {
    if (!h || !stats_host || !h->hp.ready) return fail("no host-path state");
    if (B < 0 || B > h->max_batch) return fail("batch exceeds max_batch of the handle");
    if (B == 0) return 0;
    CU_TRY(cudaSetDevice(h->device));
    CU_TRY(cudaMemcpy(stats_host, h->hp.stats, (size_t)B * CMPC_NSTAT * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

int cmpc_microbench(int device, double* fp64_tflops, double* smem_gbs) {
    CU_TRY(cudaSetDevice(device));
    int sms = 0;
    CU_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    double* out = nullptr;
    const int tpb = 256, blocks = sms * 8;
    CU_TRY(cudaMalloc(&out, (size_t)tpb * blocks * sizeof(double)));
    cudaEvent_t e0, e1;
    CU_TRY(cudaEventCreate(&e0));
    CU_TRY(cudaEventCreate(&e1));
    float ms = 0;
    if (fp64_tflops) {
        const int iters = 1 << 14;
        fp64_peak_kernel<<<blocks, tpb>>>(out, 64);
        CU_TRY(cudaEventRecord(e0));
        fp64_peak_kernel<<<blocks, tpb>>>(out, iters);
        CU_TRY(cudaEventRecord(e1));
        CU_TRY(cudaEventSynchronize(e1));
        CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
        g_launches += 2;
        *fp64_tflops = 2.0 * 8.0 * iters * (double)tpb * blocks / (ms * 1e-3) / 1e12;
    }
    if (smem_gbs) {
        const int iters = 1 << 12;
        const size_t smem = 64 * 1024;
        CU_TRY(cudaFuncSetAttribute((const void*)smem_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int b2 = sms * 3;
        smem_peak_kernel<<<b2, tpb, smem>>>(out, 8);
        CU_TRY(cudaEventRecord(e0));
        smem_peak_kernel<<<b2, tpb, smem>>>(out, iters);
        CU_TRY(cudaEventRecord(e1));
        CU_TRY(cudaEventSynchronize(e1));
        CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
        g_launches += 2;
        *smem_gbs = 16.0 * 8.0 * iters * (double)tpb * b2 / (ms * 1e-3) / 1e9;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    CU_TRY(cudaGetLastError());
    return 0;
}

#ifdef CMPC_PHASE_TIMING
/* tools only (libcmpc_timing.so): cycles and visit counts per phase of solve_one_fast since the last reset */
int cmpc_debug_phase_cycles(double* out /* 2*16 */, int reset) {
    unsigned long long h[2 * CMPC_NPHASE];
    CU_TRY(cudaMemcpyFromSymbol(h, cmpc::fast::g_phase_cycles, sizeof(h)));
    for (int i = 0; i < 2 * CMPC_NPHASE; ++i) out[i] = (double)h[i];
    if (reset) { memset(h, 0, sizeof(h)); CU_TRY(cudaMemcpyToSymbol(cmpc::fast::g_phase_cycles, h, sizeof(h))); }
    return 0;
}
#endif

int cmpc_microbench_latency(int device, double* out4) {
    CU_TRY(cudaSetDevice(device));
    double* out = nullptr;
    long long* cyc = nullptr;
    CU_TRY(cudaMalloc(&out, 32 * sizeof(double)));
    CU_TRY(cudaMalloc(&cyc, 4 * sizeof(long long)));
    const int iters = 2048;
    latency_kernel<<<1, 32>>>(out, cyc, iters);
    long long h[4];
    CU_TRY(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
    ++g_launches;
    for (int i = 0; i < 4; ++i) out4[i] = (double)h[i] / iters;
    cudaFree(out);
    cudaFree(cyc);
    return 0;
}

int cmpc_microbench_dmma(int device, double* dmma_tflops, double* dmma_latency_cycles) {
    CU_TRY(cudaSetDevice(device));
    int sms = 0;
    CU_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    double* out = nullptr;
    long long* cyc = nullptr;
    const int tpb = 256, blocks = sms * 8;
    CU_TRY(cudaMalloc(&out, (size_t)tpb * blocks * sizeof(double)));
    CU_TRY(cudaMalloc(&cyc, sizeof(long long)));
    cudaEvent_t e0, e1;
    CU_TRY(cudaEventCreate(&e0));
    CU_TRY(cudaEventCreate(&e1));
    float ms = 0;
    if (dmma_tflops) {
        const int iters = 1 << 12;
        dmma_peak_kernel<<<blocks, tpb>>>(out, 16);
        CU_TRY(cudaEventRecord(e0));
        dmma_peak_kernel<<<blocks, tpb>>>(out, iters);
        CU_TRY(cudaEventRecord(e1));
        CU_TRY(cudaEventSynchronize(e1));
        CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
        g_launches += 2;
        // one m8n8k4 = 256 FMA = 512 flop per warp
        *dmma_tflops = 512.0 * 8.0 * iters * (double)(tpb / 32) * blocks / (ms * 1e-3) / 1e12;
    }
    if (dmma_latency_cycles) {
        const int iters = 4096;
        dmma_latency_kernel<<<1, 32>>>(out, cyc, iters);
        long long hc = 0;
        CU_TRY(cudaMemcpy(&hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost));
        ++g_launches;
        *dmma_latency_cycles = (double)hc / iters;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    cudaFree(cyc);
    CU_TRY(cudaGetLastError());
    return 0;
}

}  // extern "C"
