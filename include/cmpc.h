/* cmpc.h -- C-ABI of the B200-native batched Go2 convex-MPC hot path.
 *
 * The reference (ltinphan/convex-mpc-unitree-go2 v0.3.0) has no FFI of its own: its boundary is
 * the Python class `CentroidalMPC` (convex_mpc/centroidal_mpc.py:40-120) which calls CasADi->OSQP
 * in-process.  These entry points are what a ctypes binding for that class binds;
 * each one cites the reference code it replaces.  All array arguments are raw pointers,
 * row-major, leading batch dimension B, FP64 unless stated.  `stream` is a cudaStream_t passed as
 * void* (NULL = default stream).  Every function returns 0 on success, <0 on error
 * (cmpc_last_error() gives the message).  Functions whose name ends in `_host` take HOST pointers
 * and do their own H2D/D2H; all others take DEVICE pointers and only enqueue work on `stream`.
 *
 * Variable / row orderings (identical to the reference):
 *   u     (B, 12N)   force vector in the order of w[12N:] (test_MPC.py:189-192):
 *                    index 12k + 3 leg + c, legs FL FR RL RR, c = x,y,z, world frame.
 *   X     (B, 12N)   predicted states x_1..x_N in the order of w[:12N], index 12k + i.
 *   y     (B, 28N)   duals of the condensed QP rows: [0,12N) box rows on u (== lam_x[12N:]),
 *                    [12N,28N) friction rows 16k + 4 leg + face, faces +fx,-fx,+fy,-fy
 *                    (== lam_a[12N:], centroidal_mpc.py:324-359).  Sign: y>0 at upper bounds.
 *   nu    (B, 12N)   co-states == lam_a[:12N] (dynamics equality rows, centroidal_mpc.py:287-303).
 *   mask  (B, W)     uint64 words, W = ceil(4N/64), bit leg*N + k = 1 for stance (gait.py:26-37).
 */
#ifndef CMPC_H
#define CMPC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cmpc_handle cmpc_handle;

/* solver status per QP (values follow OSQP's constants) */
#define CMPC_SOLVED               1
#define CMPC_SOLVED_INACCURATE    2
#define CMPC_MAX_ITER_REACHED    -2
#define CMPC_NON_CVX             -7
#define CMPC_TOO_MANY_FEET       -20

/* solver modes */
#define CMPC_MODE_ADMM            0   /* OSQP-equivalent ADMM only (reference OPTS: polish off)      */
#define CMPC_MODE_ACTIVE_SET      1   /* exact active-set (polish-first) with ADMM fallback (default) */

/* per-QP statistics record, (B, CMPC_NSTAT) doubles */
#define CMPC_NSTAT 8
#define CMPC_STAT_RPRIM   0   /* max constraint violation (N)                 */
#define CMPC_STAT_RDUAL   1   /* |H u + g + A'y|_inf                          */
#define CMPC_STAT_OBJ     2   /* 1/2 w'Hw + g'w of the reference QP ("cost")    */
#define CMPC_STAT_NFREE   3   /* number of stance (free) force variables      */
#define CMPC_STAT_NACT    4   /* active inequality rows at the solution       */
#define CMPC_STAT_RHO     5   /* final ADMM rho                               */
#define CMPC_STAT_ASITERS 6   /* active-set iterations                        */
#define CMPC_STAT_PATH    7   /* 0 unconstrained, 1 active-set, 2 ADMM, 3 ADMM+polish, 4 unconstrained via a Riccati sweep, 5 active set via Riccati sweeps (pre-pass 4) */

/* Replaces CentroidalMPC.__init__ (centroidal_mpc.py:41-67): allocates nothing on the device but
 * fixes the horizon N (16, 32 or 48 ...; <= 48) and the largest batch the handle will see.     */
int cmpc_create(int N, int max_batch, int device, cmpc_handle** out);
int cmpc_destroy(cmpc_handle* h);

/* Module constants of centroidal_mpc.py:12-38 (COST_MATRIX_Q/R diagonals, MU, fz_min :127, OPTS). */
int cmpc_set_params(cmpc_handle* h, const double Q[12], const double R[12], double mu, double fz_min,
                    double eps_abs, double eps_rel, int max_iter, double rho0, double sigma,
                    double alpha, int mode, int polish, int check_termination,
                    int adaptive_rho_interval);

/* Upper bound on stance foot-steps per robot (<= 4N, default 4N).  The reference's QP always has
 * 12N force variables with swing forces pinned to zero by lbx = ubx = 0 (centroidal_mpc.py:150-161);
 * the fused solver eliminates them, and a tighter bound (e.g. 4*(floor(duty*N)+1) for a periodic
 * gait) lets two CTAs share one SM.  Robots exceeding it get status CMPC_TOO_MANY_FEET.          */
int cmpc_set_max_stance(cmpc_handle* h, int nfmax);

/* Diagnostics: on != 0 routes raw-input calls through the generic kernel (the one used when the
 * caller supplies Ad/Bd/gd of arbitrary structure) instead of the closed-form fast kernel, so the
 * two implementations can be compared on the same inputs.                                        */
int cmpc_set_generic(cmpc_handle* h, int on);

/* Pre-pass (active-set mode with raw inputs) ahead of the condensed active-set kernel; robots the pre-pass cannot
 * finish reach that kernel through a device work-list.  Same optimum either way.
 *   0  off: every robot through the condensed kernel
 *   1  Riccati sweep, one robot per warp (cmpc_riccati.cuh): finishes robots whose unconstrained minimiser is feasible
 *   2  the same, register-resident, two robots per warp (cmpc_riccati2.cuh)
 *   3  = 2 with the sixteen robots of a CTA in lock-step (falls back to 2 when they do not fit shared memory, N > 16)
 *   4  (default) wrench-space projected Riccati + primal-dual active set (cmpc_wrench.cuh), four threads per robot:
 *      finishes constrained robots too (one Riccati sweep per working set, O(N) in the horizon); only robots whose
 *      working set cycles go on to the condensed kernel
 * Batches below 2 048 robots skip the pre-pass (they are latency-bound; the extra launch only adds latency).  */
int cmpc_set_prepass(cmpc_handle* h, int on);

/* Measurement aid (bench.py, SURVEY.md section 8d: per-kernel roofline fractions measured live): with profiling on,
 * cmpc_solve (raw inputs) records CUDA events on its stream before the pre-pass kernel, between the two kernels and
 * after the condensed kernel; cmpc_last_kernel_ms waits for the last such solve and returns the two durations.
 * The reference's counterpart are the wall-clock timers of centroidal_mpc.py:102-105. */
int cmpc_set_profile(cmpc_handle* h, int on);
int cmpc_last_kernel_ms(cmpc_handle* h, double* prepass_ms, double* condensed_ms);
/* The same with the pre-pass split into its sweep kernel and its certificate kernel (route 4; otherwise certificate_ms = 0). */
int cmpc_last_kernel_ms3(cmpc_handle* h, double* sweep_ms, double* certificate_ms, double* condensed_ms);

/* Workspace (section 8b "no allocation inside solve").  A handle keeps four SLOTS of device workspace (work-list,
 * counters, gain scratch, L2-resident scratch of the condensed kernel); consecutive cmpc_solve / cmpc_build calls
 * rotate over them, so AT MOST FOUR calls of one handle may be in flight at a time (on any streams), and calls on
 * one handle must come from one host thread.  cmpc_reserve sizes the slots for batches of up to B robots with the
 * current stance bound; after it, cmpc_solve allocates nothing and may be captured in a CUDA graph.  Without it the
 * first cmpc_solve reserves for max_batch robots (device-synchronising; not during capture).  A later
 * cmpc_set_max_stance or a larger batch re-reserves on the next call.  cmpc_workspace_bytes reports what
 * cmpc_reserve(h, B) holds.                                                                                   */
int cmpc_workspace_bytes(cmpc_handle* h, int B, size_t* bytes);
int cmpc_reserve(cmpc_handle* h, int B);

/* ComTraj.generate_traj (com_trajectory.py:27-211 with gait.py:21-24,40-74), batched: reference trajectory,
 * clamp of the world position target and lever arms CoM->foot over the horizon -- the producer of x_ref and
 * r_foot for cmpc_solve.  All arrays device, FP64, row-major:
 *   x0 (B,12) [p, rpy, v, omega];  R_world_to_body (B,3,3) of the real robot (go2_robot_data.py:216);
 *   foot_lever (B,4,3) measured levers, legs FL FR RL RR (go2_robot_data.py:261-269);
 *   cmd (B,4) = x_vel_des_body, y_vel_des_body, z_pos_des_body, yaw_rate_des_body;  t0 (B) = time_now;
 *   hip_offset (4,3) host, body frame (go2_robot_data.py:147-161);  pos_des_in/out (B,3): the target ComTraj
 *   carries from cycle to cycle (com_trajectory.py:13,47-61), may be the same buffer;
 *   x_ref (B,12,N), r_foot (B,4,3,N) out.  The contact table for the QP is cmpc_contact_table (gait.py:26-37). */
int cmpc_generate_traj(int device, int N, int B, const double* x0, const double* R_world_to_body,
                       const double* foot_lever, const double* cmd, const double* t0, double dt, double gait_hz,
                       double duty, const double phase_offset[4], const double hip_offset[12],
                       const double* pos_des_in, double* pos_des_out, double* x_ref, double* r_foot, void* stream);

/* Single-rigid-body closed-loop step (stand-in for MuJoCo + Pinocchio between two MPC
 * cycles, model of com_trajectory.py:234-270 held for T seconds under the first-step forces, exact ZOH).
 *   x (B,12), u (B,12N) as cmpc_solve wrote it (entries 0..11 = step 0), x_ref (B,12,N), r_foot (B,4,3,N),
 *   I_world (B,3,3), mass (B): the inputs of the cycle just solved;  I_body (3) principal inertias and
 *   stance_offset (4,3) nominal foot positions in the body frame: host.
 * Outputs (device): x_out (B,12), R_world_to_body_out (B,3,3), I_world_out (B,3,3), foot_lever_out (B,4,3) --
 * what cmpc_generate_traj / cmpc_solve read from the robot model for the next cycle.  x_out may alias x and
 * I_world_out may alias I_world.                                                                        */
int cmpc_srb_step(int device, int N, int B, const double* x, const double* u, const double* x_ref,
                  const double* r_foot, const double* I_world, const double* mass, double T, const double I_body[3],
                  const double stance_offset[12], double* x_out, double* R_world_to_body_out, double* I_world_out,
                  double* foot_lever_out, void* stream);

/* Analytic Go2 leg kinematics: the world-aligned 3 x 3 translational foot Jacobians over each
 * leg's hip / thigh / calf joints that compute_3x3_foot_Jacobian_world (go2_robot_data.py:286-300) reads out of Pinocchio,
 * from the joint angles and the base orientation -- the input of cmpc_stance_torque.  Device arrays:
 *   q_joint (B,12) joint angles, legs FL FR RL RR x (hip, thigh, calf);  R_world_to_body (B,3,3) row-major
 *   (go2_robot_data.py:216);  link (3) host = abduction offset l1, thigh length l2, calf length l3 (Go2: 0.0955, 0.213,
 *   0.213 m);  J_foot_world (B,4,3,3) out, row-major, column j = joint j;  foot_pos_body (B,4,3) out or NULL: foot
 *   relative to its hip in the body frame.  The URDF is not in the reference tree: the chain is the published Go2
 *   geometry (hip about x, thigh and calf about y), checked against finite differences of its own forward kinematics. */
int cmpc_leg_jacobian(int device, int B, const double* q_joint, const double* R_world_to_body, const double link[3],
                      double* J_foot_world, double* foot_pos_body, void* stream);

/* Stance torque mapping, the step after the path: for every leg in stance at time_now
 * (Gait.compute_current_mask, gait.py:21-24, bit-exact)  tau = clip(J^T (-f), -tau_max, tau_max)  with f the
 * first-step force of the MPC (leg_controller.py:100-101, test_MPC.py:196,227); swing legs get 0 (their torque is
 * the swing-leg controller's, leg_controller.py:66-98, not on this path).  Device arrays:
 *   J_foot_world (B,4,3,3) row-major world-aligned translational foot Jacobians over the leg's three joints
 *   (go2_robot_data.py:286-300);  u (B,12N) as cmpc_solve wrote it;  time_now (B);  tau (B,12) out;
 *   mask_now (B,4) int32 out, may be NULL.                                                               */
int cmpc_stance_torque(int device, int N, int B, const double* J_foot_world, const double* u, const double* time_now,
                       double gait_hz, double duty, const double phase_offset[4], double tau_max, double* tau,
                       int32_t* mask_now, void* stream);

/* Gait.compute_contact_table (gait.py:26-37), bit-exact.  t0 (B) device; mask_out (B, W) device. */
int cmpc_contact_table(cmpc_handle* h, int B, const double* t0, double dt, double gait_hz,
                       double duty, const double phase_offset[4], uint64_t* mask_out, void* stream);

/* traj.contact_table as the reference stores it, (B,4,N) int32 with 1 = stance
 * (com_trajectory.py:106), packed into the mask words the solver reads.                          */
int cmpc_pack_contact(cmpc_handle* h, int B, const int32_t* table, uint64_t* mask_out, void* stream);

/* ComTraj._continuousDynamics + _discreteDynamics (com_trajectory.py:221-286), closed form.
 * x_ref (B,12,N), r_foot (B,4,3,N), I_world (B,3,3), mass (B)  ->  Ad (B,12,12), Bd (B,N,12,12),
 * gd (B,12).                                                                                     */
int cmpc_dynamics(cmpc_handle* h, int B, const double* x_ref, const double* r_foot,
                  const double* I_world, const double* mass, double dt,
                  double* Ad, double* Bd, double* gd, void* stream);

/* Condensed QP data of centroidal_mpc.py:235-303 after eliminating the dynamics rows:
 * H = 2(Bqp' L Bqp + K) (B,12N,12N) dense symmetric, g (B,12N), for the full variable set.
 * Either (Ad,Bd,gd) are given (drop-in path: traj carries them) or they are NULL and the raw
 * inputs (r_foot, I_world, mass, dt) are used.  Diagnostic / parity entry: the fused solver never
 * writes H to HBM.                                                                              */
int cmpc_build(cmpc_handle* h, int B, const double* Ad, const double* Bd, const double* gd,
               const double* x0, const double* x_ref, const double* r_foot, const double* I_world,
               const double* mass, double dt, double* H, double* g, void* stream);

/* CentroidalMPC.solve_QP (centroidal_mpc.py:69-120): build + solve, fused, one CTA per robot.
 * Inputs as in cmpc_build plus mask (B,W).  In/out warm-start state: u (B,12N), y (B,28N), rho (B)
 * (warm != 0 -> use them as the initial guess, as centroidal_mpc.py:92-95 does; warm = 2: the working set of the
 * previous solution shifted by one horizon stage, SURVEY.md section 8 f4 -- the reference does not shift, :108-110).
 * Optional outputs (may be NULL): X (B,12N), nu (B,12N).  status (B) int32, iters (B) int32,
 * stats (B, CMPC_NSTAT).                                                                         */
int cmpc_solve(cmpc_handle* h, int B, const double* Ad, const double* Bd, const double* gd,
               const double* x0, const double* x_ref, const double* r_foot, const double* I_world,
               const double* mass, double dt, const uint64_t* mask, int warm,
               double* u, double* y, double* rho, double* X, double* nu,
               int32_t* status, int32_t* iters, double* stats, void* stream);

/* Streams: cmpc_solve enqueues on the caller's stream; with the Riccati route (batches >= 2 048) it also forks an internal
 * stream per workspace slot (the certificate kernel runs there while the condensed kernel serves the hand-overs on the
 * caller's stream) and joins it back before returning -- everything stays ordered with respect to the caller's stream, and
 * the fork / join is captured with the rest when the caller's stream is being captured into a CUDA graph. */

/* Same call with HOST buffers (pinned or pageable): chunked H2D -> contact table -> solve -> D2H,
 * copies overlapped with compute on two streams.  Raw-input path only (Ad/Bd computed on device).
 * t0 (B) host.  Outputs u (B,12N), status (B), iters (B) on the host; warm-start state stays
 * resident on the device inside the handle between calls (warm != 0 reuses it).                 */
int cmpc_solve_host(cmpc_handle* h, int B, const double* x0, const double* x_ref,
                    const double* r_foot, const double* I_world, const double* mass,
                    const double* t0, double dt, double gait_hz, double duty,
                    const double phase_offset[4], int warm,
                    double* u, int32_t* status, int32_t* iters);

/* One whole MPC cycle from the robot's state and command on the HOST: what test_MPC.py:173-196 does per cycle --
 * ComTraj.generate_traj (com_trajectory.py:27-211), the contact table (gait.py:26-37), solve_QP
 * (centroidal_mpc.py:69-120) and the slice U_opt = w[12N:] (test_MPC.py:189-192) -- with the reference trajectory and
 * lever arms generated on the device: 408 bytes in per robot (x0 12, R_world_to_body 9, foot_lever 12, cmd 4, t0 1,
 * pos_des 3, I_world 9, mass 1 doubles) instead of the 3.3 KB record of cmpc_solve_host.  pos_des (B,3) is read and
 * written back (the clamped world position target ComTraj carries from cycle to cycle, com_trajectory.py:47-61).
 * first_step_only != 0: u is (B,12) = U_opt[:, 0], the only column the consumer applies (test_MPC.py:196), else
 * (B,12N).  Arrays as in cmpc_generate_traj / cmpc_solve_host; all pointers host.                              */
int cmpc_cycle_host(cmpc_handle* h, int B, const double* x0, const double* R_world_to_body, const double* foot_lever,
                    const double* cmd, const double* t0, double* pos_des, const double* I_world, const double* mass,
                    double dt, double gait_hz, double duty, const double phase_offset[4], const double hip_offset[12],
                    int warm, int first_step_only, double* u, int32_t* status, int32_t* iters);

/* Per-QP statistics (B, CMPC_NSTAT) of the last cmpc_solve_host / cmpc_cycle_host call, copied to the host. */
int cmpc_host_stats(cmpc_handle* h, int B, double* stats_host);

/* Number of kernels launched by this library since process start (bench.py "gpu_launches"). */
long long cmpc_launch_count(void);

/* Micro-benchmarks used for the roofline denominators that MEASURED_PEAKS.json lacks
 * (SURVEY.md section 8d): FP64 FMA throughput (TFLOP/s) and shared-memory read bandwidth (GB/s). */
int cmpc_microbench(int device, double* fp64_tflops, double* smem_gbs);

/* FP64 tensor-core (DMMA m8n8k4) throughput (TFLOP/s) and dependent-issue latency (cycles). */
int cmpc_microbench_dmma(int device, double* dmma_tflops, double* dmma_latency_cycles);

/* Single-warp dependent-issue latencies in cycles: out4 = {DFMA, rsqrt(double)+DADD, 64-bit shuffle, LDS.64+F2I}. */
int cmpc_microbench_latency(int device, double* out4);

const char* cmpc_last_error(void);
const char* cmpc_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CMPC_H */
