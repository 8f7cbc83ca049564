// cmpc_traj.cuh -- batched ComTraj.generate_traj: the producer of the hot path's inputs.
//
// Replaces, per robot and per MPC cycle (reference: ltinphan/convex-mpc-unitree-go2, convex_mpc/):
//   com_trajectory.py:44-61    clamp of the world position target to +-0.1 m around the CoM, z from the command
//   com_trajectory.py:64-103   reference trajectory: p_des + v_world t, yaw + yaw_rate t, constant v_world / yaw rate
//   com_trajectory.py:108-201  lever arms CoM->foot over the horizon: take-off / touch-down state machine on the
//                              gait mask sampled at time_now + i dt (gait.py:21-24: NO half-step offset, unlike the
//                              contact table of gait.py:26-37), zero while in swing
//   gait.py:40-74              touchdown prediction at take-off (hip under the yawed base + drift + yaw correction),
//                              with the reference's quirk that the drift uses the BODY-frame velocity
//                              (com_trajectory.py:125-131 -> gait.py:42,58)
// Pinocchio is only used there to place a joint-less floating base (go2_robot_data.py:224-248), so the kinematics
// reduce to closed form; the hip offsets (go2_robot_data.py:147-161) are constants passed in by the caller.
//
// One thread per (robot, leg): the thread of leg l writes rows 3l..3l+2 of x_ref ([p; rpy; v; omega]) and the
// three lever-arm rows of its leg -- every row is N contiguous doubles.  HBM-bound and tiny: ~0.4 KB in,
// 3 KB out per robot.
#pragma once
#include "cmpc_core.cuh"

namespace cmpc {
namespace traj {

// gait mask at time_now + i dt as compute_current_mask forms it (gait.py:21-24 -> 26-37 with dt = 0, N = 1):
//   t = time_now + i*time_step   (com_trajectory.py:120);  inside: t + arange(1)*0 + 0/2 leaves t unchanged
CMPC_HD int stance_bit_now(double t0, double dt, int i, double period, double offset, double duty) {
    const double t = dadd_rn(t0, dmul_rn((double)i, dt));
    const double ph = dadd_rn(offset, ddiv_rn(t, period));
    double r = fmod(ph, 1.0);
    if (r != 0.0 && r < 0.0) r = dadd_rn(r, 1.0);
    return r < duty ? 1 : 0;
}

// pos_des after the clamp (com_trajectory.py:44-61)
CMPC_HD void clamp_pos_des(const double* x0, const double* pos_des_in, double z_des, double out[3]) {
    const double max_err = 0.1;
    for (int a = 0; a < 2; ++a) {
        double v = pos_des_in[a];
        if (v - x0[a] > max_err) v = x0[a] + max_err;
        if (x0[a] - v > max_err) v = x0[a] - max_err;
        out[a] = v;
    }
    out[2] = z_des;
}

// One (robot, leg).  x_ref: the robot's (12, N) block; r_leg: the leg's (3, N) block; pos_des_out may be null.
CMPC_HD void generate_leg(int N, int leg, const double* x0, const double* R_wb, const double* lever, const double* cmd,
                          double t0, double dt, double period, double duty, double offset, const double* hip,
                          const double* pos_des_in, double* pos_des_out, double* x_ref, double* r_leg) {
    double pd[3];
    clamp_pos_des(x0, pos_des_in, cmd[2], pd);
    const double yaw = x0[5], yaw_rate = cmd[3];
    const double c0 = cos(yaw), s0 = sin(yaw);
    const double vw[3] = {c0 * cmd[0] - s0 * cmd[1], s0 * cmd[0] + c0 * cmd[1], 0.0};   // R_z [vx, vy, 0]
    // rows 3 leg .. 3 leg + 2 of the reference (com_trajectory.py:84-103); column i <-> time (i+1) dt
    for (int i = 0; i < N; ++i) {
        const double t = ((double)i + 1.0) * dt;
        double v0, v1, v2;
        if (leg == 0) { v0 = pd[0] + vw[0] * t; v1 = pd[1] + vw[1] * t; v2 = pd[2] + vw[2] * t; }
        else if (leg == 1) { v0 = 0.0; v1 = 0.0; v2 = yaw + yaw_rate * t; }
        else if (leg == 2) { v0 = vw[0]; v1 = vw[1]; v2 = vw[2]; }
        else { v0 = 0.0; v1 = 0.0; v2 = yaw_rate; }
        x_ref[(3 * leg) * N + i] = v0;
        x_ref[(3 * leg + 1) * N + i] = v1;
        x_ref[(3 * leg + 2) * N + i] = v2;
    }
    if (leg == 0 && pos_des_out) { pos_des_out[0] = pd[0]; pos_des_out[1] = pd[1]; pos_des_out[2] = pd[2]; }
    // lever arms (com_trajectory.py:108-201)
    const double t_swing = (1.0 - duty) * period, t_stance = duty * period;   // gait.py:18-19
    const double pred = (t_swing + 0.5 * t_stance) / 2.0;                     // gait.py:53-54
    // body-frame velocity of the dummy model: R_world_to_body of the REAL robot times the reference velocity
    const double vb0 = R_wb[0] * vw[0] + R_wb[1] * vw[1] + R_wb[2] * vw[2];
    const double vb1 = R_wb[3] * vw[0] + R_wb[4] * vw[1] + R_wb[5] * vw[2];
    double nx = lever[0], ny = lever[1], nz = lever[2];      // next touchdown lever: starts as the measured one (:116)
    double cx = 0.0, cy = 0.0, cz = 0.0;
    int prev = 2;
    for (int i = 0; i < N; ++i) {
        const int m = stance_bit_now(t0, dt, i, period, offset, duty);
        if (m != prev) {
            if (m == 0) {
                const double t = ((double)i + 1.0) * dt;
                const double bx = pd[0] + vw[0] * t, by = pd[1] + vw[1] * t, bz = pd[2] + vw[2] * t;
                const double yi = yaw + yaw_rate * t;
                const double ci = cos(yi), si = sin(yi);
                const double hx = ci * hip[0] - si * hip[1], hy = si * hip[0] + ci * hip[1];
                const double px = bx + hx, py = by + hy;                    // nominal touchdown (hip under the base)
                const double dth = yaw_rate * pred;
                const double rx = px - bx, ry = py - by;
                const double tx = (px + vb0 * pred) + (-dth * ry);
                const double ty = (py + vb1 * pred) + (dth * rx);
                const double tz = 0.02;
                nx = tx - bx; ny = ty - by; nz = tz - bz;
                cx = 0.0; cy = 0.0; cz = 0.0;
            } else {
                cx = nx; cy = ny; cz = nz;
            }
        }
        r_leg[i] = cx;
        r_leg[N + i] = cy;
        r_leg[2 * N + i] = cz;
        prev = m;
    }
}

// ----------------------------------------------------------------------------------------------
// Single-rigid-body closed-loop step: stands in for MuJoCo + Pinocchio between two MPC
// cycles so that generate_traj -> solve -> step stays on the device.  The model is the one the MPC itself uses
// (com_trajectory.py:234-270) held for T seconds under the first-step forces; because A_c^2 = 0 the zero-order
// hold is exact in closed form:  x(T) = x + T A_c x + (T I + T^2/2 A_c)(B_c u + g_c).
// Outputs besides the state: R_world_to_body (ZYX Euler, go2_robot_data.py:213-216), the world inertia
// R I_body R^T, and nominal measured levers (stance offsets under the yawed body, z = -height) -- the three
// things generate_traj / the MPC read from the robot model.
// ----------------------------------------------------------------------------------------------
CMPC_HD void srb_step_one(int N, const double* x, const double* u0, const double* x_ref, const double* r_foot,
                          const double* I_world, double mass, double T, const double* I_body, const double* stance_off,
                          double* x_out, double* R_wb_out, double* I_out, double* lever_out) {
    DynCommon d;
    dyn_common(d, x_ref, N, I_world, mass, T);
    double F[3] = {0.0, 0.0, 0.0}, tq[3] = {0.0, 0.0, 0.0};
    for (int leg = 0; leg < 4; ++leg) {
        const double* f = u0 + 3 * leg;
        const double r[3] = {r_foot[(leg * 3 + 0) * N], r_foot[(leg * 3 + 1) * N], r_foot[(leg * 3 + 2) * N]};
        for (int a = 0; a < 3; ++a) F[a] += f[a];
        tq[0] += r[1] * f[2] - r[2] * f[1];
        tq[1] += r[2] * f[0] - r[0] * f[2];
        tq[2] += r[0] * f[1] - r[1] * f[0];
    }
    double al[3];                                           // angular acceleration I^-1 sum r x f
    for (int i = 0; i < 3; ++i) al[i] = d.Iinv[i * 3] * tq[0] + d.Iinv[i * 3 + 1] * tq[1] + d.Iinv[i * 3 + 2] * tq[2];
    const double acc[3] = {F[0] * d.minv, F[1] * d.minv, F[2] * d.minv - 9.81};
    const double h = T * T / 2.0;
    const double om[3] = {x[9], x[10], x[11]};
    const double rz_om[3] = {d.cy * om[0] + d.sy * om[1], -d.sy * om[0] + d.cy * om[1], om[2]};
    const double rz_al[3] = {d.cy * al[0] + d.sy * al[1], -d.sy * al[0] + d.cy * al[1], al[2]};
    double xn[12];
    for (int a = 0; a < 3; ++a) {
        xn[a] = x[a] + T * x[6 + a] + h * acc[a];
        xn[3 + a] = x[3 + a] + T * rz_om[a] + h * rz_al[a];
        xn[6 + a] = x[6 + a] + T * acc[a];
        xn[9 + a] = x[9 + a] + T * al[a];
    }
    for (int i = 0; i < 12; ++i) x_out[i] = xn[i];
    const double cr = cos(xn[3]), sr = sin(xn[3]), cp = cos(xn[4]), sp = sin(xn[4]), cy = cos(xn[5]), sy = sin(xn[5]);
    const double R[9] = {cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr,
                         sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr,
                         -sp, cp * sr, cp * cr};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            R_wb_out[i * 3 + j] = R[j * 3 + i];
            I_out[i * 3 + j] = R[i * 3] * I_body[0] * R[j * 3] + R[i * 3 + 1] * I_body[1] * R[j * 3 + 1] + R[i * 3 + 2] * I_body[2] * R[j * 3 + 2];
        }
    for (int leg = 0; leg < 4; ++leg) {
        lever_out[3 * leg] = cy * stance_off[3 * leg] - sy * stance_off[3 * leg + 1];
        lever_out[3 * leg + 1] = sy * stance_off[3 * leg] + cy * stance_off[3 * leg + 1];
        lever_out[3 * leg + 2] = -xn[2];
    }
}

// ----------------------------------------------------------------------------------------------
// Stance torque mapping: the step after the path.  For a leg in stance at time_now
// (gait.compute_current_mask, gait.py:21-24)  tau = J^T (-f)  with J the 3x3 world-aligned translational foot
// Jacobian over the leg's three joints (leg_controller.py:100-101, go2_robot_data.py:286-300) and f the
// first-step contact force of the MPC (test_MPC.py:196), then the motor saturation  clip(tau, -tau_max, tau_max)
// (test_MPC.py:227).  Swing legs get zero here: their torque comes from the swing-leg controller
// (leg_controller.py:66-98: operational-space PD with the joint-space inertia), which is not on this path.
// ----------------------------------------------------------------------------------------------
CMPC_HD void stance_torque_leg(const double* J, const double* f, int stance, double tau_max, double* tau) {
    for (int j = 0; j < 3; ++j) {
        double t = 0.0;
        if (stance) {
            t = -(J[0 * 3 + j] * f[0] + J[1 * 3 + j] * f[1] + J[2 * 3 + j] * f[2]);
            t = fmin(fmax(t, -tau_max), tau_max);
        }
        tau[j] = t;
    }
}

// ----------------------------------------------------------------------------------------------
// Analytic leg kinematics of the Go2: the world-aligned translational foot Jacobian over the three
// joints of a leg, what compute_3x3_foot_Jacobian_world (go2_robot_data.py:286-300) reads out of Pinocchio.  Chain of the
// Go2 description: hip (abduction) joint about x at the hip position, thigh joint about y at (0, s l1, 0) (s = +1 left, -1
// right legs), calf joint about y at (0, 0, -l2), foot at (0, 0, -l3); the hip frames are parallel to the body frame, so
//     p = Rx(q1) [ X, s l1, Z ],   X = -l2 sin q2 - l3 sin(q2+q3),   Z = -l2 cos q2 - l3 cos(q2+q3)
// is the foot relative to its hip in the body frame and J_world = R_body_to_world dp/dq.  Rb = R_world_to_body (row-major,
// as go2_robot_data.py:216 hands it out), so R_body_to_world = Rb'.
// ----------------------------------------------------------------------------------------------
CMPC_HD void leg_jacobian(const double* q3, const double* Rb, double side, double l1, double l2, double l3, double* Jw, double* p_body) {
    const double s1 = sin(q3[0]), c1 = cos(q3[0]), s2 = sin(q3[1]), c2 = cos(q3[1]);
    const double s23 = sin(q3[1] + q3[2]), c23 = cos(q3[1] + q3[2]);
    const double X = -l2 * s2 - l3 * s23, Z = -l2 * c2 - l3 * c23, y1 = side * l1;
    if (p_body) { p_body[0] = X; p_body[1] = y1 * c1 - Z * s1; p_body[2] = y1 * s1 + Z * c1; }
    // dp/dq in the body frame, column j = joint j;  dX/dq2 = Z, dZ/dq2 = -X, dX/dq3 = -l3 c23, dZ/dq3 = l3 s23
    double Jb[9];
    Jb[0] = 0.0;                  Jb[1] = Z;        Jb[2] = -l3 * c23;
    Jb[3] = -y1 * s1 - Z * c1;    Jb[4] = X * s1;   Jb[5] = -l3 * s23 * s1;
    Jb[6] = y1 * c1 - Z * s1;     Jb[7] = -X * c1;  Jb[8] = l3 * s23 * c1;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            Jw[i * 3 + j] = Rb[0 * 3 + i] * Jb[0 * 3 + j] + Rb[1 * 3 + i] * Jb[1 * 3 + j] + Rb[2 * 3 + i] * Jb[2 * 3 + j];
}

}  // namespace traj
}  // namespace cmpc
