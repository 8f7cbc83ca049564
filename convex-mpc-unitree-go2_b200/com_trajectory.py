"""Batched, GPU-resident stand-in for the reference's ``ComTraj`` (``convex_mpc/com_trajectory.py:8-211``):
the producer of the MPC's inputs.  ``generate_traj`` keeps the reference's argument list;
the robot model argument becomes a plain state record with a leading batch dimension because Pinocchio is only
used there to place a joint-less floating base (go2_robot_data.py:224-248).

    gait = Gait(3.0, 0.6)
    traj = ComTraj(state, hip_offset=HIP, device="cuda:0")
    traj.generate_traj(state, gait, time_now, vx_body, vy_body, z_des, yaw_rate, time_step)
    sol = CentroidalMPC(None, traj).solve_QP(None, traj)

All arithmetic runs in ``cmpc_generate_traj`` (csrc/cmpc_traj.cuh); PyTorch only owns the buffers.
"""
import ctypes
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from .centroidal_mpc import PHASE_OFFSET, BatchedComTraj


@dataclass
class Gait:
    """gait.py:11-19: frequency, duty and the derived periods."""
    gait_hz: float
    gait_duty: float

    @property
    def gait_period(self):
        return 1 / self.gait_hz

    @property
    def stance_time(self):
        return self.gait_duty * self.gait_period

    @property
    def swing_time(self):
        return (1 - self.gait_duty) * self.gait_period


@dataclass
class RobotState:
    """What ``generate_traj`` reads from ``PinGo2Model`` (com_trajectory.py:37-40,71,116,125; go2_robot_data.py):
    x (B,12) = compute_com_x_vec(); R_world_to_body (B,3,3); foot_lever_world (B,4,3) legs FL FR RL RR;
    mass (B,) and inertia (B,3,3) = data.Ig.mass / data.Ig.inertia."""
    x: torch.Tensor
    R_world_to_body: torch.Tensor
    foot_lever_world: torch.Tensor
    mass: torch.Tensor
    inertia: torch.Tensor


def _dev(a, device, shape):
    t = a if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a, dtype=np.float64))
    t = t.to(device=device, dtype=torch.float64).reshape(shape).contiguous()
    return t


class ComTraj(BatchedComTraj):
    def __init__(self, state, *, hip_offset, device=None, phase_offset=PHASE_OFFSET):
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.type != "cuda":
            raise _lib.CmpcError("ComTraj.generate_traj runs on a CUDA device only (no CPU fallback)")
        self._lib = _lib.load()
        x = _dev(state.x, self.device, (-1, 12))
        self.B = x.shape[0]
        self.pos_des_world = x[:, 0:3].clone()                     # com_trajectory.py:10-13
        self.hip_offset = np.ascontiguousarray(hip_offset, dtype=np.float64).reshape(4, 3)
        self.phase_offset = tuple(phase_offset)
        self.N = None

    def generate_traj(self, state, gait, time_now, x_vel_des_body, y_vel_des_body, z_pos_des_body,
                      yaw_rate_des_body, time_step, stream=None):
        """com_trajectory.py:27-211, batched.  Scalars broadcast over the batch; ``time_now`` and the four commands
        may be per-robot tensors."""
        B, dev = self.B, self.device
        N = int(gait.gait_period / time_step)                      # com_trajectory.py:66
        x0 = _dev(state.x, dev, (B, 12))
        R_wb = _dev(state.R_world_to_body, dev, (B, 3, 3))
        lever = _dev(state.foot_lever_world, dev, (B, 4, 3))

        def per_robot(v):
            t = v if isinstance(v, torch.Tensor) else torch.as_tensor(np.asarray(v, dtype=np.float64))
            return t.to(device=dev, dtype=torch.float64).reshape(-1).expand(B) if t.numel() == 1 else t.to(dev, torch.float64).reshape(B)
        cmd = torch.stack([per_robot(x_vel_des_body), per_robot(y_vel_des_body), per_robot(z_pos_des_body),
                           per_robot(yaw_rate_des_body)], dim=1).contiguous()
        t0 = per_robot(time_now).contiguous()
        if self.N != N:
            self._x_ref_buf = torch.empty(B, 12, N, dtype=torch.float64, device=dev)
            self._r_foot_buf = torch.empty(B, 4, 3, N, dtype=torch.float64, device=dev)
        s = stream if stream is not None else torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            _lib.check(self._lib.cmpc_generate_traj(
                dev.index or 0, N, B, x0.data_ptr(), R_wb.data_ptr(), lever.data_ptr(), cmd.data_ptr(), t0.data_ptr(),
                float(time_step), float(gait.gait_hz), float(gait.gait_duty), _lib.darr(self.phase_offset),
                _lib.darr(self.hip_offset.reshape(-1)), self.pos_des_world.data_ptr(), self.pos_des_world.data_ptr(),
                self._x_ref_buf.data_ptr(), self._r_foot_buf.data_ptr(), ctypes.c_void_p(s)))
        self._keep = (x0, R_wb, lever, cmd, t0)                    # inputs stay alive until the stream has run
        self._set_fields(N, x0, state, t0, time_step, gait)
        return self

    def enqueue_generate(self, state, gait, t0, cmd, time_step, stream=None):
        """Lean variant of ``generate_traj`` for device-resident loops / CUDA-graph capture: ``t0`` (B,) and ``cmd``
        (B,4) = [vx_body, vy_body, z_des, yaw_rate] are device tensors the caller owns, ``state`` holds contiguous
        FP64 device tensors; no temporaries are created once the output buffers exist."""
        B, dev = self.B, self.device
        N = int(gait.gait_period / time_step)
        if self.N != N or not hasattr(self, "_x_ref_buf"):
            self._x_ref_buf = torch.empty(B, 12, N, dtype=torch.float64, device=dev)
            self._r_foot_buf = torch.empty(B, 4, 3, N, dtype=torch.float64, device=dev)
        s = stream if stream is not None else torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            _lib.check(self._lib.cmpc_generate_traj(
                dev.index or 0, N, B, state.x.data_ptr(), state.R_world_to_body.data_ptr(), state.foot_lever_world.data_ptr(),
                cmd.data_ptr(), t0.data_ptr(), float(time_step), float(gait.gait_hz), float(gait.gait_duty),
                _lib.darr(self.phase_offset), _lib.darr(self.hip_offset.reshape(-1)), self.pos_des_world.data_ptr(),
                self.pos_des_world.data_ptr(), self._x_ref_buf.data_ptr(), self._r_foot_buf.data_ptr(), ctypes.c_void_p(s)))
        self._set_fields(N, state.x, state, t0, time_step, gait)
        return self

    def _set_fields(self, N, x0, state, t0, time_step, gait):
        dev, B = self.device, self.B
        BatchedComTraj.__init__(self, N, x0, self._x_ref_buf, time_step, m=_dev(state.mass, dev, (B,)),
                                I_com_world=_dev(state.inertia, dev, (B, 3, 3)), r_foot=self._r_foot_buf,
                                time_now=t0, gait_hz=gait.gait_hz, gait_duty=gait.gait_duty,
                                phase_offset=self.phase_offset)


def srb_step(state, traj, u, mpc_period, I_body, stance_offset, out=None, stream=None):
    """Advance every robot by one MPC period with the first-step forces of ``u`` (the solver's (B, 12N) force
    buffer, entries 0..11 = step 0) under the MPC's own single-rigid-body model -- ``cmpc_srb_step``, the
    device-resident stand-in for MuJoCo + Pinocchio.  Returns the next ``RobotState``
    (``out`` is reused when given)."""
    lib = _lib.load()
    dev = traj.device
    B, N = traj.B, traj.N
    x = _dev(state.x, dev, (B, 12))
    if out is None:
        out = RobotState(torch.empty_like(x), torch.empty(B, 3, 3, dtype=torch.float64, device=dev),
                         torch.empty(B, 4, 3, dtype=torch.float64, device=dev), traj.m,
                         torch.empty(B, 3, 3, dtype=torch.float64, device=dev))
    s = stream if stream is not None else torch.cuda.current_stream(dev).cuda_stream
    u = u.reshape(B, 12 * N)
    assert u.is_contiguous() and u.dtype == torch.float64
    with torch.cuda.device(dev):
        _lib.check(lib.cmpc_srb_step(dev.index or 0, N, B, x.data_ptr(), u.data_ptr(), traj.compute_x_ref_vec().data_ptr(),
                                     traj.r_foot.data_ptr(), traj.I_com_world.data_ptr(), traj.m.data_ptr(), float(mpc_period),
                                     _lib.darr(np.asarray(I_body, dtype=np.float64).reshape(3)),
                                     _lib.darr(np.asarray(stance_offset, dtype=np.float64).reshape(12)),
                                     out.x.data_ptr(), out.R_world_to_body.data_ptr(), out.inertia.data_ptr(),
                                     out.foot_lever_world.data_ptr(), ctypes.c_void_p(s)))
    return out


GO2_LINKS = (0.0955, 0.213, 0.213)       # abduction offset, thigh, calf (m): the published Go2 leg geometry


def leg_jacobian(q_joint, R_world_to_body, links=GO2_LINKS, with_foot_pos=False, stream=None):
    """World-aligned translational foot Jacobians (B,4,3,3) over each leg's hip / thigh / calf joints from the joint
    angles ``q_joint`` (B,12; legs FL FR RL RR) and the base orientation -- the analytic counterpart of
    ``PinGo2Model.compute_3x3_foot_Jacobian_world`` (go2_robot_data.py:286-300), ``cmpc_leg_jacobian``.  With
    ``with_foot_pos`` also the feet relative to their hips in the body frame (B,4,3)."""
    lib = _lib.load()
    dev = q_joint.device
    B = q_joint.shape[0]
    q = _dev(q_joint, dev, (B, 12))
    R = _dev(R_world_to_body, dev, (B, 3, 3))
    J = torch.empty(B, 4, 3, 3, dtype=torch.float64, device=dev)
    p = torch.empty(B, 4, 3, dtype=torch.float64, device=dev) if with_foot_pos else None
    s = stream if stream is not None else torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        _lib.check(lib.cmpc_leg_jacobian(dev.index or 0, B, q.data_ptr(), R.data_ptr(), _lib.darr(links), J.data_ptr(),
                                         p.data_ptr() if p is not None else None, ctypes.c_void_p(s)))
    return (J, p) if with_foot_pos else J


def stance_torque(J_foot_world, u, time_now, gait, N, tau_max=45.0, phase_offset=PHASE_OFFSET, stream=None):
    """tau = clip(J^T (-f), +-tau_max) for the legs in stance at ``time_now``, zero for swing legs -- the stance
    branch of ``LegController.compute_leg_torque`` (leg_controller.py:100-101) plus the motor saturation of
    test_MPC.py:227, batched (``cmpc_stance_torque``).  ``J_foot_world`` (B,4,3,3), ``u`` the solver's (B,12N) force
    buffer, ``time_now`` (B,) device tensors.  Returns (tau (B,12), mask_now (B,4) int32)."""
    lib = _lib.load()
    dev = u.device
    B = u.shape[0]
    J = _dev(J_foot_world, dev, (B, 4, 3, 3))
    t0 = _dev(time_now, dev, (B,))
    tau = torch.empty(B, 12, dtype=torch.float64, device=dev)
    mask = torch.empty(B, 4, dtype=torch.int32, device=dev)
    s = stream if stream is not None else torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        _lib.check(lib.cmpc_stance_torque(dev.index or 0, int(N), B, J.data_ptr(), u.reshape(B, 12 * N).data_ptr(), t0.data_ptr(),
                                          float(gait.gait_hz), float(gait.gait_duty), _lib.darr(phase_offset), float(tau_max),
                                          tau.data_ptr(), mask.data_ptr(), ctypes.c_void_p(s)))
    return tau, mask
