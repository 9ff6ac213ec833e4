"""Batched mirror of ``Kinematics`` (robot_gym/controllers/mpc/kinematics.py:4-133).

Same method names; every method takes / returns ``[N, ...]`` CUDA tensors and runs the analytic
sm_100a kernels (closed-form leg IK with Newton clean-up, chain FK, base-frame foot Jacobian)
instead of PyBullet's ``calculateInverseKinematics`` / ``calculateJacobian``.
"""
from __future__ import annotations

import numpy as np
import torch

from robot_gym import cuda as rg


def robot_params_from_description(robot, ctrl_constants=None) -> rg.RobotParams:
    """Fill ``rg_robot_params`` from the reference's constant modules (names unchanged):
    GetCtrlConstants (ghost/ctrl_constants.py:8-41), GetConstants (ghost/constants.py:4-43),
    GetMotorConstants (ghost/motor_constants.py:5-15) and the URDF leg chains; gait / swing /
    estimator arguments as wired in mpc_controller.py:28-45."""
    ctrl = ctrl_constants or robot.GetCtrlConstants()
    const = robot.GetConstants()
    motor = robot.GetMotorConstants()
    chains = robot.leg_chains
    p = rg.RobotParams()
    for l in range(4):
        ch = chains[l]
        for j in range(3):
            for a in range(3):
                p.legs[l].p[j][a] = float(ch["p"][j][a])
                p.legs[l].axis[j][a] = float(ch["axis"][j][a])
            for a in range(9):
                p.legs[l].r[j][a] = float(ch["r"][j][a])
        for a in range(3):
            p.legs[l].toe[a] = float(ch["toe"][a])
            p.hip_positions[l][a] = float(const.DEFAULT_HIP_POSITIONS[l][a])
        p.legs[l].ik_sign_hip = 1.0
        p.legs[l].ik_sign_knee = 1.0
        p.stance_duration[l] = float(ctrl.STANCE_DURATION_SECONDS[l])
        p.duty_factor[l] = float(ctrl.DUTY_FACTOR[l])
        p.initial_leg_phase[l] = float(ctrl.INIT_PHASE_FULL_CYCLE[l])
        p.initial_leg_state[l] = int(ctrl.INIT_LEG_STATE[l])
    for m in range(12):
        p.motor_offset[m] = float(motor.MOTOR_OFFSET[m])
        p.motor_direction[m] = float(motor.MOTOR_DIRECTION[m])
        p.motor_kp[m] = float(motor.MOTOR_POSITION_GAINS[m])
        p.motor_kd[m] = float(motor.MOTOR_VELOCITY_GAINS[m])
    p.contact_detection_phase_threshold = 0.1      # _NOMINAL_CONTACT_DETECTION_PHASE (motion_imitation default)
    p.desired_height = float(ctrl.MPC_BODY_HEIGHT)  # mpc_controller.py:44
    p.foot_clearance = 0.01                        # mpc_controller.py:45
    for a in range(3):
        p.swing_kp[a] = 0.03                       # _KP = (0.01, 0.01, 0.01) * 3
    p.swing_max_clearance = 0.1
    p.velocity_window = 20                         # mpc_controller.py:36
    # IK branches: the one that reproduces the robot's nominal stance (knee sign, ghost/constants.py:8-10)
    rg.calibrate_ik(p, np.asarray(const.INIT_MOTOR_ANGLES, dtype=np.float64))
    return p


class BatchedKinematics:
    def __init__(self, robot_workspace: rg.RobotWorkspace, device):
        self._ws = robot_workspace
        self.device = torch.device(device)

    def _lib(self):
        return rg.load()

    def ComputeMotorAnglesFromFootLocalPosition(self, foot_local_position, leg_mask=None, out=None):
        """kinematics.py:98-133 for all four legs of every env: ``[N,12]`` base-frame foot positions ->
        ``[N,12]`` motor angles ((joint - MOTOR_OFFSET) * MOTOR_DIRECTION).  ``leg_mask`` [N,4] u8
        restricts the computation (others keep the contents of ``out``)."""
        n = foot_local_position.shape[0]
        foot = foot_local_position.reshape(n, 12)
        if out is None:
            out = torch.zeros((n, 12), dtype=torch.float32, device=foot.device)
        rg.check(self._lib().rg_leg_ik(self._ws.ptr, n, rg._ptr(foot, torch.float32, (12,)),
                                       rg._ptr(leg_mask, torch.uint8, (4,), allow_none=True),
                                       rg._ptr(out, torch.float32, (12,)), rg.current_stream_ptr()))
        return out

    def ComputeFootPositionsInBaseFrame(self, motor_angles, out=None):
        """Chain FK -- what Robot.GetFootPositionsInBaseFrame (robot.py:389-397) reads from PyBullet."""
        n = motor_angles.shape[0]
        if out is None:
            out = torch.empty((n, 12), dtype=torch.float32, device=motor_angles.device)
        rg.check(self._lib().rg_leg_fk(self._ws.ptr, n, rg._ptr(motor_angles, torch.float32, (12,)),
                                       rg._ptr(out, torch.float32, (12,)), rg.current_stream_ptr()))
        return out

    def MapContactForceToJointTorques(self, contact_force, motor_angles, out=None):
        """kinematics.py:40-53 for all legs: ``[N,12]`` forces, ``[N,12]`` motor angles -> ``[N,12]`` torques."""
        n = contact_force.shape[0]
        if out is None:
            out = torch.empty((n, 12), dtype=torch.float32, device=contact_force.device)
        rg.check(self._lib().rg_force_to_torque(self._ws.ptr, n, rg._ptr(contact_force.reshape(n, 12), torch.float32, (12,)),
                                                rg._ptr(motor_angles, torch.float32, (12,)),
                                                rg._ptr(out, torch.float32, (12,)), rg.current_stream_ptr()))
        return out
