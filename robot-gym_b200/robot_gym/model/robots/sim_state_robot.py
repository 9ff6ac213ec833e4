"""Batched state provider fed by RAW rigid-body state ("next" row f2 of SURVEY.md section 8).

``Robot`` (robot_gym/model/robots/robot.py) answers the controller's getters with PyBullet queries:
``getEulerFromQuaternion`` (:79-86), ``invertTransform`` / ``multiplyTransforms`` (:185-213), joint
states (:231-236) and four ``getLinkState`` round trips per call (:367-397).  A batched simulator has
none of those calls -- it has tensors: base orientation, base twist, joint angles, contact flags.
``SimStateRobotBatch`` turns them into the getter surface ``BatchedMPCController`` reads with ONE kernel
launch (``rg_state_from_sim``); contacts, base velocity and orientation pass through untouched.
"""
from __future__ import annotations

import torch

from robot_gym import cuda as rg
from robot_gym.controllers.mpc.batched_kinematics import robot_params_from_description


class SimStateRobotBatch:
    def __init__(self, description, num_envs, device="cuda", robot_workspace=None):
        self.description = description
        self.device = torch.device(device)
        self.num_envs = int(num_envs)
        self.num_legs = 4
        self.num_motors = 12
        self._ws = robot_workspace or rg.RobotWorkspace(robot_params_from_description(description), device=self.device)
        n, dev = self.num_envs, self.device
        f32 = dict(dtype=torch.float32, device=dev)
        self.time_since_reset = torch.zeros(n, dtype=torch.float64, device=dev)
        self.foot_contacts = torch.ones((n, 4), dtype=torch.uint8, device=dev)
        self.base_velocity_world = torch.zeros((n, 3), **f32)
        self.base_orientation_xyzw = torch.zeros((n, 4), **f32)
        self.base_orientation_xyzw[:, 3] = 1.0
        self.base_rpy = torch.zeros((n, 3), **f32)
        self.base_rpy_rate = torch.zeros((n, 3), **f32)
        self.foot_positions_base = torch.zeros((n, 4, 3), **f32)
        self.motor_angles = torch.zeros((n, 12), **f32)

    def set_sim_state(self, time_since_reset, base_orientation_xyzw, base_velocity_world, base_angular_velocity_world,
                      joint_angles, foot_contacts):
        """One simulator step's worth of raw state (CUDA tensors, ``[N, ...]``); derived getters are refreshed
        by a single launch.  Tensors are referenced, not copied, where the getter is a pass-through."""
        n = self.num_envs
        self.time_since_reset = time_since_reset
        self.base_orientation_xyzw = base_orientation_xyzw
        self.base_velocity_world = base_velocity_world
        self.foot_contacts = foot_contacts
        rg.check(rg.load().rg_state_from_sim(
            self._ws.ptr, n, rg._ptr(base_orientation_xyzw, torch.float32, (4,)),
            rg._ptr(base_angular_velocity_world, torch.float32, (3,)), rg._ptr(joint_angles, torch.float32, (12,)),
            rg._ptr(self.base_rpy, torch.float32, (3,)), rg._ptr(self.base_rpy_rate, torch.float32, (3,)),
            rg._ptr(self.motor_angles, torch.float32, (12,)), rg._ptr(self.foot_positions_base.view(n, 12), torch.float32, (12,)),
            rg.current_stream_ptr()))

    # ---- description passthrough (ghost/ghost.py:7-30)
    def GetCtrlConstants(self):
        return self.description.GetCtrlConstants()

    def GetConstants(self):
        return self.description.GetConstants()

    def GetMotorConstants(self):
        return self.description.GetMotorConstants()

    @property
    def leg_chains(self):
        return self.description.leg_chains

    # ---- batched getters (robot.py names)
    def GetTimeSinceReset(self):
        return self.time_since_reset

    def GetFootContacts(self):
        return self.foot_contacts

    def GetBaseVelocity(self):
        return self.base_velocity_world

    def GetTrueBaseOrientation(self):
        return self.base_orientation_xyzw

    def GetBaseRollPitchYaw(self):
        return self.base_rpy

    def GetBaseRollPitchYawRate(self):
        return self.base_rpy_rate

    def GetFootPositionsInBaseFrame(self):
        return self.foot_positions_base

    def GetMotorAngles(self):
        return self.motor_angles

    def GetHipPositionsInBaseFrame(self):
        return self.GetConstants().DEFAULT_HIP_POSITIONS

    def GetMotorPositionGains(self):
        return self.GetMotorConstants().MOTOR_POSITION_GAINS

    def GetMotorVelocityGains(self):
        return self.GetMotorConstants().MOTOR_VELOCITY_GAINS
