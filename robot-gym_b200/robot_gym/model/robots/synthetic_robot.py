"""A batched robot-state provider with robot-gym's getter names, backed by CUDA tensors.

The third-party MPC stack pulls its inputs by calling robot getters
(robot_gym/model/robots/robot.py:79-92,169-236,389-397).  ``BatchedMPCController`` reads the same
getters, but each returns an ``[N, ...]`` CUDA tensor for N independent envs.  This class is the
synthetic source used by tests and bench.py (no physics -- PyBullet is out of scope); a
PyBullet-backed N=1 adapter only has to return the same shapes.
"""
from __future__ import annotations

import torch

from robot_gym.util.synthetic import SyntheticStates


class SyntheticRobotBatch:
    def __init__(self, description, states: SyntheticStates, device="cuda"):
        self.description = description
        self.device = torch.device(device)
        self.num_envs = len(states)
        self.num_legs = 4
        self.num_motors = 12
        self.load(states)

    _FIELDS = ("time_since_reset", "foot_contacts", "base_velocity_world", "base_orientation_xyzw", "base_rpy",
               "base_rpy_rate", "foot_positions_base", "motor_angles")

    def load(self, states: SyntheticStates, non_blocking=False):
        """Host -> device copy of a full state batch (pinned staging makes it asynchronous).  Existing device tensors
        are overwritten IN PLACE, so the storage the controller (and a captured CUDA graph) points at stays put."""
        dev = self.device
        for name in self._FIELDS:
            src = torch.from_numpy(getattr(states, name))
            cur = getattr(self, name, None)
            if isinstance(cur, torch.Tensor) and cur.shape == src.shape and cur.dtype == src.dtype and cur.device.type == dev.type:
                cur.copy_(src, non_blocking=non_blocking)
            else:
                setattr(self, name, src.to(dev, non_blocking=non_blocking))

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
