"""N = 1 adapter: a stock ``Robot`` (robot_gym/model/robots/robot.py) behind the batched getter surface.

``Simulation.build_world`` constructs the controller as ``controller_class(robot, sim.GetTimeSinceReset)``
(core/simulation.py:117) with the PyBullet-backed ``Robot``.  Its getters answer with python lists / numpy arrays
(one PyBullet query each: robot.py:79-86, :172-183, :205-236, :389-397); the kernels want ``[1, ...]`` CUDA tensors.
``PyBulletRobotAdapter`` calls the same getters the third-party stack called, packs the 35 floats + 4 contact bytes into
ONE pinned host buffer and moves them with ONE host->device copy per control step, then hands out views.
``BatchedMPCController`` wraps any robot that has no ``num_envs`` attribute in this adapter, so registering it as
``'mpc_cuda'`` in robot_gym/util/cli/mapper.py makes it selectable in the playground with no other change.

PyBullet itself is absent offline: the tests drive the adapter with a stub exposing the getter names."""
from __future__ import annotations

import numpy as np
import torch

from robot_gym.model.robots import descriptions


def _description_for(robot):
    """The leg chains / constants bundle of a stock robot: by its URDF name (ghost/constants.py: URDF_FILE) or class name."""
    name = type(robot).__name__.lower()
    try:
        name = str(robot.GetConstants().URDF_FILE).lower()
    except Exception:
        pass
    for key, desc in descriptions.ROBOTS.items():
        if key in name:
            return desc
    raise ValueError(f"no leg-chain description for robot {type(robot).__name__!r}: pass description=")


class PyBulletRobotAdapter:
    # (attribute, getter on the stock robot, number of floats)
    # foot positions first: the solve kernel reads them with 128-bit loads (16-byte aligned)
    _FLOAT_FIELDS = (("foot_positions_base", "GetFootPositionsInBaseFrame", 12), ("motor_angles", "GetMotorAngles", 12),
                     ("base_orientation_xyzw", "GetTrueBaseOrientation", 4), ("base_velocity_world", "GetBaseVelocity", 3),
                     ("base_rpy", "GetBaseRollPitchYaw", 3), ("base_rpy_rate", "GetBaseRollPitchYawRate", 3))

    def __init__(self, robot, description=None, device="cuda"):
        self._robot = robot
        self.description = description or _description_for(robot)
        self.device = torch.device(device)
        self.num_envs, self.num_legs, self.num_motors = 1, 4, 12
        n_float = sum(k for _, _, k in self._FLOAT_FIELDS)
        self._host = torch.zeros(n_float * 4 + 4, dtype=torch.uint8).pin_memory()
        self._dev = torch.zeros(n_float * 4 + 4, dtype=torch.uint8, device=self.device)
        self._host_f = self._host[:n_float * 4].view(torch.float32).numpy()
        self._host_c = self._host[n_float * 4:].numpy()
        dev_f = self._dev[:n_float * 4].view(torch.float32)
        off = 0
        for attr, _, k in self._FLOAT_FIELDS:
            setattr(self, attr, dev_f[off:off + k].view(1, k))
            off += k
        self.foot_contacts = self._dev[n_float * 4:].view(1, 4)
        self.refresh()

    def refresh(self):
        """Query the stock robot once per getter and move everything with one pinned copy (stream-ordered)."""
        off = 0
        for _, getter, k in self._FLOAT_FIELDS:
            self._host_f[off:off + k] = np.asarray(getattr(self._robot, getter)(), dtype=np.float32).reshape(-1)
            off += k
        self._host_c[:] = np.asarray(self._robot.GetFootContacts(), dtype=np.uint8).reshape(-1)
        self._dev.copy_(self._host, non_blocking=True)

    # ---- constants: straight from the stock robot (names unchanged: ghost/ghost.py:7-30)
    def GetCtrlConstants(self):
        return self._robot.GetCtrlConstants()

    def GetConstants(self):
        return self._robot.GetConstants()

    def GetMotorConstants(self):
        return self._robot.GetMotorConstants()

    @property
    def leg_chains(self):
        return self.description.leg_chains

    # ---- batched getters
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
