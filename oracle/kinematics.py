"""Oracle: leg kinematics + a state-holding robot with robot-gym's callback surface.  TEST INFRASTRUCTURE ONLY.

Restates what the reference obtains from PyBullet (third-party C, absent offline):

  * ``Robot.GetFootPositionsInBaseFrame`` (robot_gym/model/robots/robot.py:367-397): forward
    kinematics of the URDF chain, base frame = base link inertial frame;
  * ``Kinematics.MapContactForceToJointTorques`` (robot_gym/controllers/mpc/kinematics.py:13-53):
    ``calculateJacobian`` at the toe link with local position (0,0,0) -> translational Jacobian;
    PyBullet evaluates it with the floating base at the identity pose, i.e. in the base frame;
  * ``Kinematics.ComputeMotorAnglesFromFootLocalPosition`` (:98-133): ``calculateInverseKinematics``
    (solver 0 = damped least squares, seeded from the current joint state).  PyBullet's result is an
    iterate of unspecified tolerance, so IK parity is defined by FK(IK(p)) = p; this oracle iterates
    damped Newton steps to 1e-13 m from the seed pose.  It shares nothing with the CUDA closed form.

PARITY UNPINNED: see oracle/__init__.py.
"""
from __future__ import annotations

import math

import numpy as np


def _rot_axis(axis, q):
    a = np.asarray(axis, dtype=np.float64)
    a = a / np.linalg.norm(a)
    k = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + math.sin(q) * k + (1 - math.cos(q)) * (k @ k)


class LegChain:
    def __init__(self, chain):
        self.p = [np.asarray(v, dtype=np.float64) for v in chain["p"]]
        self.r = [np.asarray(v, dtype=np.float64).reshape(3, 3) for v in chain["r"]]
        self.axis = [np.asarray(v, dtype=np.float64) for v in chain["axis"]]
        self.toe = np.asarray(chain["toe"], dtype=np.float64)

    def fk(self, q, with_jacobian=False):
        """Foot position in the base frame; optionally the 3x3 translational Jacobian."""
        rot = np.eye(3)
        pos = np.zeros(3)
        origins, axes = [], []
        for j in range(3):
            pos = pos + rot @ self.p[j]
            rot = rot @ self.r[j]
            origins.append(pos.copy())
            axes.append(rot @ (self.axis[j] / np.linalg.norm(self.axis[j])))
            rot = rot @ _rot_axis(self.axis[j], q[j])
        foot = pos + rot @ self.toe
        if not with_jacobian:
            return foot
        jac = np.stack([np.cross(axes[j], foot - origins[j]) for j in range(3)], axis=1)
        return foot, jac

    def ik(self, target, seed, iters=100, tol=1e-13):
        """Damped Newton on the chain from ``seed`` (branch follows the seed, as PyBullet's does)."""
        q = np.array(seed, dtype=np.float64)
        target = np.asarray(target, dtype=np.float64)
        lam = 1e-6
        for _ in range(iters):
            foot, jac = self.fk(q, with_jacobian=True)
            err = target - foot
            if np.linalg.norm(err) < tol:
                break
            step = np.linalg.solve(jac.T @ jac + lam * np.eye(3), jac.T @ err)
            nrm = np.linalg.norm(step)
            if nrm > 0.5:
                step *= 0.5 / nrm
            q = q + step
        return (q + math.pi) % (2 * math.pi) - math.pi


def quat_to_rpy(q_xyzw):
    """``pybullet.getEulerFromQuaternion`` (called by Robot.GetBaseRollPitchYaw, robot.py:79-86): Bullet's
    ZYX Euler extraction with its two gimbal-lock branches.  PyBullet is absent: restated from the published
    btQuaternion::getEulerZYX algorithm (parity unpinned); pinned here by the euler -> quaternion -> euler
    round trip and by agreement with the rotation matrix (tests/test_oracle_locomotion.py)."""
    qx, qy, qz, qw = [float(v) for v in q_xyzw]
    sqx, sqy, sqz, sqw = qx * qx, qy * qy, qz * qz, qw * qw
    sarg = -2.0 * (qx * qz - qw * qy)
    if sarg <= -0.99999:
        return np.array([0.0, -0.5 * math.pi, 2.0 * math.atan2(qx, -qy)])
    if sarg >= 0.99999:
        return np.array([0.0, 0.5 * math.pi, 2.0 * math.atan2(-qx, qy)])
    return np.array([math.atan2(2.0 * (qy * qz + qw * qx), sqw - sqx - sqy + sqz), math.asin(sarg),
                     math.atan2(2.0 * (qx * qy + qw * qz), sqw + sqx - sqy - sqz)])


def quat_to_matrix(q_xyzw):
    """Rotation matrix of a Bullet (x, y, z, w) quaternion (getMatrixFromQuaternion)."""
    x, y, z, w = [float(v) for v in q_xyzw]
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def angular_velocity_to_local_frame(angular_velocity_world, q_xyzw):
    """Robot.TransformAngularVelocityToLocalFrame (robot.py:185-203): invertTransform + multiplyTransforms of a
    pure rotation = R(q)^T w."""
    return quat_to_matrix(q_xyzw).T @ np.asarray(angular_velocity_world, dtype=np.float64)


class OracleRobot:
    """Holds one env's state and answers the getters the third-party stack calls (robot.py:71-264)."""

    def __init__(self, description):
        self._constants = description.GetConstants()
        self._motor = description.GetMotorConstants()
        self._chains = [LegChain(c) for c in description.leg_chains]
        self.num_legs = 4
        self.num_motors = 12
        self._offset = np.asarray(self._motor.MOTOR_OFFSET, dtype=np.float64)
        self._direction = np.asarray(self._motor.MOTOR_DIRECTION, dtype=np.float64)
        self.set_state()

    # ---- state injection (synthetic state source)
    def set_state(self, base_velocity=(0, 0, 0), base_orientation=(0, 0, 0, 1), base_rpy=(0, 0, 0),
                  base_rpy_rate=(0, 0, 0), foot_positions=None, foot_contacts=(1, 1, 1, 1), motor_angles=None):
        self._base_velocity = np.asarray(base_velocity, dtype=np.float64)
        self._base_orientation = np.asarray(base_orientation, dtype=np.float64)
        self._base_rpy = np.asarray(base_rpy, dtype=np.float64)
        self._base_rpy_rate = np.asarray(base_rpy_rate, dtype=np.float64)
        self._motor_angles = (np.asarray(self._constants.INIT_MOTOR_ANGLES, dtype=np.float64)
                              if motor_angles is None else np.asarray(motor_angles, dtype=np.float64))
        if foot_positions is None:
            foot_positions = self.fk_all(self._motor_angles)
        self._foot_positions = np.asarray(foot_positions, dtype=np.float64).reshape(4, 3)
        self._foot_contacts = [bool(c) for c in foot_contacts]

    # ---- kinematics helpers
    def joint_angles(self, motor_angles):
        return np.asarray(motor_angles, dtype=np.float64) * self._direction + self._offset

    def fk_all(self, motor_angles):
        q = self.joint_angles(motor_angles)
        return np.stack([self._chains[l].fk(q[3 * l:3 * l + 3]) for l in range(4)])

    def set_sim_state(self, base_orientation, base_velocity, base_angular_velocity_world, joint_angles, foot_contacts):
        """State injection from RAW rigid-body state, through the same conversions Robot's getters apply to
        PyBullet's answers (robot.py:79-86,185-213,231-236,367-397)."""
        joint_angles = np.asarray(joint_angles, dtype=np.float64)
        motor = (joint_angles - self._offset) * self._direction
        self.set_state(base_velocity=base_velocity, base_orientation=base_orientation,
                       base_rpy=quat_to_rpy(base_orientation),
                       base_rpy_rate=angular_velocity_to_local_frame(base_angular_velocity_world, base_orientation),
                       foot_positions=None, foot_contacts=foot_contacts, motor_angles=motor)

    # ---- robot.py callback surface
    def GetFootContacts(self):
        return list(self._foot_contacts)

    def GetBaseVelocity(self):
        return tuple(self._base_velocity)

    def GetTrueBaseOrientation(self):
        return tuple(self._base_orientation)

    def GetBaseRollPitchYaw(self):
        return tuple(self._base_rpy)

    def GetBaseRollPitchYawRate(self):
        return np.asarray(self._base_rpy_rate)

    def GetFootPositionsInBaseFrame(self):
        return np.array(self._foot_positions)

    def GetHipPositionsInBaseFrame(self):
        return self._constants.DEFAULT_HIP_POSITIONS

    def GetMotorPositionGains(self):
        return self._motor.MOTOR_POSITION_GAINS

    def GetMotorVelocityGains(self):
        return self._motor.MOTOR_VELOCITY_GAINS

    def GetMotorAngles(self):
        return np.array(self._motor_angles)

    def ComputeMotorAnglesFromFootLocalPosition(self, leg_id, foot_local_position):
        """kinematics.py:98-133 -- returns (joint position indices, motor angles)."""
        idx = list(range(3 * leg_id, 3 * leg_id + 3))
        seed = self.joint_angles(self._constants.INIT_MOTOR_ANGLES)[idx]
        q = self._chains[leg_id].ik(foot_local_position, seed)
        angles = (q - self._offset[idx]) * self._direction[idx]
        return idx, angles.tolist()

    def MapContactForceToJointTorques(self, leg_id, contact_force):
        """kinematics.py:40-53 -- torque_j = (f . J[:, j]) * MOTOR_DIRECTION[j]."""
        q = self.joint_angles(self._motor_angles)[3 * leg_id:3 * leg_id + 3]
        _, jac = self._chains[leg_id].fk(q, with_jacobian=True)
        all_torques = np.matmul(np.asarray(contact_force, dtype=np.float64), jac)
        return {3 * leg_id + j: all_torques[j] * self._direction[3 * leg_id + j] for j in range(3)}


def hybrid_motor_torque(action60, q, qd):
    """HYBRID branch of RobotMotorModel.convert_to_torque (robot_gym/model/robots/simple_motor.py:128-139)."""
    a = np.asarray(action60, dtype=np.float64)
    kp, kd = a[1::5], a[3::5]
    return -1 * (kp * (np.asarray(q) - a[0::5])) - kd * (np.asarray(qd) - a[2::5]) + a[4::5]
