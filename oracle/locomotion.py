"""Oracle: per-env restatement of motion_imitation's ``mpc_controller`` python stack.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED (see oracle/__init__.py): the classes below restate, from the published
google-research/motion_imitation sources, the third-party objects robot-gym wires together in
robot_gym/controllers/mpc/mpc_controller.py:28-66 and drives at :102-109:

    OpenloopGaitGenerator      (call site mpc_controller.py:30-35)
    COMVelocityEstimator       (call site :36)           + MovingWindowFilter
    RaibertSwingLegController  (call site :38-45)
    TorqueStanceLegController  (call site :47-56)
    LocomotionController       (call site :58-65, update/get_action at :104-105)

They run ONE env in plain CPython float64, in the reference's operation order, against a robot
object that implements the callback surface of robot_gym/model/robots/robot.py (getter names
unchanged).  ``oracle.kinematics.OracleRobot`` is such an object built from the URDF leg chains.
"""
from __future__ import annotations

import collections
import copy
import math

import numpy as np

from oracle import convex_mpc

SWING, STANCE, EARLY_CONTACT, LOSE_CONTACT = 0, 1, 2, 3   # gait_generator.LegState

_NOMINAL_CONTACT_DETECTION_PHASE = 0.1


class OpenloopGaitGenerator:
    """mpc_controller/openloop_gait_generator.py."""

    def __init__(self, robot, stance_duration, duty_factor, initial_leg_phase, initial_leg_state,
                 contact_detection_phase_threshold=_NOMINAL_CONTACT_DETECTION_PHASE):
        self._robot = robot
        self._stance_duration = list(stance_duration)
        self._duty_factor = list(duty_factor)
        self._swing_duration = np.array(stance_duration) / np.array(duty_factor) - np.array(stance_duration)
        if len(initial_leg_phase) != robot.num_legs:
            raise ValueError("The number of leg phases should be the same as number of legs.")
        self._initial_leg_phase = list(initial_leg_phase)
        if len(initial_leg_state) != robot.num_legs:
            raise ValueError("The number of leg states should be the same of number of legs.")
        self._initial_leg_state = [int(s) for s in initial_leg_state]
        self._next_leg_state = []
        self._initial_state_ratio_in_cycle = []
        for state, duty in zip(self._initial_leg_state, self._duty_factor):
            if state == SWING:
                self._initial_state_ratio_in_cycle.append(1 - duty)
                self._next_leg_state.append(STANCE)
            else:
                self._initial_state_ratio_in_cycle.append(duty)
                self._next_leg_state.append(SWING)
        self._contact_detection_phase_threshold = contact_detection_phase_threshold
        self._normalized_phase = None
        self._leg_state = None
        self._desired_leg_state = None
        self.reset(0)

    def reset(self, current_time):
        self._normalized_phase = np.zeros(self._robot.num_legs)
        self._leg_state = list(self._initial_leg_state)
        self._desired_leg_state = list(self._initial_leg_state)

    @property
    def desired_leg_state(self):
        return self._desired_leg_state

    @property
    def leg_state(self):
        return self._leg_state

    @property
    def swing_duration(self):
        return self._swing_duration

    @property
    def stance_duration(self):
        return self._stance_duration

    @property
    def normalized_phase(self):
        return self._normalized_phase

    def update(self, current_time):
        contact_state = self._robot.GetFootContacts()
        for leg_id in range(self._robot.num_legs):
            full_cycle_period = self._stance_duration[leg_id] / self._duty_factor[leg_id]
            augmented_time = current_time + self._initial_leg_phase[leg_id] * full_cycle_period
            phase_in_full_cycle = math.fmod(augmented_time, full_cycle_period) / full_cycle_period
            ratio = self._initial_state_ratio_in_cycle[leg_id]
            if phase_in_full_cycle < ratio:
                self._desired_leg_state[leg_id] = self._initial_leg_state[leg_id]
                self._normalized_phase[leg_id] = phase_in_full_cycle / ratio
            else:
                self._desired_leg_state[leg_id] = self._next_leg_state[leg_id]
                self._normalized_phase[leg_id] = (phase_in_full_cycle - ratio) / (1 - ratio)
            self._leg_state[leg_id] = self._desired_leg_state[leg_id]
            if self._normalized_phase[leg_id] < self._contact_detection_phase_threshold:
                continue
            if self._leg_state[leg_id] == SWING and contact_state[leg_id]:
                self._leg_state[leg_id] = EARLY_CONTACT
            if self._leg_state[leg_id] == STANCE and not contact_state[leg_id]:
                self._leg_state[leg_id] = LOSE_CONTACT


class MovingWindowFilter:
    """mpc_controller/com_velocity_estimator.py helper: moving mean with Neumaier summation."""

    def __init__(self, window_size):
        assert window_size > 0
        self._window_size = window_size
        self._value_deque = collections.deque(maxlen=window_size)
        self._sum = 0
        self._correction = 0

    def _neumaier_sum(self, value):
        new_sum = self._sum + value
        if abs(self._sum) >= abs(value):
            self._correction += (self._sum - new_sum) + value
        else:
            self._correction += (value - new_sum) + self._sum
        self._sum = new_sum

    def calculate_average(self, new_value):
        deque_len = len(self._value_deque)
        if deque_len < self._value_deque.maxlen:
            pass
        else:
            self._neumaier_sum(-self._value_deque[0])
        self._neumaier_sum(new_value)
        self._value_deque.append(new_value)
        return (self._sum + self._correction) / self._window_size


def quat_rotate_inverse(q_xyzw, v):
    """pybullet invertTransform((0,0,0), q) then multiplyTransforms((0,0,0), q_inv, v, identity):
    R(q)^T v with Bullet's quaternion->matrix formula (btMatrix3x3::setRotation)."""
    x, y, z, w = (float(c) for c in q_xyzw)
    d = x * x + y * y + z * z + w * w
    s = 2.0 / d
    xs, ys, zs = x * s, y * s, z * s
    wx, wy, wz = w * xs, w * ys, w * zs
    xx, xy, xz = x * xs, x * ys, x * zs
    yy, yz, zz = y * ys, y * zs, z * zs
    m = np.array([[1.0 - (yy + zz), xy - wz, xz + wy],
                  [xy + wz, 1.0 - (xx + zz), yz - wx],
                  [xz - wy, yz + wx, 1.0 - (xx + yy)]])
    return m.T @ np.asarray(v, dtype=np.float64)


class COMVelocityEstimator:
    """mpc_controller/com_velocity_estimator.py."""

    def __init__(self, robot, window_size=20):
        self._robot = robot
        self._window_size = window_size
        self.reset(0)

    @property
    def com_velocity_body_frame(self):
        return self._com_velocity_body_frame

    @property
    def com_velocity_world_frame(self):
        return self._com_velocity_world_frame

    def reset(self, current_time):
        del current_time
        self._velocity_filter_x = MovingWindowFilter(self._window_size)
        self._velocity_filter_y = MovingWindowFilter(self._window_size)
        self._velocity_filter_z = MovingWindowFilter(self._window_size)
        self._com_velocity_world_frame = np.array((0, 0, 0))
        self._com_velocity_body_frame = np.array((0, 0, 0))

    def update(self, current_time):
        del current_time
        velocity = self._robot.GetBaseVelocity()
        vx = self._velocity_filter_x.calculate_average(velocity[0])
        vy = self._velocity_filter_y.calculate_average(velocity[1])
        vz = self._velocity_filter_z.calculate_average(velocity[2])
        self._com_velocity_world_frame = np.array((vx, vy, vz))
        base_orientation = self._robot.GetTrueBaseOrientation()
        self._com_velocity_body_frame = quat_rotate_inverse(base_orientation, self._com_velocity_world_frame)


_KP = np.array([0.01, 0.01, 0.01]) * 3.


def _gen_parabola(phase, start, mid, end):
    mid_phase = 0.5
    delta_1 = mid - start
    delta_2 = end - start
    delta_3 = mid_phase ** 2 - mid_phase
    coef_a = (delta_1 - delta_2 * mid_phase) / delta_3
    coef_b = (delta_2 * mid_phase ** 2 - delta_1) / delta_3
    coef_c = start
    return coef_a * phase ** 2 + coef_b * phase + coef_c


def _gen_swing_foot_trajectory(input_phase, start_pos, end_pos):
    phase = input_phase
    if input_phase <= 0.5:
        phase = 0.8 * math.sin(input_phase * math.pi)
    else:
        phase = 0.8 + (input_phase - 0.5) * 0.4
    x = (1 - phase) * start_pos[0] + phase * end_pos[0]
    y = (1 - phase) * start_pos[1] + phase * end_pos[1]
    max_clearance = 0.1
    mid = max(end_pos[2], start_pos[2]) + max_clearance
    z = _gen_parabola(phase, start_pos[2], mid, end_pos[2])
    return (x, y, z)


class RaibertSwingLegController:
    """mpc_controller/raibert_swing_leg_controller.py."""

    def __init__(self, robot, gait_generator, state_estimator, desired_speed, desired_twisting_speed,
                 desired_height, foot_clearance):
        self._robot = robot
        self._state_estimator = state_estimator
        self._gait_generator = gait_generator
        self._last_leg_state = gait_generator.desired_leg_state
        self.desired_speed = np.array((desired_speed[0], desired_speed[1], 0))
        self.desired_twisting_speed = desired_twisting_speed
        self._desired_height = np.array((0, 0, desired_height - foot_clearance))
        self._joint_angles = None
        self._phase_switch_foot_local_position = None
        self.foot_targets = {}            # oracle-only: last swing trajectory point per leg (for tests)
        self.reset(0)

    def reset(self, current_time):
        del current_time
        # NOTE: this aliases the gait generator's list (no copy), so the first update() after a
        # reset sees new_state == last_state for every leg and never latches.  Kept on purpose.
        self._last_leg_state = self._gait_generator.desired_leg_state
        self._phase_switch_foot_local_position = np.array(self._robot.GetFootPositionsInBaseFrame(), dtype=np.float64)
        self._joint_angles = {}
        self.foot_targets = {}

    def update(self, current_time):
        del current_time
        new_leg_state = self._gait_generator.desired_leg_state
        for leg_id, state in enumerate(new_leg_state):
            if state == SWING and state != self._last_leg_state[leg_id]:
                self._phase_switch_foot_local_position[leg_id] = self._robot.GetFootPositionsInBaseFrame()[leg_id]
        self._last_leg_state = copy.deepcopy(new_leg_state)

    def get_action(self):
        com_velocity = self._state_estimator.com_velocity_body_frame
        com_velocity = np.array((com_velocity[0], com_velocity[1], 0))
        _, _, yaw_dot = self._robot.GetBaseRollPitchYawRate()
        hip_positions = self._robot.GetHipPositionsInBaseFrame()
        for leg_id, leg_state in enumerate(self._gait_generator.leg_state):
            if leg_state in (STANCE, EARLY_CONTACT):
                continue
            hip_offset = hip_positions[leg_id]
            twisting_vector = np.array((-hip_offset[1], hip_offset[0], 0))
            hip_horizontal_velocity = com_velocity + yaw_dot * twisting_vector
            target_hip_horizontal_velocity = self.desired_speed + self.desired_twisting_speed * twisting_vector
            foot_target_position = (
                hip_horizontal_velocity * self._gait_generator.stance_duration[leg_id] / 2 -
                _KP * (target_hip_horizontal_velocity - hip_horizontal_velocity)
            ) - self._desired_height + np.array((hip_offset[0], hip_offset[1], 0))
            foot_position = _gen_swing_foot_trajectory(
                self._gait_generator.normalized_phase[leg_id],
                self._phase_switch_foot_local_position[leg_id], foot_target_position)
            self.foot_targets[leg_id] = np.array(foot_position)
            joint_ids, joint_angles = self._robot.ComputeMotorAnglesFromFootLocalPosition(leg_id, foot_position)
            for joint_id, joint_angle in zip(joint_ids, joint_angles):
                self._joint_angles[joint_id] = (joint_angle, leg_id)
        action = {}
        kps = self._robot.GetMotorPositionGains()
        kds = self._robot.GetMotorVelocityGains()
        for joint_id, joint_angle_leg_id in self._joint_angles.items():
            leg_id = joint_angle_leg_id[1]
            if self._gait_generator.desired_leg_state[leg_id] == SWING:
                action[joint_id] = (joint_angle_leg_id[0], kps[joint_id], 0, kds[joint_id], 0)
        return action


_FORCE_DIMENSION = 3


class TorqueStanceLegController:
    """mpc_controller/torque_stance_leg_controller.py (the mpc_osqp call is oracle.convex_mpc)."""

    def __init__(self, robot, gait_generator, state_estimator, desired_speed=(0, 0), desired_twisting_speed=0,
                 desired_body_height=0.45, body_mass=220 / 9.8,
                 body_inertia=(0.07335, 0, 0, 0, 0.25068, 0, 0, 0, 0.25447), num_legs=4,
                 friction_coeffs=(0.45, 0.45, 0.45, 0.45), mpc_params=None, mpc_solver=None):
        # mpc_solver: callable with the signature of convex_mpc.compute_contact_forces (default) -- the tests that
        # step hundreds of envs pass oracle.c_oracle.compute_contact_forces, the same QP solved by the C port
        self._mpc_solver = mpc_solver or convex_mpc.compute_contact_forces
        self._robot = robot
        self._gait_generator = gait_generator
        self._state_estimator = state_estimator
        self.desired_speed = desired_speed
        self.desired_twisting_speed = desired_twisting_speed
        self._desired_body_height = desired_body_height
        self._num_legs = num_legs
        self._friction_coeffs = np.array(friction_coeffs)
        self._params = mpc_params or convex_mpc.MpcParams(mass=body_mass, inertia=tuple(body_inertia),
                                                           num_legs=num_legs, friction_coeffs=tuple(friction_coeffs))
        self.last_contact_forces = None
        self.last_foot_contact_state = None

    def reset(self, current_time):
        del current_time

    def update(self, current_time):
        del current_time

    def get_action(self):
        desired_com_position = np.array((0., 0., self._desired_body_height), dtype=np.float64)
        desired_com_velocity = np.array((self.desired_speed[0], self.desired_speed[1], 0.), dtype=np.float64)
        desired_com_roll_pitch_yaw = np.array((0., 0., 0.), dtype=np.float64)
        desired_com_angular_velocity = np.array((0., 0., self.desired_twisting_speed), dtype=np.float64)
        foot_contact_state = np.array(
            [(leg_state in (STANCE, EARLY_CONTACT)) for leg_state in self._gait_generator.desired_leg_state],
            dtype=np.int32)
        com_roll_pitch_yaw = np.array(self._robot.GetBaseRollPitchYaw(), dtype=np.float64)
        com_roll_pitch_yaw[2] = 0
        predicted_contact_forces = self._mpc_solver(
            self._params,
            np.asarray(self._state_estimator.com_velocity_body_frame, dtype=np.float64),
            com_roll_pitch_yaw,
            np.asarray(self._robot.GetBaseRollPitchYawRate(), dtype=np.float64),
            foot_contact_state,
            np.array(self._robot.GetFootPositionsInBaseFrame().flatten(), dtype=np.float64),
            desired_com_position, desired_com_velocity, desired_com_roll_pitch_yaw,
            desired_com_angular_velocity, com_position=[0])
        contact_forces = {}
        for i in range(self._num_legs):
            contact_forces[i] = np.array(predicted_contact_forces[i * _FORCE_DIMENSION:(i + 1) * _FORCE_DIMENSION])
        self.last_contact_forces = np.array(predicted_contact_forces[:3 * self._num_legs])
        self.last_foot_contact_state = foot_contact_state
        action = {}
        for leg_id, force in contact_forces.items():
            motor_torques = self._robot.MapContactForceToJointTorques(leg_id, force)
            for joint_id, torque in motor_torques.items():
                action[joint_id] = (0, 0, 0, 0, torque)
        return action


class LocomotionController:
    """mpc_controller/locomotion_controller.py."""

    def __init__(self, robot, gait_generator, state_estimator, swing_leg_controller, stance_leg_controller, clock):
        self._robot = robot
        self._clock = clock
        self._reset_time = self._clock()
        self._time_since_reset = 0
        self._gait_generator = gait_generator
        self._state_estimator = state_estimator
        self._swing_leg_controller = swing_leg_controller
        self._stance_leg_controller = stance_leg_controller

    @property
    def swing_leg_controller(self):
        return self._swing_leg_controller

    @property
    def stance_leg_controller(self):
        return self._stance_leg_controller

    @property
    def gait_generator(self):
        return self._gait_generator

    @property
    def state_estimator(self):
        return self._state_estimator

    def reset(self):
        self._reset_time = self._clock()
        self._time_since_reset = 0
        self._gait_generator.reset(self._time_since_reset)
        self._state_estimator.reset(self._time_since_reset)
        self._swing_leg_controller.reset(self._time_since_reset)
        self._stance_leg_controller.reset(self._time_since_reset)

    def update(self):
        self._time_since_reset = self._clock() - self._reset_time
        self._gait_generator.update(self._time_since_reset)
        self._state_estimator.update(self._time_since_reset)
        self._swing_leg_controller.update(self._time_since_reset)
        self._stance_leg_controller.update(self._time_since_reset)

    def get_action(self):
        swing_action = self._swing_leg_controller.get_action()
        stance_action = self._stance_leg_controller.get_action()
        action = []
        for joint_id in range(self._robot.num_motors):
            if joint_id in swing_action:
                action.extend(swing_action[joint_id])
            else:
                assert joint_id in stance_action
                action.extend(stance_action[joint_id])
        return np.array(action, dtype=np.float32)


def build_mpc_controller(robot, clock, constants, mpc_params=None, mpc_solver=None):
    """``MPCController._setup_controller`` (robot_gym/controllers/mpc/mpc_controller.py:28-66)."""
    gait = OpenloopGaitGenerator(robot, stance_duration=constants.STANCE_DURATION_SECONDS,
                                 duty_factor=constants.DUTY_FACTOR,
                                 initial_leg_phase=constants.INIT_PHASE_FULL_CYCLE,
                                 initial_leg_state=constants.INIT_LEG_STATE)
    est = COMVelocityEstimator(robot, window_size=20)
    sw = RaibertSwingLegController(robot, gait, est, desired_speed=(0.0, 0.0), desired_twisting_speed=0.0,
                                   desired_height=constants.MPC_BODY_HEIGHT, foot_clearance=0.01)
    stc = TorqueStanceLegController(robot, gait, est, desired_speed=(0.0, 0.0), desired_twisting_speed=0.0,
                                    desired_body_height=constants.MPC_BODY_HEIGHT, body_mass=constants.MPC_BODY_MASS,
                                    body_inertia=constants.MPC_BODY_INERTIA, mpc_params=mpc_params, mpc_solver=mpc_solver)
    return LocomotionController(robot, gait, est, sw, stc, clock)


def update_controller_params(controller, constants, params):
    """``MPCController.update_controller_params`` (mpc_controller.py:83-100)."""
    if len(params) == 2:
        vx, wz = params
        vy = 0.
    else:
        vx, vy, wz = params
    lin_speed = [vx + constants.VX_OFFSET, vy + constants.VY_OFFSET, 0.]
    ang_speed = wz + constants.WZ_OFFSET
    controller.swing_leg_controller.desired_speed = lin_speed
    controller.swing_leg_controller.desired_twisting_speed = ang_speed
    controller.stance_leg_controller.desired_speed = lin_speed
    controller.stance_leg_controller.desired_twisting_speed = ang_speed
