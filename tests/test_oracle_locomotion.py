"""Pins for the restated gait generator / estimator / swing controller / kinematics / glue
(oracle/locomotion.py, oracle/kinematics.py): known answers the survey lists and
the reference's own in-tree numbers (constants, hybrid motor model outputs)."""
import math
import os

import numpy as np
import pytest

from oracle import kinematics, locomotion
from oracle.locomotion import EARLY_CONTACT, LOSE_CONTACT, STANCE, SWING
from robot_gym.model.robots.descriptions import GHOST, K3LSO
from robot_gym.util import synthetic


def _gait(robot, ctrl=None):
    ctrl = ctrl or GHOST.GetCtrlConstants()
    return locomotion.OpenloopGaitGenerator(robot, ctrl.STANCE_DURATION_SECONDS, ctrl.DUTY_FACTOR,
                                            ctrl.INIT_PHASE_FULL_CYCLE, ctrl.INIT_LEG_STATE)


def test_gait_initial_state_and_period():
    robot = kinematics.OracleRobot(GHOST)
    robot.set_state(foot_contacts=(0, 1, 1, 0))
    g = _gait(robot)
    assert g.desired_leg_state == [SWING, STANCE, STANCE, SWING]        # reset state: ghost/ctrl_constants.py:32-37
    g.update(0.0)
    # legs 0/3 start at phase 0.9 of their SWING-first cycle (swing ratio 0.4): 0.9 >= 0.4, so they are
    # 5/6 through the following STANCE and lift off 0.05 s later; legs 1/2 start their stance.
    period = 0.3 / 0.6
    assert math.fmod(0.0 + 0.9 * period, period) / period == pytest.approx(0.9)
    assert g.desired_leg_state == [STANCE, STANCE, STANCE, STANCE]
    np.testing.assert_allclose(g.normalized_phase, [(0.9 - 0.4) / 0.6, 0.0, 0.0, (0.9 - 0.4) / 0.6], atol=1e-12)
    g.update(0.051)
    assert g.desired_leg_state == [SWING, STANCE, STANCE, SWING]
    # cycle period 0.5 s: states repeat
    g.update(0.123)
    a = (list(g.desired_leg_state), np.array(g.normalized_phase))
    g.update(0.123 + 4 * period)
    assert list(g.desired_leg_state) == a[0]
    np.testing.assert_allclose(g.normalized_phase, a[1], atol=1e-12)


def test_gait_contact_overrides_respect_threshold():
    robot = kinematics.OracleRobot(GHOST)
    ctrl = GHOST.GetCtrlConstants()
    g = _gait(robot)
    # find a time where leg 1 (initially STANCE) is in stance with normalized phase > 0.1
    robot.set_state(foot_contacts=(1, 0, 1, 1))
    g.update(0.1)
    assert g.desired_leg_state[1] == STANCE and g.normalized_phase[1] > 0.1
    assert g.leg_state[1] == LOSE_CONTACT
    robot.set_state(foot_contacts=(1, 0, 1, 1))
    g.update(0.01)                                   # phase 0.01/0.3 < 0.1: no contact detection yet
    assert g.normalized_phase[1] < 0.1 and g.leg_state[1] == STANCE
    # a swing leg that touches down late in its swing becomes EARLY_CONTACT
    for t in np.arange(0.0, 0.5, 0.001):
        robot.set_state(foot_contacts=(1, 1, 1, 1))
        g.update(float(t))
        for leg in range(4):
            if g.desired_leg_state[leg] == SWING and g.normalized_phase[leg] >= 0.1:
                assert g.leg_state[leg] == EARLY_CONTACT
            if g.desired_leg_state[leg] == SWING and g.normalized_phase[leg] < 0.1:
                assert g.leg_state[leg] == SWING


def test_generator_gait_matches_oracle_planned_contacts():
    """util/synthetic.desired_stance (host logic used to build bench inputs) == oracle gait."""
    robot = kinematics.OracleRobot(GHOST)
    g = _gait(robot)
    ctrl = GHOST.GetCtrlConstants()
    ts = np.arange(0, 600) * 0.001
    planned = synthetic.desired_stance(ts, ctrl.STANCE_DURATION_SECONDS, ctrl.DUTY_FACTOR,
                                       ctrl.INIT_PHASE_FULL_CYCLE, ctrl.INIT_LEG_STATE)
    for k, t in enumerate(ts):
        g.update(float(t))
        assert [s == STANCE for s in g.desired_leg_state] == list(planned[k])


def test_moving_window_filter_divides_by_window_even_when_not_full():
    f = locomotion.MovingWindowFilter(4)
    assert f.calculate_average(4.0) == 1.0           # (4)/4, not 4/1
    assert f.calculate_average(4.0) == 2.0
    f.calculate_average(4.0); f.calculate_average(4.0)
    assert f.calculate_average(8.0) == 5.0           # oldest dropped: (4+4+4+8)/4
    # Neumaier compensation keeps tiny values next to huge ones
    g = locomotion.MovingWindowFilter(3)
    g.calculate_average(1e16); g.calculate_average(1.0)
    assert g.calculate_average(-1e16) == pytest.approx(1.0 / 3)


def test_estimator_rotates_into_body_frame():
    robot = kinematics.OracleRobot(GHOST)
    yaw = 0.5 * math.pi
    robot.set_state(base_velocity=(1.0, 0.0, 0.0), base_orientation=(0, 0, math.sin(yaw / 2), math.cos(yaw / 2)))
    est = locomotion.COMVelocityEstimator(robot, window_size=1)
    est.update(0)
    np.testing.assert_allclose(est.com_velocity_body_frame, [0.0, -1.0, 0.0], atol=1e-12)


def test_swing_trajectory_endpoints_and_apex():
    start, end = (0.1, -0.1, -0.40), (0.2, -0.12, -0.41)
    p0 = locomotion._gen_swing_foot_trajectory(0.0, start, end)
    p1 = locomotion._gen_swing_foot_trajectory(1.0, start, end)
    np.testing.assert_allclose(p0, start, atol=1e-12)
    np.testing.assert_allclose(p1, end, atol=1e-12)
    # warped phase 0.5 is reached at sin(pi p) = 0.625; the parabola peaks there at max(z)+0.1
    p_mid = math.asin(0.625) / math.pi
    apex = locomotion._gen_swing_foot_trajectory(p_mid, start, end)
    assert apex[2] == pytest.approx(max(start[2], end[2]) + 0.1)


@pytest.mark.parametrize("desc", [GHOST, K3LSO])
def test_fk_nominal_and_ik_round_trip(desc):
    robot = kinematics.OracleRobot(desc)
    q0 = np.asarray(desc.GetConstants().INIT_MOTOR_ANGLES, dtype=np.float64)
    feet = robot.fk_all(q0)
    if desc is GHOST:
        np.testing.assert_allclose(feet, synthetic.GHOST_NOMINAL_FEET, atol=6e-5)
    assert np.ptp(feet[:, 2]) < 2e-3                   # mirrored init angles: feet at (almost) equal height
    rng = np.random.default_rng(0)
    for _ in range(10):
        q = q0 + rng.uniform(-0.3, 0.3, 12)
        target = robot.fk_all(q)
        for leg in range(4):
            idx, ang = robot.ComputeMotorAnglesFromFootLocalPosition(leg, target[leg])
            assert idx == [3 * leg, 3 * leg + 1, 3 * leg + 2]
            np.testing.assert_allclose(ang, q[idx], atol=1e-9)
            back = robot.fk_all(np.where(np.isin(np.arange(12), idx), np.resize(ang, 12)[np.arange(12) % 3], q))


def test_jacobian_matches_finite_differences_and_torque_map():
    robot = kinematics.OracleRobot(GHOST)
    chain = robot._chains[1]
    q = np.array([0.1, 0.7, -1.2])
    _, jac = chain.fk(q, with_jacobian=True)
    num = np.zeros((3, 3))
    for j in range(3):
        dq = np.zeros(3); dq[j] = 1e-6
        num[:, j] = (chain.fk(q + dq) - chain.fk(q - dq)) / 2e-6
    np.testing.assert_allclose(jac, num, atol=1e-8)
    robot.set_state(motor_angles=np.tile(q, 4))
    tau = robot.MapContactForceToJointTorques(1, [1.0, 2.0, -30.0])
    np.testing.assert_allclose([tau[3], tau[4], tau[5]], np.array([1.0, 2.0, -30.0]) @ jac, atol=1e-12)


def test_hybrid_motor_model_matches_reference_outputs(reference_constants):
    kat = reference_constants["hybrid_motor_kat"]      # produced by the reference's own RobotMotorModel
    for cmd, q, qd, tau in zip(kat["commands"], kat["q"], kat["qd"], kat["torque"]):
        np.testing.assert_allclose(kinematics.hybrid_motor_torque(cmd, q, qd), tau, rtol=0, atol=1e-12)


def test_locomotion_glue_action_layout_and_first_update_alias():
    ctrl = GHOST.GetCtrlConstants()
    robot = kinematics.OracleRobot(GHOST)
    clock = {"t": 0.0}
    ctl = locomotion.build_mpc_controller(robot, lambda: clock["t"], ctrl)
    ctl.reset()
    locomotion.update_controller_params(ctl, ctrl, (0.2, 0.1))
    assert ctl.swing_leg_controller.desired_speed == [0.2 + ctrl.VX_OFFSET, 0.0 + ctrl.VY_OFFSET, 0.0]
    assert ctl.stance_leg_controller.desired_twisting_speed == 0.1 + ctrl.WZ_OFFSET
    latch0 = np.array(ctl.swing_leg_controller._phase_switch_foot_local_position)
    robot.set_state(foot_positions=robot.GetFootPositionsInBaseFrame() + 0.01, foot_contacts=(0, 1, 1, 0))
    clock["t"] = 0.06                                  # legs 0/3 lifted off at 0.05 s
    ctl.update()
    assert ctl.gait_generator.desired_leg_state == [SWING, STANCE, STANCE, SWING]
    # first update after reset: last_leg_state aliases the gait's list -> no latch although legs 0/3 swing
    np.testing.assert_array_equal(ctl.swing_leg_controller._phase_switch_foot_local_position, latch0)
    action = ctl.get_action()
    assert action.shape == (60,) and action.dtype == np.float32
    a = action.reshape(12, 5)
    swing_motors = [0, 1, 2, 9, 10, 11]
    np.testing.assert_array_equal(a[swing_motors, 1], 220.0)             # kp
    np.testing.assert_array_equal(a[swing_motors, 3], [1, 2, 2, 1, 2, 2])  # kd
    np.testing.assert_array_equal(a[swing_motors, 4], 0.0)
    stance_motors = [3, 4, 5, 6, 7, 8]
    np.testing.assert_array_equal(a[stance_motors, :4], 0.0)
    assert np.abs(a[stance_motors, 4]).max() > 0.1


def test_control_step_oracle_matches_frozen_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "control_step_oracle_golden.npz"))
    ctrl = GHOST.GetCtrlConstants()
    n_steps = 6
    seq = synthetic.make_state_sequence(int(g["n_env"]), int(g["n_steps"]), GHOST)[:n_steps]
    e = 2
    robot = kinematics.OracleRobot(GHOST)
    clock = {"t": 0.0}
    def load(k):
        s = seq[k]
        robot.set_state(base_velocity=s.base_velocity_world[e].astype(np.float64),
                        base_orientation=s.base_orientation_xyzw[e].astype(np.float64),
                        base_rpy=s.base_rpy[e].astype(np.float64), base_rpy_rate=s.base_rpy_rate[e].astype(np.float64),
                        foot_positions=s.foot_positions_base[e].astype(np.float64), foot_contacts=s.foot_contacts[e],
                        motor_angles=s.motor_angles[e].astype(np.float64))
        clock["t"] = float(s.time_since_reset[e])
    load(0)
    ctl = locomotion.build_mpc_controller(robot, lambda: clock["t"], ctrl)
    ctl.reset()
    for k in range(n_steps):
        load(k)
        s = seq[k]
        for leg_ctl in (ctl.swing_leg_controller, ctl.stance_leg_controller):
            leg_ctl.desired_speed = [float(s.command[e, 0]), float(s.command[e, 1]), 0.0]
            leg_ctl.desired_twisting_speed = float(s.command[e, 2])
        ctl.update()
        action = ctl.get_action()
        np.testing.assert_array_equal(ctl.gait_generator.desired_leg_state, g["desired_leg_state"][k, e])
        np.testing.assert_array_equal(ctl.gait_generator.leg_state, g["leg_state"][k, e])
        np.testing.assert_array_equal(np.asarray(ctl.gait_generator.normalized_phase).view(np.int64),
                                      g["normalized_phase"][k, e].view(np.int64))     # bit-exact
        np.testing.assert_allclose(action, g["actions"][k, e], rtol=1e-6, atol=1e-6)


def test_quaternion_conversions_round_trip_and_gimbal_branches():
    """State-provider conversions (robot.py:79-86,185-203): euler -> quaternion -> euler is the identity
    away from gimbal lock, the Euler angles reproduce the rotation matrix, and R(q)^T w is the inverse
    rotation of w."""
    rng = np.random.default_rng(11)
    rpy = np.column_stack([rng.uniform(-3.1, 3.1, 200), rng.uniform(-1.5, 1.5, 200), rng.uniform(-3.1, 3.1, 200)])
    quat = synthetic.euler_to_quat_xyzw(rpy)
    for e, q in zip(rpy, quat):
        back = kinematics.quat_to_rpy(q)
        assert np.abs(back - e).max() < 1e-9
        r, p, y = back
        rz = np.array([[math.cos(y), -math.sin(y), 0], [math.sin(y), math.cos(y), 0], [0, 0, 1]])
        ry = np.array([[math.cos(p), 0, math.sin(p)], [0, 1, 0], [-math.sin(p), 0, math.cos(p)]])
        rx = np.array([[1, 0, 0], [0, math.cos(r), -math.sin(r)], [0, math.sin(r), math.cos(r)]])
        assert np.abs(rz @ ry @ rx - kinematics.quat_to_matrix(q)).max() < 1e-12
        w = rng.uniform(-2, 2, 3)
        local = kinematics.angular_velocity_to_local_frame(w, q)
        assert np.abs(kinematics.quat_to_matrix(q) @ local - w).max() < 1e-12
    # gimbal lock: pitch = +-90 deg keeps roll = 0 and folds everything into yaw
    for sign in (1.0, -1.0):
        q = synthetic.euler_to_quat_xyzw(np.array([[0.0, sign * math.pi / 2, 0.7]]))[0]
        r, p, y = kinematics.quat_to_rpy(q)
        assert r == 0.0 and abs(p - sign * math.pi / 2) < 1e-12 and abs(y - 0.7) < 1e-6
    assert np.all(kinematics.quat_to_rpy([0, 0, 0, 1]) == 0.0)


def test_oracle_robot_sim_state_injection_matches_direct_state():
    robot = kinematics.OracleRobot(GHOST)
    motor = np.asarray(GHOST.GetConstants().INIT_MOTOR_ANGLES, dtype=np.float64) + 0.05
    joints = robot.joint_angles(motor)
    q = synthetic.euler_to_quat_xyzw(np.array([[0.1, -0.2, 0.3]]))[0]
    robot.set_sim_state(q, (0.3, 0.0, -0.1), (0.2, -0.4, 0.5), joints, (1, 0, 0, 1))
    assert np.abs(np.array(robot.GetBaseRollPitchYaw()) - [0.1, -0.2, 0.3]).max() < 1e-12
    assert np.abs(robot.GetMotorAngles() - motor).max() < 1e-12
    assert np.abs(robot.GetFootPositionsInBaseFrame() - robot.fk_all(motor)).max() == 0.0
    assert np.abs(kinematics.quat_to_matrix(q) @ robot.GetBaseRollPitchYawRate() - [0.2, -0.4, 0.5]).max() < 1e-12
    assert robot.GetFootContacts() == [True, False, False, True]
