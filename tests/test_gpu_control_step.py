"""GPU parity of the full control step and of its parts (gait / estimator / swing / IK / FK /
Jacobian^T / pack / hybrid motor) against the oracle, through the C ABI and through the
BatchedMPCController drop-in.  Gait phase and leg-state indices are compared BIT-EXACTLY."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import kinematics, locomotion
from robot_gym import cuda as rg
from robot_gym.controllers.mpc.batched_kinematics import BatchedKinematics, robot_params_from_description
from robot_gym.controllers.mpc.batched_mpc_controller import BatchedMPCController
from robot_gym.model.robots.descriptions import GHOST, K3LSO, with_gait
from robot_gym.model.robots.sim_state_robot import SimStateRobotBatch
from robot_gym.model.robots.synthetic_robot import SyntheticRobotBatch
from robot_gym.util import synthetic

pytestmark = pytest.mark.gpu


def _dev(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


@pytest.mark.parametrize("schedule", ["trot", "walk", "bound"])
def test_gait_step_is_bit_exact(rg_lib, cuda_device, schedule):
    desc = with_gait(GHOST, schedule)
    ctrl = desc.GetCtrlConstants()
    ws = rg.RobotWorkspace(robot_params_from_description(desc), device=cuda_device)
    n = 6000
    rng = np.random.default_rng(1)
    t = np.concatenate([np.arange(3000) * 0.001, rng.uniform(0, 30.0, n - 3000)])      # grid times + arbitrary doubles
    contacts = (rng.uniform(0, 1, (n, 4)) < 0.5).astype(np.uint8)
    desired = torch.empty((n, 4), dtype=torch.int32, device=cuda_device)
    state = torch.empty_like(desired)
    phase = torch.empty((n, 4), dtype=torch.float64, device=cuda_device)
    p = ctypes.c_void_p
    t_dev, c_dev = _dev(t, cuda_device), _dev(contacts, cuda_device)     # keep alive across the launch
    rg.check(rg_lib.rg_gait_step(ws.ptr, n, p(t_dev.data_ptr()), p(c_dev.data_ptr()),
                                 p(desired.data_ptr()), p(state.data_ptr()), p(phase.data_ptr()), None))
    torch.cuda.synchronize()
    robot = kinematics.OracleRobot(desc)
    gait = locomotion.OpenloopGaitGenerator(robot, ctrl.STANCE_DURATION_SECONDS, ctrl.DUTY_FACTOR,
                                            ctrl.INIT_PHASE_FULL_CYCLE, ctrl.INIT_LEG_STATE)
    d, s, ph = desired.cpu().numpy(), state.cpu().numpy(), phase.cpu().numpy()
    for i in range(n):
        robot.set_state(foot_contacts=contacts[i])
        gait.update(float(t[i]))
        assert list(d[i]) == gait.desired_leg_state and list(s[i]) == gait.leg_state, i
        assert ph[i].tobytes() == np.asarray(gait.normalized_phase, dtype=np.float64).tobytes(), i


@pytest.mark.parametrize("desc", [GHOST, K3LSO])
def test_fk_ik_torque_against_oracle(rg_lib, cuda_device, desc):
    ws = rg.RobotWorkspace(robot_params_from_description(desc), device=cuda_device)
    kin = BatchedKinematics(ws, cuda_device)
    robot = kinematics.OracleRobot(desc)
    rng = np.random.default_rng(2)
    n = 512
    q0 = np.asarray(desc.GetConstants().INIT_MOTOR_ANGLES, dtype=np.float64)
    q = (q0[None] + rng.uniform(-0.35, 0.35, (n, 12))).astype(np.float32)
    feet = kin.ComputeFootPositionsInBaseFrame(_dev(q, cuda_device))
    back = kin.ComputeMotorAnglesFromFootLocalPosition(feet)
    forces = rng.uniform(-60, 60, (n, 12)).astype(np.float32)
    tau = kin.MapContactForceToJointTorques(_dev(forces, cuda_device), _dev(q, cuda_device))
    torch.cuda.synchronize()
    feet, back, tau = feet.cpu().numpy(), back.cpu().numpy(), tau.cpu().numpy()
    assert np.abs(back - q).max() < 2e-5                       # IK(FK(q)) = q  (float32 storage)
    for i in range(0, n, 16):
        ref_feet = robot.fk_all(q[i].astype(np.float64))
        assert np.abs(feet[i].reshape(4, 3) - ref_feet).max() < 1e-6
        robot.set_state(motor_angles=q[i].astype(np.float64))
        for leg in range(4):
            t = robot.MapContactForceToJointTorques(leg, forces[i, 3 * leg:3 * leg + 3].astype(np.float64))
            ref = np.array([t[3 * leg + j] for j in range(3)])
            assert np.abs(tau[i, 3 * leg:3 * leg + 3] - ref).max() < 1e-4 * max(1.0, np.abs(ref).max())
            idx, ang = robot.ComputeMotorAnglesFromFootLocalPosition(leg, feet[i].reshape(4, 3)[leg].astype(np.float64))
            assert np.abs(back[i, idx] - np.array(ang)).max() < 2e-5
    # leg_mask leaves unmasked legs untouched
    mask = np.zeros((n, 4), dtype=np.uint8); mask[:, 1] = 1
    out = torch.full((n, 12), 9.0, dtype=torch.float32, device=cuda_device)
    kin.ComputeMotorAnglesFromFootLocalPosition(_dev(feet, cuda_device), leg_mask=_dev(mask, cuda_device), out=out)
    o = out.cpu().numpy()
    assert np.all(o[:, [0, 1, 2, 6, 7, 8, 9, 10, 11]] == 9.0) and np.abs(o[:, 3:6] - q[:, 3:6]).max() < 2e-5


def test_hybrid_motor_kernel_matches_reference_outputs(rg_lib, cuda_device, reference_constants):
    kat = reference_constants["hybrid_motor_kat"]      # outputs of the reference's own RobotMotorModel
    a = _dev(np.array(kat["commands"], dtype=np.float32), cuda_device)
    q = _dev(np.array(kat["q"], dtype=np.float32), cuda_device)
    qd = _dev(np.array(kat["qd"], dtype=np.float32), cuda_device)
    tau = torch.empty_like(q)
    p = ctypes.c_void_p
    rg.check(rg_lib.rg_hybrid_motor_torque(a.shape[0], p(a.data_ptr()), p(q.data_ptr()), p(qd.data_ptr()), p(tau.data_ptr()), None))
    torch.cuda.synchronize()
    np.testing.assert_allclose(tau.cpu().numpy(), np.array(kat["torque"]), rtol=2e-5, atol=2e-3)


class _SeqRobot(SyntheticRobotBatch):
    pass


def test_control_step_matches_oracle_golden_over_40_steps(rg_lib, cuda_device, golden_dir):
    """BatchedMPCController (the drop-in) stepped 40 times against the frozen output of the restated
    LocomotionController: leg states / phases bit-exact, hybrid actions within tolerance."""
    g = np.load(os.path.join(golden_dir, "control_step_oracle_golden.npz"))
    n_env, n_steps = int(g["n_env"]), int(g["n_steps"])
    seq = synthetic.make_state_sequence(n_env, n_steps, GHOST)
    robot = SyntheticRobotBatch(GHOST, seq[0], device=cuda_device)
    ctl = BatchedMPCController(robot, robot.GetTimeSinceReset)          # reset() at seq[0]'s clock
    assert ctl.MOTOR_CONTROL_MODE == 3 and ctl.get_standing_action() == (0., 0.)
    ctrl = GHOST.GetCtrlConstants()
    worst_force = 0.0
    for k in range(n_steps):
        robot.load(seq[k])
        cmd = torch.from_numpy(seq[k].command).to(cuda_device)
        raw = cmd - torch.tensor([ctrl.VX_OFFSET, ctrl.VY_OFFSET, ctrl.WZ_OFFSET], dtype=torch.float32, device=cuda_device)
        ctl.update_controller_params(raw)
        ctl.command.copy_(cmd)                                          # exactly the float32 commands the golden used
        action = ctl.get_action()
        torch.cuda.synchronize()
        np.testing.assert_array_equal(ctl.desired_leg_state.cpu().numpy(), g["desired_leg_state"][k])
        np.testing.assert_array_equal(ctl.leg_state.cpu().numpy(), g["leg_state"][k])
        assert ctl.normalized_phase.cpu().numpy().tobytes() == g["normalized_phase"][k].tobytes()
        np.testing.assert_allclose(ctl.com_velocity_body.cpu().numpy(), g["com_velocity_body"][k], rtol=0, atol=2e-7)
        f, ref_f = ctl.contact_forces.cpu().numpy(), g["contact_forces"][k]
        worst_force = max(worst_force, np.abs(f - ref_f).max() / max(1.0, np.abs(ref_f).max()))
        a, ref = action.cpu().numpy().reshape(n_env, 12, 5), g["actions"][k].reshape(n_env, 12, 5)
        np.testing.assert_array_equal(a[:, :, [1, 2, 3]], ref[:, :, [1, 2, 3]])     # kp, qdot, kd
        assert np.abs(a[:, :, 0] - ref[:, :, 0]).max() < 5e-5                        # swing joint targets [rad]
        assert np.abs(a[:, :, 4] - ref[:, :, 4]).max() < 1e-4 * max(1.0, np.abs(ref[:, :, 4]).max())
    assert worst_force < 1e-4


def test_controller_interface_single_env_and_reset(rg_lib, cuda_device):
    st = synthetic.make_states(1, GHOST, seed=4)
    robot = SyntheticRobotBatch(GHOST, st, device=cuda_device)
    clock = {"t": 0.0}
    ctl = BatchedMPCController(robot, lambda: clock["t"])
    ctl.update_controller_params((0.2, 0.1))                            # (vx, wz), vy = 0 (mpc_controller.py:84-86)
    c = GHOST.GetCtrlConstants()
    np.testing.assert_allclose(ctl.command.cpu().numpy()[0], [0.2 + c.VX_OFFSET, c.VY_OFFSET, 0.1 + c.WZ_OFFSET], atol=1e-7)
    ctl._mpc_controller.stance_leg_controller.desired_twisting_speed = 0.3
    assert float(ctl._mpc_controller.swing_leg_controller.desired_twisting_speed[0]) == pytest.approx(0.3)
    clock["t"] = 0.06
    a = ctl.get_action()
    assert isinstance(a, np.ndarray) and a.shape == (60,) and a.dtype == np.float32     # what ApplyStepAction takes
    assert ctl.kinematics_model is not None
    assert ctl.swing_joint_valid.cpu().numpy().sum() == 2                # legs 0/3 swing at t = 0.06
    ctl.reset()
    assert ctl.swing_joint_valid.cpu().numpy().sum() == 0
    assert float(ctl.reset_time[0]) == 0.06 and int(ctl.last_leg_state[0, 0]) == -1
    with pytest.raises(ValueError):
        ctl.update_controller_params((1.0,))


def test_controller_batched_subset_reset_and_stats(rg_lib, cuda_device):
    st = synthetic.make_states(256, GHOST, seed=8)
    robot = SyntheticRobotBatch(GHOST, st, device=cuda_device)
    t = torch.zeros(256, dtype=torch.float64, device=cuda_device)
    ctl = BatchedMPCController(robot, lambda: t)
    t += 0.2
    ctl.update_controller_params(torch.tensor([[0.1, 0.0, 0.2]] * 256))
    a = ctl.get_action()
    assert a.shape == (256, 60) and a.is_cuda
    ids = torch.tensor([3, 7], device=cuda_device)
    ctl.reset(ids)
    assert ctl.reset_time[3].item() == 0.2 and ctl.reset_time[4].item() == 0.0
    stats = ctl.rollout_stats().cpu().numpy()
    assert stats[0] == 256 and stats[4] == 256 and stats[6] == 0
    # most solves verify from the cold-start active-set iteration (0 interior-point iterations)
    assert 0 <= stats[1] / 256 <= 12 and 1 <= stats[3] / 256 <= 12


@pytest.mark.parametrize("desc", [GHOST, K3LSO])
def test_state_provider_from_raw_sim_state_matches_oracle(rg_lib, cuda_device, desc):
    """rg_state_from_sim: rpy, body-frame angular velocity, motor angles and FK foot
    positions from raw rigid-body state, against the oracle's restatement of the Robot getters -- including
    envs at both gimbal-lock branches -- and a controller driven through the provider equals one driven
    through the synthetic getter source fed with the oracle's conversions."""
    n = 300
    rng = np.random.default_rng(5)
    rpy = np.column_stack([rng.uniform(-3.0, 3.0, n), rng.uniform(-1.5, 1.5, n), rng.uniform(-3.0, 3.0, n)])
    rpy[0] = (0.0, np.pi / 2, 0.4); rpy[1] = (0.0, -np.pi / 2, -1.1); rpy[2] = 0.0
    quat = synthetic.euler_to_quat_xyzw(rpy).astype(np.float32)
    w_world = rng.uniform(-2, 2, (n, 3)).astype(np.float32)
    v_world = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    oracle_robot = kinematics.OracleRobot(desc)
    motor0 = np.asarray(desc.GetConstants().INIT_MOTOR_ANGLES, dtype=np.float64)
    motor = motor0[None] + rng.uniform(-0.3, 0.3, (n, 12))
    joints = np.stack([oracle_robot.joint_angles(m) for m in motor]).astype(np.float32)
    contacts = (rng.uniform(0, 1, (n, 4)) < 0.6).astype(np.uint8)
    robot = SimStateRobotBatch(desc, n, device=cuda_device)
    t = torch.zeros(n, dtype=torch.float64, device=cuda_device)
    robot.set_sim_state(t, _dev(quat, cuda_device), _dev(v_world, cuda_device), _dev(w_world, cuda_device),
                        _dev(joints, cuda_device), _dev(contacts, cuda_device))
    torch.cuda.synchronize()
    g_rpy, g_rate = robot.GetBaseRollPitchYaw().cpu().numpy(), robot.GetBaseRollPitchYawRate().cpu().numpy()
    g_motor, g_feet = robot.GetMotorAngles().cpu().numpy(), robot.GetFootPositionsInBaseFrame().cpu().numpy()
    for i in range(n):
        q64 = quat[i].astype(np.float64)
        oracle_robot.set_sim_state(q64, v_world[i], w_world[i].astype(np.float64), joints[i].astype(np.float64), contacts[i])
        ref_rpy = np.array(oracle_robot.GetBaseRollPitchYaw())
        # compare rotations, not angles (float32 quaternions near +-pi wrap; gimbal envs fold roll into yaw)
        assert np.abs(kinematics.quat_to_matrix(synthetic.euler_to_quat_xyzw(g_rpy[i:i + 1].astype(np.float64))[0])
                      - kinematics.quat_to_matrix(q64 / np.linalg.norm(q64))).max() < 2e-5
        if abs(abs(ref_rpy[1]) - np.pi / 2) > 1e-2:
            d = np.abs(g_rpy[i] - ref_rpy); d = np.minimum(d, 2 * np.pi - d)
            assert d.max() < 5e-6
        assert np.abs(g_rate[i] - oracle_robot.GetBaseRollPitchYawRate()).max() < 2e-6
        assert np.abs(g_motor[i] - oracle_robot.GetMotorAngles()).max() < 1e-6
        assert np.abs(g_feet[i] - oracle_robot.GetFootPositionsInBaseFrame()).max() < 1e-6
    assert g_rpy[0, 0] == 0.0 and abs(g_rpy[0, 1] - np.pi / 2) < 1e-6          # gimbal branches taken
    assert g_rpy[1, 0] == 0.0 and abs(g_rpy[1, 1] + np.pi / 2) < 1e-6
    # NULL outputs are skipped, NULL inputs rejected
    lib = rg.load()
    assert lib.rg_state_from_sim(robot._ws.ptr, n, quat.ctypes.data, None, joints.ctypes.data, None, None, None, None, None) == -1
    only_feet = torch.zeros((n, 12), dtype=torch.float32, device=cuda_device)
    rg.check(lib.rg_state_from_sim(robot._ws.ptr, n, rg._ptr(_dev(quat, cuda_device), torch.float32, (4,)), None,
                                   rg._ptr(_dev(joints, cuda_device), torch.float32, (12,)), None, None, None,
                                   rg._ptr(only_feet, torch.float32, (12,)), None))
    torch.cuda.synchronize()
    assert np.array_equal(only_feet.cpu().numpy(), g_feet.reshape(n, 12))
    assert lib.rg_state_from_sim(robot._ws.ptr, 0, None, None, None, None, None, None, None, None) == 0   # empty batch

    # the provider drives the controller exactly like the synthetic getter source with the same derived state
    small = n if n < 64 else 64
    sl = slice(0, small)
    states = synthetic.SyntheticStates(
        time_since_reset=np.zeros(small), foot_contacts=contacts[sl], base_velocity_world=v_world[sl],
        base_orientation_xyzw=quat[sl], base_rpy=g_rpy[sl], base_rpy_rate=g_rate[sl],
        foot_positions_base=g_feet[sl].reshape(small, 12), motor_angles=g_motor[sl],
        com_velocity_body=np.zeros((small, 3), np.float32), planned_contacts=contacts[sl], command=np.zeros((small, 3), np.float32),
        com_height=np.zeros(small, np.float32))
    ref_robot = SyntheticRobotBatch(desc, states, device=cuda_device)
    sim_robot = SimStateRobotBatch(desc, small, device=cuda_device)
    tt = torch.full((small,), 0.05, dtype=torch.float64, device=cuda_device)
    sim_robot.set_sim_state(tt, _dev(quat[sl], cuda_device), _dev(v_world[sl], cuda_device), _dev(w_world[sl], cuda_device),
                            _dev(joints[sl], cuda_device), _dev(contacts[sl], cuda_device))
    ref_robot.time_since_reset = tt
    a_ref = BatchedMPCController(ref_robot, ref_robot.GetTimeSinceReset)
    a_sim = BatchedMPCController(sim_robot, sim_robot.GetTimeSinceReset)
    for c in (a_ref, a_sim):
        c.update_controller_params(torch.tensor([[0.2, 0.0, 0.1]] * small))
    out_ref, out_sim = a_ref.get_action().cpu().numpy(), a_sim.get_action().cpu().numpy()
    assert np.array_equal(out_ref, out_sim)


def test_controller_warm_start_matches_cold_controller_over_a_rollout(rg_lib, cuda_device):
    """BatchedMPCController(warm_start=True) (the default) and warm_start=False produce the same actions over a
    rollout with advancing gait and drifting state; the warm one needs fewer active-set rounds."""
    n = 256
    st = synthetic.make_states(n, GHOST, seed=9)
    robots = [SyntheticRobotBatch(GHOST, st, device=cuda_device) for _ in range(2)]
    clock = torch.zeros(n, dtype=torch.float64, device=cuda_device)
    ctl = [BatchedMPCController(robots[0], lambda: clock, warm_start=True), BatchedMPCController(robots[1], lambda: clock, warm_start=False)]
    assert ctl[0].mpc_active_set is not None and ctl[1].mpc_active_set is None
    cmd = torch.tensor([[0.2, 0.0, 0.1]] * n)
    for c in ctl:
        c.update_controller_params(cmd)
    rng = np.random.default_rng(1)
    rounds = np.zeros(2)
    for step in range(30):
        clock += 0.002
        drift = torch.from_numpy(rng.normal(0, 0.003, (n, 3)).astype(np.float32)).to(cuda_device)
        for r in robots:
            r.base_velocity_world += drift
            r.base_rpy[:, :2] += 0.1 * drift[:, :2]
        acts = [c.get_action().clone() for c in ctl]
        torch.cuda.synchronize()
        a0, a1 = acts[0].cpu().numpy(), acts[1].cpu().numpy()
        assert np.abs(a0 - a1).max() < 1e-4 * max(1.0, np.abs(a1).max()), step
        for k, c in enumerate(ctl):
            info = c.solve_info.cpu().numpy()
            assert np.all((info[:, rg.RG_INFO_STATUS] & (rg.RG_STATUS_POLISHED | rg.RG_STATUS_NO_STANCE)) != 0)
            if step > 0:
                rounds[k] += info[:, rg.RG_INFO_POLISH_ROUNDS].mean()
    assert rounds[0] < 0.8 * rounds[1], rounds
    ctl[0].reset(torch.tensor([0, 5], device=cuda_device))
    assert int(ctl[0].mpc_active_set[5, 0]) == -1 and int(ctl[0].mpc_active_set[5].max()) == -1
