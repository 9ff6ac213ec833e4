"""the vectorised gym-env shim (N envs, one process) and the N = 1 adapter that puts a
stock PyBullet-backed ``Robot`` behind the batched controller.  Physics is a pluggable callback (PyBullet is out of
scope and absent offline): the tests use the synthetic joint integrator and a stub of the Robot getter surface."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import kinematics
from robot_gym import cuda as rg
from robot_gym.controllers.mpc.batched_mpc_controller import BatchedMPCController
from robot_gym.core import sim_constants
from robot_gym.gym.batched_env import BatchedRobotGymEnv
from robot_gym.model.robots.descriptions import GHOST, K3LSO
from robot_gym.model.robots.synthetic_robot import SyntheticRobotBatch
from robot_gym.util import synthetic
from robot_gym.util.cli import mapper

pytestmark = pytest.mark.gpu


def test_registration_beside_mpc():
    assert mapper.CONTROLLERS["mpc_cuda"] is BatchedMPCController            # util/cli/mapper.py:7-9
    assert mapper.ROBOTS["ghost"] is GHOST and mapper.ROBOTS["k3lso"] is K3LSO
    assert BatchedMPCController.MOTOR_CONTROL_MODE == 3                      # read at core/simulation.py:113,178


def test_env_shim_4096_envs_100_steps_equals_the_unshimmed_controller(rg_lib, cuda_device):
    """step(actions[N,2]) over 4096 envs x 100 control steps with per-env resets: the hybrid commands the shim applies
    are bit-identical to those of a second, un-shimmed BatchedMPCController reading the same state; clocks advance by
    ACTION_REPEAT ticks per step; fallen envs (no foot in contact, robot_gym_env.py:155-165) are reset individually."""
    n, steps = 4096, 100
    env = BatchedRobotGymEnv(GHOST, n, controller_class="mpc_cuda", device=cuda_device)
    sim = env.simulation
    shadow = BatchedMPCController(sim.robot, sim.GetTimeSinceReset)
    obs = env.reset()
    shadow.reset()
    assert obs.shape == (n, 37) and obs.is_cuda
    gen = torch.Generator(device="cpu").manual_seed(3)
    n_resets = 0
    saw_swing = torch.zeros((), dtype=torch.bool, device=cuda_device)
    max_joint_travel = torch.zeros((), dtype=torch.float32, device=cuda_device)
    for k in range(steps):
        act = torch.stack([torch.rand(n, generator=gen) * 0.35, (torch.rand(n, generator=gen) - 0.5) * 0.8], dim=1).to(cuda_device)
        shadow.update_controller_params(act)
        expect = shadow.get_action().clone()
        if k in (20, 55):                                          # a few envs lose every contact: they must be reset
            sim.physics.forced_airborne[torch.tensor([7, 1000 + k], device=cuda_device)] = True
        t_before = sim.GetTimeSinceReset().clone()
        obs, reward, done, info = env.step(act)
        assert torch.equal(env.last_action, expect), k
        assert obs.shape == (n, 37) and reward.shape == (n,) and done.shape == (n,) and done.dtype == torch.bool
        assert torch.isfinite(obs).all() and torch.isfinite(env.last_action).all()
        fell = done.nonzero().flatten()
        alive = (~done).nonzero().flatten()
        dt = sim_constants.ACTION_REPEAT * sim_constants.SIMULATION_TIME_STEP
        assert torch.allclose(sim.GetTimeSinceReset()[alive], t_before[alive] + dt, atol=1e-12)
        if fell.numel():
            n_resets += fell.numel()
            assert torch.equal(info["reset_env_ids"], fell)
            assert torch.all(sim.GetTimeSinceReset()[fell] == 0) and torch.all(env.episode_steps[fell] == 0)
            shadow.reset(fell)
        assert int(sim.controller.unverified_count()) == 0
        saw_swing |= (sim.robot.GetFootContacts()[alive] == 0).any()
        max_joint_travel = torch.maximum(max_joint_travel, (sim.physics.joint_angles - sim.physics._q0).abs().max())
    assert n_resets >= 4
    # the loop is closed: swing legs lifted (contacts switched off on envs that did not fall) and joints left the start pose
    assert bool(saw_swing) and float(max_joint_travel) > 0.05


@pytest.mark.parametrize("fuse", [True, False])
def test_motor_model_in_the_step_matches_the_reference_formula(rg_lib, cuda_device, fuse):
    """The torque the physics callback receives = strength_ratios * HYBRID motor model * MOTOR_DIRECTION
    (simple_motor.py:128-140, robot.py:291-292), whether the first tick's evaluation is fused into the controller's
    epilogue or launched on its own; strength ratios (set_strength_ratios) act per env and per motor."""
    n = 512
    env = BatchedRobotGymEnv(GHOST, n, device=cuda_device, fuse_motor_model=fuse, auto_reset=False)
    sim = env.simulation
    rng = np.random.default_rng(0)
    ratios = rng.uniform(0.5, 1.0, (n, 12)).astype(np.float32)
    sim.set_strength_ratios(torch.from_numpy(ratios))
    seen = []
    real_step = sim.physics.step

    def spy(torques, dt):
        seen.append((torques.clone(), sim.robot.GetMotorAngles().clone(), sim.robot.GetMotorVelocities().clone()))
        real_step(torques, dt)
    sim.physics.step = spy
    for _ in range(3):
        env.step((0.2, 0.1))
    torch.cuda.synchronize()
    assert len(seen) == 3 * sim_constants.ACTION_REPEAT
    action = env.last_action.cpu().numpy()
    direction = np.asarray(GHOST.GetMotorConstants().MOTOR_DIRECTION)
    for tick in range(sim_constants.ACTION_REPEAT):
        tau, q, qd = (t.cpu().numpy() for t in seen[2 * sim_constants.ACTION_REPEAT + tick])
        for e in range(0, n, 37):
            ref = kinematics.hybrid_motor_torque(action[e], q[e].astype(np.float64), qd[e].astype(np.float64)) * ratios[e] * direction
            assert np.abs(tau[e] - ref).max() < 2e-4 * max(1.0, np.abs(ref).max()), (tick, e)


class _StubBulletRobot:
    """The getter surface of robot.py the controller reads, answering with python lists / numpy like the stock Robot."""

    def __init__(self, desc, st, i):
        self._d, self._st, self._i = desc, st, i

    def GetCtrlConstants(self): return self._d.GetCtrlConstants()
    def GetConstants(self): return self._d.GetConstants()
    def GetMotorConstants(self): return self._d.GetMotorConstants()
    def GetFootContacts(self): return [bool(v) for v in self._st.foot_contacts[self._i]]
    def GetBaseVelocity(self): return tuple(float(v) for v in self._st.base_velocity_world[self._i])
    def GetTrueBaseOrientation(self): return tuple(float(v) for v in self._st.base_orientation_xyzw[self._i])
    def GetBaseRollPitchYaw(self): return np.asarray(self._st.base_rpy[self._i], dtype=np.float64)
    def GetBaseRollPitchYawRate(self): return np.asarray(self._st.base_rpy_rate[self._i], dtype=np.float64)
    def GetFootPositionsInBaseFrame(self): return np.asarray(self._st.foot_positions_base[self._i], dtype=np.float64).reshape(4, 3)
    def GetMotorAngles(self): return np.asarray(self._st.motor_angles[self._i], dtype=np.float64)


class Ghost(_StubBulletRobot):          # the adapter finds the leg chains by the stock class name (ghost.Ghost / k3lso.K3lso)
    pass


def test_single_env_adapter_for_a_stock_robot(rg_lib, cuda_device):
    """controller_class(robot, sim.GetTimeSinceReset) with a stock (non-batched) Robot, as Simulation.build_world does
    (core/simulation.py:117): the controller wraps it in PyBulletRobotAdapter, get_action() returns the [60] float32
    numpy command ApplyStepAction takes, and it equals the command of the batched path on the same state."""
    seq = synthetic.make_state_sequence(1, 12, GHOST, seed=77)
    stub = Ghost(GHOST, seq[0], 0)
    clock = {"t": float(seq[0].time_since_reset[0])}
    ctl = BatchedMPCController(stub, lambda: clock["t"])
    ref_robot = SyntheticRobotBatch(GHOST, seq[0], device=cuda_device)
    ref = BatchedMPCController(ref_robot, ref_robot.GetTimeSinceReset, squeeze_single=False, use_graph=False)   # eager launches
    assert ctl.num_envs == 1 and ctl._adapter is not None
    for k in range(12):
        stub._st = seq[k]
        clock["t"] = float(seq[k].time_since_reset[0])
        ref_robot.load(seq[k])
        for c in (ctl, ref):
            c.update_controller_params((0.2, 0.0, 0.1))
        a = ctl.get_action()
        b = ref.get_action()
        torch.cuda.synchronize()
        assert isinstance(a, np.ndarray) and a.shape == (60,) and a.dtype == np.float32
        np.testing.assert_array_equal(a, b.cpu().numpy()[0])
    ctl.reset()
    assert float(ctl.reset_time[0]) == clock["t"]


@pytest.mark.parametrize("n", [1, 700])
def test_graph_replay_of_the_control_step_equals_eager_launches(rg_lib, cuda_device, n):
    """Small batches replay the step as ONE CUDA graph (rg_control_step_graph_*): same commands, bit for bit, as the
    three eager launches, over a rollout whose state is updated in place; a provider that hands out fresh tensors
    every step makes the controller fall back to eager launches instead of re-capturing for ever."""
    seq = synthetic.make_state_sequence(n, 16, GHOST, seed=123)
    robots = [SyntheticRobotBatch(GHOST, seq[0], device=cuda_device) for _ in range(2)]
    ctl_g = BatchedMPCController(robots[0], robots[0].GetTimeSinceReset, squeeze_single=False, use_graph=True)
    ctl_e = BatchedMPCController(robots[1], robots[1].GetTimeSinceReset, squeeze_single=False, use_graph=False)
    launches = []
    for k in range(16):
        for r in robots:
            r.load(seq[k])                                   # in place: pointers stay put
        for c in (ctl_g, ctl_e):
            c.update_controller_params((0.25, 0.0, -0.1))
        before = rg.launch_count()
        a_g = ctl_g.get_action().clone()
        launches.append(rg.launch_count() - before)
        a_e = ctl_e.get_action().clone()
        torch.cuda.synchronize()
        assert torch.equal(a_g, a_e), k
        assert torch.equal(ctl_g.solve_info, ctl_e.solve_info)
    assert ctl_g._graph is not None and ctl_g._use_graph
    assert launches[0] >= 6 and all(l == 3 for l in launches[1:])          # create = eager step + capture; replay = 3 kernels
    # fresh tensors every step -> stale pointers -> after a few re-captures the controller goes eager
    for k in range(6):
        robots[0].base_rpy = robots[0].base_rpy.clone()
        ctl_g.get_action()
    assert not ctl_g._use_graph and ctl_g._graph is None
