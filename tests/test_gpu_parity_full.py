"""GPU parity at BASELINE's full sizes and over the whole config-4 grid, through the C ABI.

  * every env of the 4096- and 65536-env batches against the C oracle (dense reference-style build + interior point +
    exact polish, pinned to the numpy oracle at 1e-12 in tests/test_c_oracle.py); an env the two disagree on beyond
    1e-5 is re-solved by the numpy oracle, and nothing may exceed 1e-4 relative (north_star's tolerance);
  * all 12 cells of horizon {5, 10, 20} x schedule {trot, pace, bound, walk}, 256 envs each, whole horizon;
  * a SOLVER-INDEPENDENT pin: the dense QP is built by oracle.convex_mpc.build_qp and the kernel's float64
    solution must satisfy its KKT conditions (stationarity, primal feasibility, multiplier signs) -- no oracle
    solver is involved in that check;
  * the full control step over 256 envs x 10 control steps (ghost) and 96 x 8 (k3lso) against the restated
    LocomotionController, and at 65536 envs against a 700-env controller over a subset of the same envs (every env is
    independent of its batch: gait bit-identical, forces and actions equal to rounding);
  * six random parameter sets (body, weights, regularisation, planning step, friction coefficients, horizon, schedule),
    384 envs each, every env against the C oracle built from the same parameters;
  * the standalone entry points the fused step does not exercise (rg_swing_targets, rg_com_velocity_update,
    rg_pack_hybrid_action, plain rg_mpc_build_solve), stream re-entrancy, the two-kernel / one-kernel solve.

"Oracle" = the in-repo restatement; reference parity is unpinned (oracle/__init__.py)."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import c_oracle, convex_mpc as cm, kinematics, locomotion
from robot_gym import cuda as rg
from robot_gym.controllers.mpc.batched_kinematics import robot_params_from_description
from robot_gym.controllers.mpc.batched_mpc_controller import BatchedMPCController
from robot_gym.model.robots.descriptions import GHOST, K3LSO, with_gait
from robot_gym.model.robots.synthetic_robot import SyntheticRobotBatch
from robot_gym.util import synthetic

pytestmark = pytest.mark.gpu
REL_TOL = 1e-4            # north_star: forces within 1e-4 relative of the oracle optimum
RECHECK = 1e-5            # C oracle vs kernel beyond this: the numpy oracle arbitrates
CORES = os.cpu_count() or 1


def _dev(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _params(horizon=10, **overrides):
    ctrl = GHOST.GetCtrlConstants()
    p = rg.default_mpc_params(ctrl.MPC_BODY_MASS, ctrl.MPC_BODY_INERTIA, ctrl.MPC_BODY_HEIGHT, horizon)
    for k, v in overrides.items():
        setattr(p, k, v)
    return p


def _solve(dev, st, horizon=10, want_horizon=False, f64=False, ws=None, **overrides):
    ws = ws or rg.MpcWorkspace(_params(horizon, **overrides), device=dev, max_envs=max(1, len(st)))
    h64 = torch.empty((len(st), horizon, 12), dtype=torch.float64, device=dev) if f64 else None
    f, hf, info = rg.mpc_build_solve(ws, _dev(st.com_velocity_body, dev), _dev(st.base_rpy, dev), _dev(st.base_rpy_rate, dev),
                                     _dev(st.planned_contacts, dev), _dev(st.foot_positions_base, dev), _dev(st.command, dev),
                                     want_horizon=want_horizon, horizon_forces_f64=h64)
    torch.cuda.synchronize()
    return (f.cpu().numpy(), None if hf is None else hf.cpu().numpy(), info.cpu().numpy(),
            None if h64 is None else h64.cpu().numpy())


def _numpy_oracle(st, i, horizon):
    ctrl = GHOST.GetCtrlConstants()
    return cm.compute_contact_forces(
        cm.MpcParams(horizon=horizon), st.com_velocity_body[i].astype(np.float64), st.base_rpy[i].astype(np.float64),
        st.base_rpy_rate[i].astype(np.float64), st.planned_contacts[i], st.foot_positions_base[i].astype(np.float64),
        [0, 0, ctrl.MPC_BODY_HEIGHT], [float(st.command[i, 0]), float(st.command[i, 1]), 0.0], [0, 0, 0],
        [0, 0, float(st.command[i, 2])])


def _assert_all_envs_match(gpu, ref, st, horizon, what):
    """gpu / ref: [N, m] forces (first step or whole horizon).  Every env must agree."""
    n = len(gpu)
    scale = np.maximum(1.0, np.abs(ref).reshape(n, -1).max(axis=1))
    rel = np.abs(gpu - ref).reshape(n, -1).max(axis=1) / scale
    suspects = np.flatnonzero(rel > RECHECK)
    assert len(suspects) <= max(8, n // 200), (what, len(suspects), float(rel.max()))    # the C oracle is exact to ~1e-7
    for i in suspects:
        exact = _numpy_oracle(st, int(i), horizon)[:gpu.shape[1]]
        rel[i] = np.abs(gpu[i] - exact).max() / max(1.0, np.abs(exact).max())
    assert rel.max() < REL_TOL, (what, int(rel.argmax()), float(rel.max()))
    return float(rel.max()), len(suspects)


@pytest.mark.parametrize("n", [4096, 65536])
def test_every_env_of_the_full_batches_matches_the_oracle(rg_lib, cuda_device, n):
    """BASELINE config[1] (4096 envs) and the config[2] batch size (65536): first-step forces of EVERY env."""
    st = synthetic.make_states(n, GHOST)
    f, _, info, _ = _solve(cuda_device, st)
    ref, _, _ = c_oracle.solve_batch(cm.MpcParams(), st, GHOST.GetCtrlConstants().MPC_BODY_HEIGHT, n_threads=CORES)
    worst, n_suspect = _assert_all_envs_match(f, ref, st, 10, f"{n} envs")
    status = info[:, rg.RG_INFO_STATUS]
    assert np.all(status & (rg.RG_STATUS_POLISHED | rg.RG_STATUS_NO_STANCE))
    print(f"[{n} envs] worst relative force error vs oracle {worst:.2e}; {n_suspect} envs re-checked with the numpy oracle")


@pytest.mark.parametrize("schedule", ["trot", "pace", "bound", "walk"])
@pytest.mark.parametrize("horizon", [5, 10, 20])
def test_config4_grid_every_cell_against_the_oracle(rg_lib, cuda_device, horizon, schedule):
    """BASELINE config[3]: horizon x contact schedule, 256 envs per cell, the WHOLE horizon of forces."""
    desc = with_gait(GHOST, schedule)
    n = 256
    st = synthetic.make_states(n, desc, schedule_ctrl=desc.GetCtrlConstants(), seed=100 * horizon + len(schedule))
    f, hf, info, _ = _solve(cuda_device, st, horizon, want_horizon=True)
    _, ref, _ = c_oracle.solve_batch(cm.MpcParams(horizon=horizon), st, GHOST.GetCtrlConstants().MPC_BODY_HEIGHT,
                                     n_threads=CORES, want_horizon=True)
    _assert_all_envs_match(hf.reshape(n, -1), ref.reshape(n, -1), st, horizon, f"h={horizon} {schedule}")
    np.testing.assert_array_equal(f, hf[:, 0, :])
    polished = (info[:, rg.RG_INFO_STATUS] & (rg.RG_STATUS_POLISHED | rg.RG_STATUS_NO_STANCE)) != 0
    assert polished.mean() >= 0.99, (horizon, schedule, polished.mean())


@pytest.mark.parametrize("schedule", ["trot", "pace", "bound", "walk"])
def test_kkt_certificate_of_the_kernel_solution(rg_lib, cuda_device, schedule):
    """Solver-independent: build the dense QP the way mpc_osqp does (oracle.convex_mpc.build_qp) and check that the
    kernel's float64 solution is a KKT point of it -- stationarity with multipliers recovered by least squares on
    the rows it sits on, primal feasibility, multiplier signs.  P is positive definite, so a KKT point is THE optimum.
    Tolerances are relative to the gradient scale max(1, |q|_inf) (stationarity, signs) and to fz_max (feasibility)."""
    desc = with_gait(GHOST, schedule)
    ctrl = desc.GetCtrlConstants()
    n = 512
    st = synthetic.make_states(n, desc, schedule_ctrl=ctrl, seed=211)
    _, _, info, h64 = _solve(cuda_device, st, want_horizon=False, f64=True)
    mp = cm.MpcParams()
    polished = (info[:, rg.RG_INFO_STATUS] & rg.RG_STATUS_POLISHED) != 0
    worst = dict(stationarity=0.0, primal=0.0, dual_sign=0.0)
    for i in range(n):
        if not st.planned_contacts[i].any():
            assert np.all(h64[i] == 0.0)
            continue
        qp = cm.build_qp(mp, st.com_velocity_body[i].astype(np.float64), st.base_rpy[i].astype(np.float64),
                         st.base_rpy_rate[i].astype(np.float64), st.planned_contacts[i],
                         st.foot_positions_base[i].astype(np.float64), [0, 0, ctrl.MPC_BODY_HEIGHT],
                         [float(st.command[i, 0]), float(st.command[i, 1]), 0.0], [0, 0, 0], [0, 0, float(st.command[i, 2])])
        x = -h64[i].reshape(-1)                        # the kernel returns the negated solution, like mpc_osqp
        cert = cm.kkt_certificate(qp.p_mat, qp.q_vec, qp.c_mat, qp.lb, qp.ub, x, act_tol=1e-8)
        qs = max(1.0, float(np.abs(qp.q_vec).max()))
        tol = 1e-6 if polished[i] else 1e-3            # an unverified solve is the best interior-point iterate
        assert cert["stationarity"] <= tol * qs, (schedule, i, cert["stationarity"], qs)
        assert cert["primal"] <= 1e-8 * mp.fz_max, (schedule, i, cert["primal"])
        assert cert["dual_sign"] <= tol * qs, (schedule, i, cert["dual_sign"])
        if polished[i]:
            worst["stationarity"] = max(worst["stationarity"], cert["stationarity"] / qs)
            worst["dual_sign"] = max(worst["dual_sign"], cert["dual_sign"] / qs)
            worst["primal"] = max(worst["primal"], cert["primal"] / mp.fz_max)
    assert polished.mean() >= 0.99
    print(f"[{schedule}] KKT certificate of {int(polished.sum())} verified solves: " +
          ", ".join(f"{k} {v:.1e}" for k, v in worst.items()))


def test_two_kernel_and_one_kernel_solves_agree_and_the_queue_is_reusable(rg_lib, cuda_device):
    """two_kernel_solve = 1 (lean active-set kernel + complete solver on the queued envs) and = 0 (one complete
    kernel) reach the same verified optimum; the bound gait exercises the queue, repeated solves on ONE workspace
    show that the fallback kernel re-arms the queue, and a batch larger than the workspace's queue falls back to
    the single kernel."""
    desc = with_gait(GHOST, "bound")
    st = synthetic.make_states(1024, desc, schedule_ctrl=desc.GetCtrlConstants(), seed=77)
    ws2 = rg.MpcWorkspace(_params(), device=cuda_device, max_envs=1024)
    f1, hf1, info1, _ = _solve(cuda_device, st, want_horizon=True, two_kernel_solve=0)
    for rep in range(3):
        f2, hf2, info2, _ = _solve(cuda_device, st, want_horizon=True, ws=ws2)
        scale = np.maximum(1.0, np.abs(hf1).reshape(1024, -1).max(axis=1))
        gap = np.abs(hf1 - hf2).reshape(1024, -1).max(axis=1) / scale
        both = ((info1[:, 2] & info2[:, 2]) & rg.RG_STATUS_POLISHED) != 0
        assert both.mean() > 0.99 and gap[both].max() < 1e-5, (rep, both.mean(), gap[both].max())
        assert gap.max() < 1e-3
        queued = (info2[:, rg.RG_INFO_STATUS] & rg.RG_STATUS_ACTIVE_SET_ONLY) == 0
        assert 0.02 < queued.mean() < 0.5                        # the fallback queue was used ...
        assert np.all(info2[queued, rg.RG_INFO_IPM_ITERS] > 0)   # ... and those envs went through the interior point
    small_ws = rg.MpcWorkspace(_params(), device=cuda_device, max_envs=16)
    f3, hf3, info3, _ = _solve(cuda_device, st, want_horizon=True, ws=small_ws)
    np.testing.assert_array_equal(hf3, hf1)                      # same single kernel, bitwise
    # trot: nothing is queued, results are bit-identical between the two modes
    st_t = synthetic.make_states(2048, GHOST, seed=78)
    a, _, ia, _ = _solve(cuda_device, st_t, two_kernel_solve=0)
    b, _, ib, _ = _solve(cuda_device, st_t, two_kernel_solve=1)
    np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(ia, ib)


def test_plain_build_solve_entry_point_and_zero_yaw(rg_lib, cuda_device):
    """rg_mpc_build_solve called directly (no warm-start buffer) equals the warm entry point with an unknown seed;
    zero_yaw through rg_mpc_build_solve_io equals passing yaw = 0."""
    st = synthetic.make_states(300, GHOST, seed=12)
    st.base_rpy[:, 2] = np.linspace(-0.5, 0.5, 300).astype(np.float32)
    dev = cuda_device
    ws = rg.MpcWorkspace(_params(), device=dev, max_envs=300)
    t = {k: _dev(getattr(st, k), dev) for k in ("com_velocity_body", "base_rpy", "base_rpy_rate", "planned_contacts",
                                                  "foot_positions_base", "command")}
    p = lambda x: ctypes.c_void_p(x.data_ptr())
    f_plain = torch.empty((300, 12), dtype=torch.float32, device=dev)
    info_plain = torch.empty((300, 4), dtype=torch.int32, device=dev)
    rg.check(rg_lib.rg_mpc_build_solve(ws.ptr, 300, p(t["com_velocity_body"]), p(t["base_rpy"]), p(t["base_rpy_rate"]),
                                       p(t["planned_contacts"]), p(t["foot_positions_base"]), p(t["command"]), None,
                                       p(f_plain), None, p(info_plain), None))
    f_warm, _, info_warm = rg.mpc_build_solve(ws, *t.values(), active_set=rg.new_active_set(300, 10, dev))
    torch.cuda.synchronize()
    assert torch.equal(f_plain, f_warm) and torch.equal(info_plain, info_warm)
    ref = _numpy_oracle(st, 17, 10)[:12]
    assert np.abs(f_plain[17].cpu().numpy() - ref).max() < REL_TOL * max(1.0, np.abs(ref).max())
    f_zero, _, _ = rg.mpc_build_solve(ws, *t.values(), zero_yaw=True)
    rpy0 = t["base_rpy"].clone(); rpy0[:, 2] = 0
    f_ref, _, _ = rg.mpc_build_solve(ws, t["com_velocity_body"], rpy0, *list(t.values())[2:])
    torch.cuda.synchronize()
    assert torch.equal(f_zero, f_ref) and not torch.equal(f_zero, f_plain)


def test_solves_on_two_streams_are_independent(rg_lib, cuda_device):
    """Re-entrancy claimed by include/rg_cuda.h: two solves in flight on two streams with disjoint workspaces and
    outputs give exactly the results of the same solves run one after the other."""
    dev = cuda_device
    sts = [synthetic.make_states(3000, GHOST, seed=s) for s in (21, 22)]
    wss = [rg.MpcWorkspace(_params(), device=dev, max_envs=3000) for _ in sts]
    ins = [[_dev(getattr(st, k), dev) for k in ("com_velocity_body", "base_rpy", "base_rpy_rate", "planned_contacts",
                                                  "foot_positions_base", "command")] for st in sts]
    serial = [rg.mpc_build_solve(ws, *a)[0].clone() for ws, a in zip(wss, ins)]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(device=dev) for _ in sts]
    outs = [None, None]
    for rep in range(3):
        for k, (ws, a, s) in enumerate(zip(wss, ins, streams)):
            with torch.cuda.stream(s):
                outs[k] = rg.mpc_build_solve(ws, *a)[0]
        torch.cuda.synchronize()
        for k in range(2):
            assert torch.equal(outs[k], serial[k]), (rep, k)


def test_stale_workspace_is_reported_not_dereferenced(rg_lib, cuda_device):
    """A workspace overwritten behind the library's back (here: zeroed) must not be indexed with the recorded
    horizon: the kernel flags RG_STATUS_BAD_WORKSPACE and returns zero forces."""
    st = synthetic.make_states(64, GHOST, seed=5)
    ws = rg.MpcWorkspace(_params(), device=cuda_device, max_envs=64)
    ws.buffer.zero_()
    f, _, info, _ = _solve(cuda_device, st, ws=ws)
    assert np.all(f == 0) and np.all(info[:, rg.RG_INFO_STATUS] == rg.RG_STATUS_BAD_WORKSPACE)


def test_wrapper_rejects_short_misplaced_and_misaligned_tensors(rg_lib, cuda_device):
    st = synthetic.make_states(32, GHOST, seed=6)
    dev = cuda_device
    ws = rg.MpcWorkspace(_params(), device=dev, max_envs=32)
    t = [_dev(getattr(st, k), dev) for k in ("com_velocity_body", "base_rpy", "base_rpy_rate", "planned_contacts",
                                              "foot_positions_base", "command")]
    short = list(t); short[0] = t[0][:16].contiguous()
    with pytest.raises(ValueError, match="rows"):
        rg.mpc_build_solve(ws, *short)
    shifted = list(t)
    raw = torch.zeros(32 * 4 + 1, dtype=torch.uint8, device=dev)
    shifted[3] = raw[1:].view(32, 4)                                   # contact bytes off the 4-byte grid
    with pytest.raises(ValueError, match="aligned"):
        rg.mpc_build_solve(ws, *shifted)
    robot = SyntheticRobotBatch(GHOST, st, device=dev)
    ctl = BatchedMPCController(robot, robot.GetTimeSinceReset)
    robot.base_rpy = robot.base_rpy[:8].contiguous()                   # a provider that returns a sub-batch
    with pytest.raises(ValueError, match="rows"):
        ctl.get_action()


# ------------------------------------------------------------------------------------------------ config 3 at scale
@pytest.mark.parametrize("robot_name,n_env,n_steps", [("ghost", 256, 10), ("k3lso", 96, 8)])
def test_control_step_256_envs_10_steps_against_the_restated_controller(rg_lib, cuda_device, robot_name, n_env, n_steps):
    """BASELINE config[2] in parity form: BatchedMPCController over 256 envs x 10 control steps (ghost; 96 x 8 with the
    k3lso constants) against one restated LocomotionController per env (oracle/locomotion.py with the C port as its QP
    solver): gait states and phases bit-exact, estimator, forces, swing targets and the 60-float hybrid actions within
    tolerance."""
    desc = {"ghost": GHOST, "k3lso": K3LSO}[robot_name]
    ctrl = desc.GetCtrlConstants()
    seq = synthetic.make_state_sequence(n_env, n_steps, desc, seed=synthetic.SEED + 5)
    robot = SyntheticRobotBatch(desc, seq[0], device=cuda_device)
    ctl = BatchedMPCController(robot, robot.GetTimeSinceReset)
    acts, des, sta, pha, frc = [], [], [], [], []
    for k in range(n_steps):
        robot.load(seq[k])
        ctl.command.copy_(torch.from_numpy(seq[k].command).to(cuda_device))
        a = ctl.get_action()
        torch.cuda.synchronize()
        acts.append(a.cpu().numpy().copy()); des.append(ctl.desired_leg_state.cpu().numpy().copy())
        sta.append(ctl.leg_state.cpu().numpy().copy()); pha.append(ctl.normalized_phase.cpu().numpy().copy())
        frc.append(ctl.contact_forces.cpu().numpy().copy())
        assert int(ctl.unverified_count()) == 0
    worst_f = worst_q = worst_tau = 0.0
    for e in range(n_env):
        orobot = kinematics.OracleRobot(desc)
        clock = {"t": 0.0}

        def load(k):
            s = seq[k]
            orobot.set_state(base_velocity=s.base_velocity_world[e].astype(np.float64),
                             base_orientation=s.base_orientation_xyzw[e].astype(np.float64),
                             base_rpy=s.base_rpy[e].astype(np.float64), base_rpy_rate=s.base_rpy_rate[e].astype(np.float64),
                             foot_positions=s.foot_positions_base[e].astype(np.float64), foot_contacts=s.foot_contacts[e],
                             motor_angles=s.motor_angles[e].astype(np.float64))
            clock["t"] = float(s.time_since_reset[e])
        load(0)
        octl = locomotion.build_mpc_controller(orobot, lambda: clock["t"], ctrl, mpc_solver=c_oracle.compute_contact_forces)
        octl.reset()
        for k in range(n_steps):
            load(k)
            s = seq[k]
            for leg_ctl in (octl.swing_leg_controller, octl.stance_leg_controller):
                leg_ctl.desired_speed = [float(s.command[e, 0]), float(s.command[e, 1]), 0.0]
                leg_ctl.desired_twisting_speed = float(s.command[e, 2])
            octl.update()
            ref = octl.get_action().reshape(12, 5)
            assert list(des[k][e]) == octl.gait_generator.desired_leg_state, (e, k)
            assert list(sta[k][e]) == octl.gait_generator.leg_state, (e, k)
            assert pha[k][e].tobytes() == np.asarray(octl.gait_generator.normalized_phase, dtype=np.float64).tobytes(), (e, k)
            rf = octl.stance_leg_controller.last_contact_forces
            worst_f = max(worst_f, np.abs(frc[k][e] - rf).max() / max(1.0, np.abs(rf).max()))
            a = acts[k][e].reshape(12, 5)
            np.testing.assert_array_equal(a[:, [1, 2, 3]], ref[:, [1, 2, 3]])
            worst_q = max(worst_q, np.abs(a[:, 0] - ref[:, 0]).max())
            worst_tau = max(worst_tau, np.abs(a[:, 4] - ref[:, 4]).max() / max(1.0, np.abs(ref[:, 4]).max()))
    assert worst_f < REL_TOL and worst_q < 5e-5 and worst_tau < REL_TOL, (worst_f, worst_q, worst_tau)
    print(f"[control step {robot_name} {n_env} x {n_steps}] worst force {worst_f:.1e} rel, swing joint target {worst_q:.1e} rad, torque {worst_tau:.1e} rel")


# ------------------------------------------------------------------------------------------------ standalone parts
def test_standalone_estimator_swing_and_pack_entry_points(rg_lib, cuda_device):
    """rg_com_velocity_update, rg_swing_targets and rg_pack_hybrid_action called one by one over 8 control steps
    reproduce the restated python classes (COMVelocityEstimator, RaibertSwingLegController up to the IK call,
    LocomotionController's merge) -- the fused rg_control_step never calls these three kernels."""
    dev = cuda_device
    desc, ctrl = GHOST, GHOST.GetCtrlConstants()
    n, n_steps = 96, 8
    rws = rg.RobotWorkspace(robot_params_from_description(desc), device=dev)
    seq = synthetic.make_state_sequence(n, n_steps, desc, seed=synthetic.SEED + 9)
    W = 20
    f32, f64, i32, u8 = torch.float32, torch.float64, torch.int32, torch.uint8
    z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dev)
    window, wsum, wcorr, wcount, whead = z((n, 3, W), f64), z((n, 3), f64), z((n, 3), f64), z((n,), i32), z((n,), i32)
    v_body, v_world = z((n, 3), f32), z((n, 3), f32)
    desired, state, phase = z((n, 4), i32), z((n, 4), i32), z((n, 4), f64)
    last_state = torch.full((n, 4), -1, dtype=i32, device=dev)
    latch = _dev(seq[0].foot_positions_base, dev).clone()
    target = z((n, 12), f32)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    # oracle side: one estimator + gait + swing controller per env
    orobots = [kinematics.OracleRobot(desc) for _ in range(n)]
    stacks = []
    for e in range(n):
        s0 = seq[0]
        orobots[e].set_state(base_velocity=s0.base_velocity_world[e].astype(np.float64),
                             base_orientation=s0.base_orientation_xyzw[e].astype(np.float64),
                             base_rpy=s0.base_rpy[e].astype(np.float64), base_rpy_rate=s0.base_rpy_rate[e].astype(np.float64),
                             foot_positions=s0.foot_positions_base[e].astype(np.float64), foot_contacts=s0.foot_contacts[e],
                             motor_angles=s0.motor_angles[e].astype(np.float64))
        gait = locomotion.OpenloopGaitGenerator(orobots[e], ctrl.STANCE_DURATION_SECONDS, ctrl.DUTY_FACTOR,
                                                ctrl.INIT_PHASE_FULL_CYCLE, ctrl.INIT_LEG_STATE)
        est = locomotion.COMVelocityEstimator(orobots[e], window_size=W)
        sw = locomotion.RaibertSwingLegController(orobots[e], gait, est, desired_speed=(0.0, 0.0), desired_twisting_speed=0.0,
                                                  desired_height=ctrl.MPC_BODY_HEIGHT, foot_clearance=0.01)
        gait.reset(0.0); est.reset(0.0); sw.reset(0.0)
        stacks.append((gait, est, sw))
    t0 = seq[0].time_since_reset.copy()
    worst_v = worst_t = 0.0
    for k in range(n_steps):
        s = seq[k]
        t_dev = _dev(s.time_since_reset - t0, dev)
        contacts = _dev(s.foot_contacts, dev)
        vel, quat = _dev(s.base_velocity_world, dev), _dev(s.base_orientation_xyzw, dev)
        feet, rate, cmd = _dev(s.foot_positions_base, dev), _dev(s.base_rpy_rate, dev), _dev(s.command, dev)
        rg.check(rg_lib.rg_gait_step(rws.ptr, n, P(t_dev), P(contacts), P(desired), P(state), P(phase), None))
        rg.check(rg_lib.rg_com_velocity_update(rws.ptr, n, P(vel), P(quat), P(window), P(wsum), P(wcorr), P(wcount), P(whead),
                                               P(v_body), P(v_world), None))
        rg.check(rg_lib.rg_swing_targets(rws.ptr, n, P(desired), P(state), P(phase), P(feet), P(v_body), P(rate), P(cmd),
                                         P(last_state), P(latch), P(target), None))
        torch.cuda.synchronize()
        vb, vw, tg, st_np = v_body.cpu().numpy(), v_world.cpu().numpy(), target.cpu().numpy(), state.cpu().numpy()
        for e in range(n):
            gait, est, sw = stacks[e]
            orobots[e].set_state(base_velocity=s.base_velocity_world[e].astype(np.float64),
                                 base_orientation=s.base_orientation_xyzw[e].astype(np.float64),
                                 base_rpy=s.base_rpy[e].astype(np.float64), base_rpy_rate=s.base_rpy_rate[e].astype(np.float64),
                                 foot_positions=s.foot_positions_base[e].astype(np.float64), foot_contacts=s.foot_contacts[e],
                                 motor_angles=s.motor_angles[e].astype(np.float64))
            now = float(s.time_since_reset[e] - t0[e])
            gait.update(now); est.update(now); sw.update(now)
            worst_v = max(worst_v, np.abs(vb[e] - np.asarray(est.com_velocity_body_frame)).max(),
                          np.abs(vw[e] - np.asarray(est.com_velocity_world_frame)).max())
            sw.desired_speed = np.array([float(s.command[e, 0]), float(s.command[e, 1]), 0.0])
            sw.desired_twisting_speed = float(s.command[e, 2])
            sw.foot_targets = {}
            sw.get_action()
            targets = sw.foot_targets                          # {leg: base-frame foot target} of the non-stance legs
            assert sorted(targets) == [l for l in range(4) if st_np[e, l] not in (rg.RG_LEG_STANCE, rg.RG_LEG_EARLY_CONTACT)], (e, k)
            for leg, ref in targets.items():
                worst_t = max(worst_t, np.abs(tg[e, 3 * leg:3 * leg + 3] - ref).max())
    assert worst_v < 5e-7 and worst_t < 2e-6, (worst_v, worst_t)
    # pack: swing 5-tuples for legs whose desired state is SWING and that hold a stored IK result, torques elsewhere
    rng = np.random.default_rng(3)
    swing_angles = rng.uniform(-1, 1, (n, 12)).astype(np.float32)
    valid = (rng.uniform(0, 1, (n, 4)) < 0.7).astype(np.uint8)
    torques = rng.uniform(-30, 30, (n, 12)).astype(np.float32)
    action = z((n, 60), f32)
    sa, va, tq = _dev(swing_angles, dev), _dev(valid, dev), _dev(torques, dev)
    rg.check(rg_lib.rg_pack_hybrid_action(rws.ptr, n, P(desired), P(sa), P(va), P(tq), P(action), None))
    torch.cuda.synchronize()
    a = action.cpu().numpy().reshape(n, 12, 5)
    d = desired.cpu().numpy()
    mc = desc.GetMotorConstants()
    for e in range(n):
        for m in range(12):
            leg = m // 3
            if d[e, leg] == rg.RG_LEG_SWING and valid[e, leg]:
                exp = (swing_angles[e, m], np.float32(mc.MOTOR_POSITION_GAINS[m]), 0.0, np.float32(mc.MOTOR_VELOCITY_GAINS[m]), 0.0)
            else:
                exp = (0.0, 0.0, 0.0, 0.0, torques[e, m])
            assert tuple(a[e, m]) == tuple(np.float32(x) for x in exp), (e, m)


def test_weakly_active_row_does_not_cycle(rg_lib, cuda_device):
    """Env 17809 of the prefix-stable batch has a friction row whose multiplier is -1.7e-10 at the optimum (no strict
    complementarity): the active-set rounds used to flip it in and out until the budget ran out (status: interior point
    only).  Rows that come back after being dropped are now sticky; the solve must verify and match the oracle."""
    st = synthetic.make_states_sharded(17809 - 3, 17809 + 3, GHOST)
    for two in (0, 1):
        f, hf, info, _ = _solve(cuda_device, st, want_horizon=True, two_kernel_solve=two)
        assert np.all(info[:, rg.RG_INFO_STATUS] & rg.RG_STATUS_POLISHED), info
        ref = _numpy_oracle(st, 3, 10)
        assert np.abs(hf[3].reshape(-1) - ref).max() < REL_TOL * max(1.0, np.abs(ref).max())



@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4, 5])
def test_random_parameter_sets_every_env_against_the_oracle(rg_lib, cuda_device, seed):
    """The recalled third-party constants are run-time parameters: random bodies (mass, full inertia
    tensor), MPC weights, regularisation, planning step, friction coefficients (one per pyramid row), horizon and
    contact schedule -- 384 envs per set, every env against the C oracle built from the same parameters."""
    rng = np.random.default_rng(1000 + seed)
    horizon = int(rng.choice([5, 10, 20]))
    schedule = str(rng.choice(["trot", "walk", "pace"]))
    mass = float(rng.uniform(8.0, 40.0))
    a = rng.normal(size=(3, 3)) * 0.05
    inertia = (np.diag(rng.uniform(0.05, 0.8, 3)) + a @ a.T)                     # symmetric positive definite
    base_w = np.array([5, 5, 0.2, 0, 0, 10, 0.5, 0.5, 0.2, 0.2, 0.2, 0.1, 0], dtype=np.float64)
    weights = tuple(float(x) for x in base_w * rng.uniform(0.3, 3.0, 13))
    mp = cm.MpcParams(mass=mass, inertia=tuple(float(x) for x in inertia.reshape(-1)), horizon=horizon,
                      dt=float(rng.uniform(0.02, 0.04)), weights=weights, alpha=float(10 ** rng.uniform(-5.5, -4.5)),
                      friction_coeffs=tuple(float(x) for x in rng.uniform(0.3, 0.8, 4)))
    desc = with_gait(GHOST, schedule)
    st = synthetic.make_states(384, desc, schedule_ctrl=desc.GetCtrlConstants(), seed=2000 + seed)
    height = GHOST.GetCtrlConstants().MPC_BODY_HEIGHT
    overrides = dict(mass=mp.mass, inertia=mp.inertia, dt=mp.dt, weights=mp.weights, alpha=mp.alpha,
                     friction_coeffs=mp.friction_coeffs, fz_max=mp.fz_max, fz_min=mp.fz_min)
    f, hf, info, _ = _solve(cuda_device, st, horizon=horizon, want_horizon=True, **overrides)
    assert np.all(info[:, rg.RG_INFO_STATUS] & (rg.RG_STATUS_POLISHED | rg.RG_STATUS_NO_STANCE))
    _, ref = c_oracle.solve_batch(mp, st, height, n_threads=CORES, want_horizon=True)[:2]
    n = len(st)
    gpu, ref = hf.reshape(n, -1), ref.reshape(n, -1)
    rel = np.abs(gpu - ref).max(axis=1) / np.maximum(1.0, np.abs(ref).max(axis=1))
    suspects = np.flatnonzero(rel > RECHECK)
    assert len(suspects) <= 8, (seed, horizon, schedule, len(suspects), float(rel.max()))
    for i in suspects:                                                           # the numpy oracle arbitrates
        i = int(i)
        exact = cm.compute_contact_forces(mp, st.com_velocity_body[i].astype(np.float64), st.base_rpy[i].astype(np.float64),
                                          st.base_rpy_rate[i].astype(np.float64), st.planned_contacts[i],
                                          st.foot_positions_base[i].astype(np.float64), [0, 0, height],
                                          [float(st.command[i, 0]), float(st.command[i, 1]), 0.0], [0, 0, 0],
                                          [0, 0, float(st.command[i, 2])])
        rel[i] = np.abs(gpu[i] - exact).max() / max(1.0, np.abs(exact).max())
    assert rel.max() < REL_TOL, (seed, horizon, schedule, int(rel.argmax()), float(rel.max()))


PROLOGUE_OUTPUTS = ("com_velocity_body", "desired_leg_state", "leg_state", "normalized_phase", "mpc_contact_state", "swing_foot_target",
                    "swing_joint_angles", "swing_joint_valid", "last_leg_state", "phase_switch_foot_local_position")


def test_control_step_at_65536_envs_is_batch_size_independent(rg_lib, cuda_device):
    """BASELINE config[2] at its full size: three control steps of one BatchedMPCController over 65536 envs (lean
    kernel + fallback queue, four launches per step) against a second controller that only sees 700 of those envs
    (one wave: the single complete kernel, one graph launch).  Every env is independent of its batch: gait states and
    phases must be bit-identical, forces and actions equal to rounding, and nothing may be left unverified.  The
    small controller's envs are the ones the 256-env test pins to the restated oracle controller."""
    import dataclasses
    n_env, n_steps, n_sub = 65536, 3, 700
    seq = synthetic.make_state_sequence(n_env, n_steps, GHOST, seed=synthetic.SEED + 9)
    idx = np.sort(np.random.default_rng(7).choice(n_env, n_sub, replace=False))
    sub = [type(s)(**{f.name: getattr(s, f.name)[idx] for f in dataclasses.fields(s)}) for s in seq]
    out = []
    for states in (seq, sub):
        robot = SyntheticRobotBatch(GHOST, states[0], device=cuda_device)
        ctl = BatchedMPCController(robot, robot.GetTimeSinceReset)
        rec = []
        for k in range(n_steps):
            robot.load(states[k])
            ctl.command.copy_(torch.from_numpy(states[k].command).to(cuda_device))
            a = ctl.get_action()
            torch.cuda.synchronize()
            assert int(ctl.unverified_count()) == 0
            pro = [getattr(ctl, name).cpu().numpy().copy() for name in PROLOGUE_OUTPUTS]
            rec.append((a.cpu().numpy().copy(), ctl.leg_state.cpu().numpy().copy(), ctl.normalized_phase.cpu().numpy().copy(),
                        ctl.contact_forces.cpu().numpy().copy(), pro))
        out.append(rec)
    for k in range(n_steps):
        a_big, s_big, p_big, f_big, pro_big = out[0][k]
        a_sub, s_sub, p_sub, f_sub, pro_sub = out[1][k]
        assert np.array_equal(s_big[idx], s_sub) and p_big[idx].tobytes() == p_sub.tobytes(), k
        # the big batch runs the per-env prologue (step_prologue_env_kernel), the small one the per-(env, leg) mapping:
        # same routines per leg and per estimator axis, so every prologue output is bit-identical
        for name, big, small in zip(PROLOGUE_OUTPUTS, pro_big, pro_sub):
            assert big[idx].tobytes() == small.tobytes(), (k, name)
        assert np.all(np.isfinite(a_big))
        scale = np.maximum(1.0, np.abs(f_sub).max(axis=1, keepdims=True))
        assert (np.abs(f_big[idx] - f_sub) / scale).max() < 1e-6, k
        assert np.abs(a_big[idx] - a_sub).max() <= 1e-5 * max(1.0, np.abs(a_sub).max()), k
