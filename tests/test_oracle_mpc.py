"""Pins for the convex-MPC oracle (oracle/convex_mpc.py).

The reference ships no golden vectors for this path (parity unpinned), so the oracle is
pinned by solver-independent certificates and implementation-independent known answers:
KKT optimality, symmetry, swing legs exactly zero, the closed-form discretisation identity, the
alpha -> 0 limit m g / 4, and the frozen oracle goldens under tests/golden/.
"""
import os

import numpy as np
import pytest
import scipy.linalg

from oracle import convex_mpc as cm
from robot_gym.model.robots.descriptions import GHOST
from robot_gym.util import synthetic

SYM_FEET = np.array([[0.2, -0.15, -0.42], [0.2, 0.15, -0.42], [-0.2, -0.15, -0.42], [-0.2, 0.15, -0.42]])


def _solve(params, st, i):
    ctrl = GHOST.GetCtrlConstants()
    return cm.compute_contact_forces(
        params, st.com_velocity_body[i].astype(np.float64), st.base_rpy[i].astype(np.float64),
        st.base_rpy_rate[i].astype(np.float64), st.planned_contacts[i], st.foot_positions_base[i].astype(np.float64),
        [0, 0, ctrl.MPC_BODY_HEIGHT], [float(st.command[i, 0]), float(st.command[i, 1]), 0.0], [0, 0, 0],
        [0, 0, float(st.command[i, 2])], return_info=True)


def test_symmetric_stand_known_answers():
    p = cm.MpcParams()
    f = cm.compute_contact_forces(p, [0, 0, 0], [0, 0, 0], [0, 0, 0], [1, 1, 1, 1], SYM_FEET.ravel(),
                                  [0, 0, 0.42], [0, 0, 0], [0, 0, 0], [0, 0, 0])
    first = f[:12].reshape(4, 3)
    assert np.abs(first[:, :2]).max() < 1e-8                 # no tangential force, any alpha (symmetry)
    assert np.ptp(first[:, 2]) < 1e-8                        # four equal vertical forces
    assert first[0, 2] < 0                                   # returned force = -(GRF): pushes the ground down
    # alpha > 0 reshapes the profile over the horizon: NOT m g / 4 ...
    assert abs(-first[0, 2] - 47.5) > 1.0
    # ... but m g / 4 = 47.5 N is the alpha -> 0 limit
    p0 = cm.MpcParams(alpha=1e-9)
    f0 = cm.compute_contact_forces(p0, [0, 0, 0], [0, 0, 0], [0, 0, 0], [1, 1, 1, 1], SYM_FEET.ravel(),
                                   [0, 0, 0.42], [0, 0, 0], [0, 0, 0], [0, 0, 0])
    assert abs(-f0[2] - 190.0 / 4) < 1e-2


def test_swing_legs_are_exactly_zero_and_bounds_hold():
    p = cm.MpcParams()
    st = synthetic.make_states(12, GHOST)
    for i in range(12):
        f, info = _solve(p, st, i)
        force = -f.reshape(p.horizon, 4, 3)                  # ground reaction forces
        swing = st.planned_contacts[i] == 0
        assert np.all(force[:, swing, :] == 0.0)
        fz = force[:, ~swing, 2]
        assert fz.min() >= p.fz_min - 1e-6 and fz.max() <= p.fz_max + 1e-6
        mu = p.friction_coeffs[0]
        assert np.all(np.abs(force[:, ~swing, 0]) <= mu * fz + 1e-6)
        assert np.all(np.abs(force[:, ~swing, 1]) <= mu * fz + 1e-6)


@pytest.mark.parametrize("weights", [None, (5, 5, 0.2, 0, 0, 10, 0., 0., 1., 1., 1., 0., 0)])
def test_kkt_certificate_on_random_states(weights):
    p = cm.MpcParams() if weights is None else cm.MpcParams(weights=weights)
    st = synthetic.make_states(10, GHOST, seed=5)
    for i in range(10):
        f, info = _solve(p, st, i)
        qp = info["qp"]
        k = cm.kkt_certificate(qp.p_mat, qp.q_vec, qp.c_mat, qp.lb, qp.ub, info["x"])
        scale = max(1.0, np.abs(qp.q_vec).max())
        assert k["stationarity"] < 1e-9 * scale, k
        assert k["primal"] < 1e-9 * p.fz_max and k["dual_sign"] < 1e-9 * scale and k["complementarity"] == 0.0
        assert info["polished"]


def test_closed_form_discretisation_identity():
    """[[A,B],[0,0]] is nilpotent of index 3: expm = I + M dt + M^2 dt^2/2 exactly,
    and A_d^d B_d = B dt + (d + 1/2) dt^2 A B -- the identity the CUDA kernel is built on."""
    rng = np.random.default_rng(3)
    dt = 0.025
    for _ in range(20):
        rpy = rng.uniform(-0.4, 0.4, 3)
        feet = SYM_FEET + rng.uniform(-0.05, 0.05, (4, 3))
        a = cm.calculate_a_mat(rpy)
        rot = cm.rpy_to_rot_zyx(rpy)
        inv_i = rot @ np.linalg.inv(np.diag([0.07335, 0.25068, 0.25447])) @ rot.T
        b = cm.calculate_b_mat(9.8 / 190, inv_i, feet)
        m = np.zeros((25, 25))
        m[:13, :13], m[:13, 13:] = a, b
        assert np.abs(np.linalg.matrix_power(m, 3)).max() == 0.0
        a_exp, b_exp = cm.calculate_exponentials(a, b, dt)
        assert np.abs(a_exp - (np.eye(13) + a * dt + a @ a * dt * dt / 2)).max() < 1e-15
        assert np.abs(b_exp - (b * dt + a @ b * dt * dt / 2)).max() < 1e-15
        ad = np.eye(13)
        for d in range(10):
            assert np.abs(ad @ b_exp - (b * dt + (d + 0.5) * dt * dt * (a @ b))).max() < 1e-14
            ad = a_exp @ ad


def test_constraint_rows_and_bounds_layout():
    c = cm.update_constraints_matrix((0.45, 0.45, 0.45, 0.45), 2, 4)
    assert c.shape == (40, 24)
    np.testing.assert_array_equal(c[:5, :3], [[-1, 0, 0.45], [1, 0, 0.45], [0, -1, 0.45], [0, 1, 0.45], [0, 0, 1]])
    contact = np.array([[1, 0, 1, 1], [1, 0, 1, 1]], dtype=float)
    lb, ub = cm.calculate_constraint_bounds(contact, 1900.0, 19.0, 0.45, 2)
    assert lb[4] == 19.0 and ub[4] == 1900.0 and ub[0] == pytest.approx(1.45 * 1900.0)
    assert np.all(lb[5:10] == 0) and np.all(ub[5:10] == 0)   # swing foot: 0 <= C f <= 0


def test_com_height_estimate_and_explicit_height():
    p = cm.MpcParams()
    args = ([0.1, 0, 0], [0.05, -0.1, 0], [0, 0, 0.1], [1, 0, 0, 1], (SYM_FEET + [0, 0, 0.02]).ravel(),
            [0, 0, 0.42], [0.2, 0.08, 0], [0, 0, 0], [0, 0, 0.1])
    qp = cm.build_qp(p, *args)
    feet_w = (cm.foot_rotation_xyz(np.array(args[1])) @ (SYM_FEET + [0, 0, 0.02]).T).T
    assert qp.com_z == pytest.approx(abs(feet_w[[0, 3], 2].mean()))
    qp2 = cm.build_qp(p, *args, com_position=[0, 0, 0.5])
    assert qp2.com_z == 0.5 and qp2.x0[5] == 0.5


def test_no_stance_legs_gives_zero():
    f = cm.compute_contact_forces(cm.MpcParams(), [0, 0, 0], [0, 0, 0], [0, 0, 0], [0, 0, 0, 0], SYM_FEET.ravel(),
                                  [0, 0, 0.42], [0, 0, 0], [0, 0, 0], [0, 0, 0])
    assert np.all(f == 0)


def test_mirror_symmetry():
    """Reflecting the state across the sagittal plane (y -> -y, swapping left and right legs) must
    reflect the forces: a property of the formulation, independent of any implementation."""
    p = cm.MpcParams()
    rng = np.random.default_rng(11)
    feet = SYM_FEET + rng.uniform(-0.03, 0.03, (4, 3))
    v, w, rpy = rng.uniform(-0.3, 0.3, 3), rng.uniform(-0.3, 0.3, 3), np.array([0.1, -0.05, 0.0])
    f = cm.compute_contact_forces(p, v, rpy, w, [1, 1, 1, 1], feet.ravel(), [0, 0, 0.42], [0.2, 0.05, 0], [0, 0, 0], [0, 0, 0.1])
    sy = np.array([1, -1, 1.0])
    feet_m = (feet * sy)[[1, 0, 3, 2]]
    f_m = cm.compute_contact_forces(p, v * sy, rpy * [-1, 1, -1], w * [-1, 1, -1], [1, 1, 1, 1], feet_m.ravel(),
                                    [0, 0, 0.42], [0.2, -0.05, 0], [0, 0, 0], [0, 0, -0.1])
    a = f[:12].reshape(4, 3)
    b = (f_m[:12].reshape(4, 3) * sy)[[1, 0, 3, 2]]
    assert np.abs(a - b).max() < 1e-7 * max(1.0, np.abs(a).max())


def test_oracle_matches_frozen_goldens(golden_dir):
    g = np.load(os.path.join(golden_dir, "mpc_oracle_golden.npz"))
    for horizon in (10, 5, 20):
        n = int(g[f"mpc_h{horizon}_n"])
        st = synthetic.make_states(n, GHOST, seed=synthetic.SEED + horizon)
        p = cm.MpcParams(horizon=horizon)
        for i in range(0, n, max(1, n // 6)):
            f, _ = _solve(p, st, i)
            ref = g[f"mpc_h{horizon}_forces"][i]
            assert np.abs(f - ref).max() <= 1e-8 * max(1.0, np.abs(ref).max())


def test_structured_prototype_matches_dense_oracle():
    """The Kronecker/Woodbury algebra the CUDA kernel uses (tools/prototype_structured_ipm.py)
    against the dense reference-style pipeline."""
    import importlib.util, sys
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "prototype_structured_ipm.py")
    spec = importlib.util.spec_from_file_location("prototype_structured_ipm", path)
    proto = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(proto)
    p = cm.MpcParams()
    st = synthetic.make_states(4, GHOST, seed=9)
    ctrl = GHOST.GetCtrlConstants()
    for i in range(4):
        args = (st.com_velocity_body[i].astype(np.float64), st.base_rpy[i].astype(np.float64),
                st.base_rpy_rate[i].astype(np.float64), st.planned_contacts[i], st.foot_positions_base[i].astype(np.float64),
                [0, 0, ctrl.MPC_BODY_HEIGHT], [float(st.command[i, 0]), float(st.command[i, 1]), 0.0], [0, 0, 0],
                [0, 0, float(st.command[i, 2])])
        ref = cm.compute_contact_forces(p, *args)
        sol, _ = proto.solve_structured(p, proto.build_structured(p, *args), tol=1e-12, max_iter=40)
        assert np.abs(-sol.reshape(-1) - ref).max() < 1e-4 * max(1.0, np.abs(ref).max())
