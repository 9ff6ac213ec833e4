"""Oracle: convex-MPC stance QP, dense float64 restatement.  TEST INFRASTRUCTURE ONLY.

Restates ``mpc_osqp.ConvexMpc`` of motion_imitation==0.0.5 (third-party, absent
from /root/reference; reached from
robot_gym/controllers/mpc/mpc_controller.py:47-56 through
``TorqueStanceLegController.get_action``).  PARITY UNPINNED: the reference ships
no golden vectors for this path, see oracle/__init__.py.

The build below follows the dense pipeline of ``ConvexMpc::ComputeContactForces``
step by step (A/B matrices, matrix exponential of [[A,B],[0,0]]*dt, stacked
A_qp/B_qp, P = 2(B_qp^T L B_qp + alpha I), q = 2 B_qp^T L (A_qp x0 - x_ref),
5 friction-pyramid rows per foot per step) on purpose: it shares no code and no
algebraic shortcut with the CUDA kernels (which use the closed-form
discretisation and a Kronecker/Woodbury structure), so agreement between the
two is evidence for both.

State order (kStateDim = 13):
    x = [roll, pitch, yaw,  px, py, pz,  wx, wy, wz,  vx, vy, vz,  -g]
"""
from __future__ import annotations

import dataclasses
import math
from typing import Optional, Sequence

import numpy as np
import scipy.linalg

K_STATE_DIM = 13
K_CONSTRAINT_DIM = 5
K3 = 3


@dataclasses.dataclass
class MpcParams:
    """Constructor arguments of ``ConvexMpc`` plus the constants compiled into mpc_osqp.

    mass / inertia: robot_gym/model/robots/ghost/ctrl_constants.py:8-9 (passed at
    robot_gym/controllers/mpc/mpc_controller.py:54-55).  Everything else is a
    recalled motion_imitation default (confidence [M]) and is
    therefore a *runtime* parameter here and in the CUDA library.
    """
    mass: float = 190.0 / 9.8
    inertia: Sequence[float] = (0.07335, 0, 0, 0, 0.25068, 0, 0, 0, 0.25447)
    num_legs: int = 4
    horizon: int = 10                       # _PLANNING_HORIZON_STEPS
    dt: float = 0.025                       # _PLANNING_TIMESTEP
    weights: Sequence[float] = (5, 5, 0.2, 0, 0, 10, 0.5, 0.5, 0.2, 0.2, 0.2, 0.1, 0)  # _MPC_WEIGHTS
    alpha: float = 1e-5
    friction_coeffs: Sequence[float] = (0.45, 0.45, 0.45, 0.45)
    gravity: float = 9.8                    # kGravity
    fz_max_scale: float = 10.0              # kMaxScale
    fz_min_scale: float = 0.1               # kMinScale

    @property
    def fz_max(self) -> float:
        return self.mass * self.gravity * self.fz_max_scale

    @property
    def fz_min(self) -> float:
        return self.mass * self.gravity * self.fz_min_scale


# --------------------------------------------------------------------------- rotations
def _rot_x(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=np.float64)


def _rot_y(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float64)


def _rot_z(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float64)


def rpy_to_rot_zyx(rpy):
    """``ConvertRpyToRot``: q = yaw * pitch * roll  ->  R = Rz Ry Rx (body -> world)."""
    return _rot_z(rpy[2]) @ _rot_y(rpy[1]) @ _rot_x(rpy[0])


def foot_rotation_xyz(rpy):
    """Rotation used for the foot lever arms in ``ComputeContactForces``:
    AngleAxis(roll,X) * AngleAxis(pitch,Y) * AngleAxis(yaw,Z)  ->  R = Rx Ry Rz.

    This multiplication order differs from ``ConvertRpyToRot`` (recalled quirk of
    mpc_osqp.cc, confidence [M]); it is restated as is.  With the yaw the python
    wrapper forces to zero the two differ only at second order in roll*pitch.
    """
    return _rot_x(rpy[0]) @ _rot_y(rpy[1]) @ _rot_z(rpy[2])


def skew(v):
    """``ConvertToSkewSymmetric``."""
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], dtype=np.float64)


# --------------------------------------------------------------------------- dynamics
def calculate_a_mat(rpy):
    """``CalculateAMat``: continuous-time A (13x13)."""
    cy, sy = math.cos(rpy[2]), math.sin(rpy[2])
    cp, tp = math.cos(rpy[1]), math.tan(rpy[1])
    a = np.zeros((K_STATE_DIM, K_STATE_DIM))
    a[0:3, 6:9] = np.array([[cy / cp, sy / cp, 0.0],
                            [-sy, cy, 0.0],
                            [cy * tp, sy * tp, 1.0]])
    a[3, 9] = a[4, 10] = a[5, 11] = 1.0
    a[11, 12] = 1.0
    return a


def calculate_b_mat(inv_mass, inv_inertia_world, foot_positions_world):
    """``CalculateBMat``: continuous-time B (13 x 3k)."""
    k = foot_positions_world.shape[0]
    b = np.zeros((K_STATE_DIM, K3 * k))
    for i in range(k):
        b[6:9, 3 * i:3 * i + 3] = inv_inertia_world @ skew(foot_positions_world[i])
        b[9, 3 * i] = b[10, 3 * i + 1] = b[11, 3 * i + 2] = inv_mass
    return b


def calculate_exponentials(a_mat, b_mat, dt):
    """``CalculateExponentials``: expm([[A,B],[0,0]] dt) -> (A_exp, B_exp), scipy Pade."""
    s, m = K_STATE_DIM, b_mat.shape[1]
    ab = np.zeros((s + m, s + m))
    ab[:s, :s] = a_mat * dt
    ab[:s, s:] = b_mat * dt
    e = scipy.linalg.expm(ab)
    return e[:s, :s], e[:s, s:]


def calculate_qp_mats(a_exp, b_exp, weights, alpha, horizon):
    """``CalculateQpMats``: stacked A_qp (13h x 13), B_qp (13h x 3kh) and P = 2(B^T L B + alpha I)."""
    s, m = K_STATE_DIM, b_exp.shape[1]
    a_qp = np.zeros((s * horizon, s))
    a_pow = np.eye(s)
    anb = []
    for i in range(horizon):
        anb.append(a_pow @ b_exp)          # A^i B
        a_pow = a_exp @ a_pow
        a_qp[i * s:(i + 1) * s] = a_pow    # A^(i+1)
    b_qp = np.zeros((s * horizon, m * horizon))
    for i in range(horizon):
        for j in range(i + 1):
            b_qp[i * s:(i + 1) * s, j * m:(j + 1) * m] = anb[i - j]
    big_l = np.tile(np.asarray(weights, dtype=np.float64), horizon)
    p_mat = 2.0 * (b_qp.T @ (big_l[:, None] * b_qp) + alpha * np.eye(m * horizon))
    return a_qp, b_qp, big_l, p_mat


def update_constraints_matrix(friction_coeffs, horizon, num_legs):
    """``UpdateConstraintsMatrix``: 5 rows per foot per step.

    Recalled quirk (confidence [M]): row r of EVERY foot's pyramid uses
    ``friction_coeff[r]`` -- the 4-vector is indexed by pyramid row, not by leg.  The only
    call site passes four equal values (0.45), for which both readings coincide.
    """
    mu = friction_coeffs
    blk = np.array([[-1, 0, mu[0]], [1, 0, mu[1]], [0, -1, mu[2]], [0, 1, mu[3]], [0, 0, 1]], dtype=np.float64)
    nb = horizon * num_legs
    c = np.zeros((K_CONSTRAINT_DIM * nb, K3 * nb))
    for i in range(nb):
        c[5 * i:5 * i + 5, 3 * i:3 * i + 3] = blk
    return c


def calculate_constraint_bounds(contact_state, fz_max, fz_min, friction_coeff, horizon):
    """``CalculateConstraintBounds``: contact_state is (horizon x num_legs) of 0/1."""
    k = contact_state.shape[1]
    lb = np.zeros(K_CONSTRAINT_DIM * horizon * k)
    ub = np.zeros_like(lb)
    for i in range(horizon):
        for j in range(k):
            row = (i * k + j) * K_CONSTRAINT_DIM
            c = float(contact_state[i, j])
            lb[row + 4] = fz_min * c
            ub[row:row + 4] = (friction_coeff + 1.0) * fz_max * c
            ub[row + 4] = fz_max * c
    return lb, ub


def estimate_com_height_simple(foot_positions_world, contacts, fallback):
    """``EstimateCoMHeightSimple``: |mean z of the contact feet| (used because the python
    wrapper passes ``com_position=[0]``).  The reference DCHECKs >0 contacts; with none
    every force is pinned to zero by the bounds, so the height is irrelevant: ``fallback``."""
    n = int(np.sum(contacts != 0))
    if n == 0:
        return float(fallback)
    return abs(float(np.sum(foot_positions_world[contacts != 0, 2])) / n)


@dataclasses.dataclass
class QpProblem:
    p_mat: np.ndarray
    q_vec: np.ndarray
    c_mat: np.ndarray
    lb: np.ndarray
    ub: np.ndarray
    x0: np.ndarray
    com_z: float


def build_qp(params: MpcParams, com_velocity, rpy, angular_velocity, contacts, foot_positions_base,
             desired_com_position, desired_com_velocity, desired_rpy, desired_angular_velocity,
             com_position: Optional[Sequence[float]] = None) -> QpProblem:
    """Dense QP of ``ConvexMpc::ComputeContactForces`` (everything before the OSQP call)."""
    h, k, dt = params.horizon, params.num_legs, params.dt
    rpy = np.asarray(rpy, dtype=np.float64)
    contacts = np.asarray(contacts).astype(np.int64)
    feet_base = np.asarray(foot_positions_base, dtype=np.float64).reshape(k, 3)
    feet_world = (foot_rotation_xyz(rpy) @ feet_base.T).T
    if com_position is not None and len(com_position) == 3:
        com_z = float(com_position[2])
    else:
        com_z = estimate_com_height_simple(feet_world, contacts, desired_com_position[2])

    g = params.gravity
    x0 = np.array([rpy[0], rpy[1], rpy[2], 0.0, 0.0, com_z,
                   angular_velocity[0], angular_velocity[1], angular_velocity[2],
                   com_velocity[0], com_velocity[1], com_velocity[2], -g], dtype=np.float64)
    x_ref = np.zeros(K_STATE_DIM * h)
    for i in range(h):
        o = i * K_STATE_DIM
        x_ref[o + 0] = desired_rpy[0]
        x_ref[o + 1] = desired_rpy[1]
        x_ref[o + 2] = rpy[2] + dt * (i + 1) * desired_angular_velocity[2]
        x_ref[o + 3] = dt * (i + 1) * desired_com_velocity[0]
        x_ref[o + 4] = dt * (i + 1) * desired_com_velocity[1]
        x_ref[o + 5] = desired_com_position[2]
        x_ref[o + 6] = 0.0
        x_ref[o + 7] = 0.0
        x_ref[o + 8] = desired_angular_velocity[2]
        x_ref[o + 9] = desired_com_velocity[0]
        x_ref[o + 10] = desired_com_velocity[1]
        x_ref[o + 11] = 0.0
        x_ref[o + 12] = -g

    a_mat = calculate_a_mat(rpy)
    rot = rpy_to_rot_zyx(rpy)
    inv_inertia = np.linalg.inv(np.asarray(params.inertia, dtype=np.float64).reshape(3, 3))
    inv_inertia_world = rot @ inv_inertia @ rot.T
    b_mat = calculate_b_mat(1.0 / params.mass, inv_inertia_world, feet_world)
    a_exp, b_exp = calculate_exponentials(a_mat, b_mat, dt)
    a_qp, b_qp, big_l, p_mat = calculate_qp_mats(a_exp, b_exp, params.weights, params.alpha, h)
    state_diff = a_qp @ x0 - x_ref
    q_vec = 2.0 * b_qp.T @ (big_l * state_diff)

    contact_state = np.tile((contacts != 0).astype(np.float64)[None, :], (h, 1))
    lb, ub = calculate_constraint_bounds(contact_state, params.fz_max, params.fz_min,
                                         params.friction_coeffs[0], h)
    c_mat = update_constraints_matrix(params.friction_coeffs, h, k)
    return QpProblem(p_mat, q_vec, c_mat, lb, ub, x0, com_z)


# --------------------------------------------------------------------------- QP solve
def kkt_certificate(p_mat, q_vec, c_mat, lb, ub, x, y=None, act_tol=1e-7):
    """Solver-independent optimality certificate for  min 1/2 x'Px + q'x  s.t. lb <= Cx <= ub.

    If ``y`` is None the multipliers are recovered by least squares on the active rows.
    Returns dict(stationarity, primal, dual_sign, complementarity) -- all should be ~0.
    """
    cx = c_mat @ x
    scale = max(1.0, float(np.max(np.abs(ub))))
    at_lo = cx <= lb + act_tol * scale
    at_hi = cx >= ub - act_tol * scale
    primal = float(max(np.max(lb - cx), np.max(cx - ub), 0.0))
    grad = p_mat @ x + q_vec
    if y is None:
        act = at_lo | at_hi
        y = np.zeros_like(cx)
        if np.any(act):
            sol, *_ = np.linalg.lstsq(c_mat[act].T, -grad, rcond=None)
            y[act] = sol
    stat = float(np.max(np.abs(grad + c_mat.T @ y)))
    eq = at_lo & at_hi
    bad_lo = np.where(at_lo & ~eq, np.maximum(y, 0.0), 0.0)      # at lower bound y must be <= 0
    bad_hi = np.where(at_hi & ~eq, np.maximum(-y, 0.0), 0.0)     # at upper bound y must be >= 0
    free = ~(at_lo | at_hi)
    comp = float(np.max(np.abs(y[free]))) if np.any(free) else 0.0
    return dict(stationarity=stat, primal=primal,
                dual_sign=float(max(bad_lo.max(), bad_hi.max())), complementarity=comp, y=y)


def _solve_equality_qp(p_mat, q_vec, c_act, b_act):
    """min 1/2 x'Px + q'x  s.t. c_act x = b_act  (null-space method, float64)."""
    n = p_mat.shape[0]
    if c_act.shape[0] == 0:
        return np.linalg.solve(p_mat, -q_vec), np.zeros(0)
    qmat, rmat, piv = scipy.linalg.qr(c_act.T, mode="full", pivoting=True)
    rank = int(np.sum(np.abs(np.diag(rmat)) > 1e-10 * max(1.0, abs(rmat[0, 0]))))
    y_basis, z_basis = qmat[:, :rank], qmat[:, rank:]
    # particular solution: c_act x = b_act with x in range(y_basis)
    rp = rmat[:rank, :]
    b_piv = b_act[piv]
    xy, *_ = np.linalg.lstsq(rp.T, b_piv, rcond=None)
    x_part = y_basis @ xy
    if z_basis.shape[1] > 0:
        hz = z_basis.T @ p_mat @ z_basis
        xz = np.linalg.solve(hz, -z_basis.T @ (p_mat @ x_part + q_vec))
        x = x_part + z_basis @ xz
    else:
        x = x_part
    lam, *_ = np.linalg.lstsq(c_act.T, -(p_mat @ x + q_vec), rcond=None)
    return x, lam


def solve_qp(p_mat, q_vec, c_mat, lb, ub, tol=1e-10, max_iter=60, polish=True):
    """Dense Mehrotra predictor-corrector interior point + active-set polish (float64).

    Stands in for the reference's ``OSQP(eps=1e-3, polish=True)``: because P is positive
    definite (alpha > 0) the optimum is unique, so any method that satisfies the KKT
    certificate has found the same point the reference's polished solve converges to.
    Rows with lb == ub (swing feet: 0 <= Cx <= 0 pins f = 0) are eliminated first.
    """
    n = p_mat.shape[0]
    # --- eliminate variables pinned by equality blocks (swing feet)
    nblk = n // 3
    free_blk = np.array([not np.all(ub[5 * b:5 * b + 5] == lb[5 * b:5 * b + 5]) for b in range(nblk)])
    fidx = np.flatnonzero(np.repeat(free_blk, 3))
    ridx = np.flatnonzero(np.repeat(free_blk, 5))
    x_full = np.zeros(n)
    info = dict(iters=0, n_free=len(fidx))
    if len(fidx) == 0:
        return x_full, info
    pm = p_mat[np.ix_(fidx, fidx)]
    qv = q_vec[fidx]
    cm = c_mat[np.ix_(ridx, fidx)]
    lo, hi = lb[ridx], ub[ridx]
    gm = np.vstack([cm, -cm])
    hv = np.concatenate([hi, -lo])
    m = gm.shape[0]

    # strictly feasible start: fz midway in log scale, no tangential force
    x = np.zeros(len(fidx))
    fz0 = math.sqrt(max(lo[4], 1e-3) * hi[4])
    x[2::3] = fz0
    s = hv - gm @ x
    assert np.all(s > 0)
    lam = np.maximum(1.0, np.abs(qv).max()) / s * 1e-1 + 0.0
    it = 0
    for it in range(1, max_iter + 1):
        r_d = pm @ x + qv + gm.T @ lam
        r_p = gm @ x + s - hv
        mu = float(s @ lam) / m
        scale = max(1.0, float(np.abs(qv).max()))
        if max(np.abs(r_d).max() / scale, np.abs(r_p).max() / max(1.0, np.abs(hv).max()), mu / scale) < tol:
            break
        d = lam / s
        phi = pm + gm.T @ (d[:, None] * gm)
        try:
            cho = scipy.linalg.cho_factor(phi)
        except np.linalg.LinAlgError:
            break       # deep iterate lost positive definiteness to rounding: the polish takes over from here

        def newton(r_c):
            rhs = -r_d - gm.T @ ((-r_c + lam * r_p) / s)
            dx = scipy.linalg.cho_solve(cho, rhs)
            ds = -r_p - gm @ dx
            dl = (-r_c - lam * ds) / s
            return dx, ds, dl

        def max_step(v, dv):
            neg = dv < 0
            return min(1.0, float(np.min(-v[neg] / dv[neg]))) if np.any(neg) else 1.0

        dx_a, ds_a, dl_a = newton(s * lam)
        a_aff = min(max_step(s, ds_a), max_step(lam, dl_a))
        mu_aff = float((s + a_aff * ds_a) @ (lam + a_aff * dl_a)) / m
        sigma = (mu_aff / mu) ** 3
        dx, ds, dl = newton(s * lam + ds_a * dl_a - sigma * mu)
        a = min(1.0, 0.99 * min(max_step(s, ds), max_step(lam, dl)))
        x, s, lam = x + a * dx, s + a * ds, lam + a * dl
    info["iters"] = it
    info["mu"] = float(s @ lam) / m

    if polish:
        # Primal-dual active-set refinement started from the interior-point guess: solve the
        # equality-constrained QP on the guessed active rows exactly, add violated rows, drop
        # rows whose multiplier has the wrong sign, repeat.  From an IPM point at 1e-10 this
        # settles in one or two rounds and yields multipliers for the KKT certificate.
        cx = cm @ x
        span = np.maximum(1.0, np.abs(hi))
        y_hi, y_lo = lam[:len(hi)], lam[len(hi):]
        side = np.zeros(len(hi), dtype=np.int64)            # +1 at upper bound, -1 at lower, 0 free
        side[(y_hi > 1e-7 * np.maximum(1.0, y_hi.max())) & (hi - cx < 1e-5 * span)] = 1
        side[(y_lo > 1e-7 * np.maximum(1.0, y_lo.max())) & (cx - lo < 1e-5 * span)] = -1
        info["polished"] = False
        feas_tol = 1e-11 * max(1.0, float(np.abs(hi).max()))
        for _ in range(20):
            rows = np.flatnonzero(side)
            b_act = np.where(side[rows] > 0, hi[rows], lo[rows])
            xp, yp = _solve_equality_qp(pm, qv, cm[rows], b_act)
            cxp = cm @ xp
            viol_hi, viol_lo = cxp - hi, lo - cxp
            viol_hi[rows] = 0.0
            viol_lo[rows] = 0.0
            changed = False
            if max(viol_hi.max(), viol_lo.max()) > feas_tol:
                add_hi = viol_hi > feas_tol
                add_lo = viol_lo > feas_tol
                side[add_hi] = 1
                side[add_lo] = -1
                changed = True
            wrong = (side[rows] * yp) < -1e-12 * max(1.0, float(np.abs(yp).max()) if len(yp) else 1.0)
            if np.any(wrong):
                side[rows[wrong]] = 0
                changed = True
            if not changed:
                info["polished"] = True
                x = xp
                y_full = np.zeros(len(hi))
                y_full[rows] = yp
                info["y_free_rows"] = (ridx, y_full)
                break
    x_full[fidx] = x
    return x_full, info


def compute_contact_forces(params: MpcParams, com_velocity, rpy, angular_velocity, contacts,
                           foot_positions_base, desired_com_position, desired_com_velocity,
                           desired_rpy, desired_angular_velocity, com_position=None,
                           return_info=False):
    """``ConvexMpc.compute_contact_forces``: returns the NEGATED QP solution (3*k*h values; the
    stance controller uses the first 3k).  mpc_osqp returns ``-solution``: the force the
    leg applies to the ground, which ``MapContactForceToJointTorques`` turns into torques
    (robot_gym/controllers/mpc/kinematics.py:40-53)."""
    qp = build_qp(params, com_velocity, rpy, angular_velocity, contacts, foot_positions_base,
                  desired_com_position, desired_com_velocity, desired_rpy, desired_angular_velocity,
                  com_position)
    x, info = solve_qp(qp.p_mat, qp.q_vec, qp.c_mat, qp.lb, qp.ub)
    if return_info:
        info["qp"] = qp
        info["x"] = x
        return -x, info
    return -x
