"""Experiment: how large is the optimal active set on the bench batch, and how many rounds
does a cold-start primal-dual active-set iteration (no interior point first) need?

Uses the numpy oracle (test infrastructure) only; informs the CUDA solver's start strategy.
"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(REPO, "robot-gym_b200")); sys.path.insert(0, REPO)
import numpy as np
from robot_gym.model.robots.descriptions import GHOST, with_gait
from robot_gym.util import synthetic
from oracle import convex_mpc as cm


def pdas(pm, qv, cmx, lo, hi, side, max_rounds=20):
    feas_tol = 1e-9 * max(1.0, float(np.abs(hi).max()))
    for rnd in range(1, max_rounds + 1):
        rows = np.flatnonzero(side)
        b_act = np.where(side[rows] > 0, hi[rows], lo[rows])
        xp, yp = cm._solve_equality_qp(pm, qv, cmx[rows], b_act)
        cxp = cmx @ xp
        vh, vl = cxp - hi, lo - cxp
        vh[rows] = 0; vl[rows] = 0
        changed = False
        if max(vh.max(), vl.max()) > feas_tol:
            side[vh > feas_tol] = 1; side[vl > feas_tol] = -1; changed = True
        wrong = (side[rows] * yp) < -1e-12 * max(1.0, float(np.abs(yp).max()) if len(yp) else 1.0)
        if np.any(wrong):
            side[rows[wrong]] = 0; changed = True
        if not changed:
            return rnd, xp, side
    return -1, xp, side


def main(n=300, gait=None, horizon=10):
    desc = GHOST if gait is None else with_gait(GHOST, gait)
    ctrl = desc.GetCtrlConstants()
    st = synthetic.make_states(4096, desc)
    mp = cm.MpcParams(horizon=horizon)
    rounds, nact, kinds = [], [], np.zeros(5)
    for i in range(n):
        qp = cm.build_qp(mp, st.com_velocity_body[i].astype(np.float64), st.base_rpy[i].astype(np.float64),
                         st.base_rpy_rate[i].astype(np.float64), st.planned_contacts[i],
                         st.foot_positions_base[i].astype(np.float64), [0, 0, ctrl.MPC_BODY_HEIGHT],
                         [st.command[i, 0], st.command[i, 1], 0.0], [0, 0, 0], [0, 0, float(st.command[i, 2])])
        nblk = qp.p_mat.shape[0] // 3
        free = np.array([not np.all(qp.ub[5*b:5*b+5] == qp.lb[5*b:5*b+5]) for b in range(nblk)])
        fidx = np.flatnonzero(np.repeat(free, 3)); ridx = np.flatnonzero(np.repeat(free, 5))
        if len(fidx) == 0:
            continue
        pm = qp.p_mat[np.ix_(fidx, fidx)]; qv = qp.q_vec[fidx]; cmx = qp.c_mat[np.ix_(ridx, fidx)]
        lo, hi = qp.lb[ridx], qp.ub[ridx]
        r, xp, side = pdas(pm, qv, cmx, lo, hi, np.zeros(len(hi), dtype=np.int64))
        xref, info = cm.solve_qp(qp.p_mat, qp.q_vec, qp.c_mat, qp.lb, qp.ub)
        err = np.abs(xp - xref[fidx]).max() / max(1.0, np.abs(xref).max())
        rounds.append(r if err < 1e-6 else -2)
        nact.append(int(np.count_nonzero(side)))
        for k in range(5):
            kinds[k] += np.count_nonzero(side[k::5])
    rounds = np.array(rounds); nact = np.array(nact)
    print(f"gait={gait or 'trot'} h={horizon} n={len(rounds)}: cold PDAS rounds hist",
          dict(zip(*np.unique(rounds, return_counts=True))), "mean active", nact.mean(),
          "frac empty", np.mean(nact == 0), "active rows by pyramid row", kinds)


if __name__ == "__main__":
    main()
    main(n=150, gait="pace")
    main(n=150, gait="bound")
