"""Experiment: at which horizon steps do the active rows of the optimal solution sit? (oracle only)"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(REPO, "robot-gym_b200")); sys.path.insert(0, REPO)
import numpy as np
from robot_gym.model.robots.descriptions import GHOST, with_gait
from robot_gym.util import synthetic
from oracle import convex_mpc as cm

def main(n=200, gait=None, h=10):
    desc = GHOST if gait is None else with_gait(GHOST, gait)
    ctrl = desc.GetCtrlConstants()
    st = synthetic.make_states(4096, desc)
    mp = cm.MpcParams(horizon=h)
    hist = np.zeros(h); first = np.zeros(h + 1); viol_first = np.zeros(h + 1)
    for i in range(n):
        args = (mp, st.com_velocity_body[i].astype(np.float64), st.base_rpy[i].astype(np.float64),
                st.base_rpy_rate[i].astype(np.float64), st.planned_contacts[i], st.foot_positions_base[i].astype(np.float64),
                [0, 0, ctrl.MPC_BODY_HEIGHT], [st.command[i, 0], st.command[i, 1], 0.0], [0, 0, 0], [0, 0, float(st.command[i, 2])])
        qp = cm.build_qp(*args)
        x, info = cm.solve_qp(qp.p_mat, qp.q_vec, qp.c_mat, qp.lb, qp.ub)
        cx = qp.c_mat @ x
        tol = 1e-7 * qp.ub.max()
        swing = qp.ub == qp.lb
        act = ((cx > qp.ub - tol) | (cx < qp.lb + tol)) & ~swing
        t_of_row = np.arange(len(cx)) // 20
        for t in range(h): hist[t] += act[t_of_row == t].sum()
        ts = t_of_row[act]
        first[ts.min() if len(ts) else h] += 1
        free = np.repeat(~np.all((qp.ub == qp.lb).reshape(-1, 5), axis=1), 3)
        if not free.any(): continue
        xu = np.zeros_like(x); xu[free] = np.linalg.solve(qp.p_mat[np.ix_(free, free)], -qp.q_vec[free])
        cu = qp.c_mat @ xu
        viol = ((cu > qp.ub + tol) | (cu < qp.lb - tol)) & ~swing
        tv = t_of_row[viol]
        viol_first[tv.min() if len(tv) else h] += 1
    print(f"{gait or 'trot'} h={h}: active rows per step", (hist / n).round(2), "\n   first active step hist", first.astype(int), "\n   first violated step of the unconstrained minimiser", viol_first.astype(int))

if __name__ == "__main__":
    main(); main(100, "pace"); main(100, "walk")
