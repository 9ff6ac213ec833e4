"""Experiment: which rows are active at the optimum beyond the built-in guess (fz_min at the last step)? (oracle only)"""
import os, sys, collections
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(REPO, "robot-gym_b200")); sys.path.insert(0, REPO)
import numpy as np
from robot_gym.model.robots.descriptions import GHOST
from robot_gym.util import synthetic
from oracle import convex_mpc as cm
ctrl = GHOST.GetCtrlConstants(); st = synthetic.make_states(4096, GHOST); mp = cm.MpcParams(horizon=10)
n = 300; extra = collections.Counter(); nextra = []; missing = 0
fz8 = []
for i in range(n):
    qp = cm.build_qp(mp, st.com_velocity_body[i].astype(np.float64), st.base_rpy[i].astype(np.float64), st.base_rpy_rate[i].astype(np.float64),
                     st.planned_contacts[i], st.foot_positions_base[i].astype(np.float64), [0, 0, ctrl.MPC_BODY_HEIGHT],
                     [st.command[i, 0], st.command[i, 1], 0.0], [0, 0, 0], [0, 0, float(st.command[i, 2])])
    x, info = cm.solve_qp(qp.p_mat, qp.q_vec, qp.c_mat, qp.lb, qp.ub)
    cx = qp.c_mat @ x; tol = 1e-7 * qp.ub.max(); swing = qp.ub == qp.lb
    lo_act = (cx < qp.lb + tol) & ~swing; hi_act = (cx > qp.ub - tol) & ~swing
    cnt = 0
    for r in np.flatnonzero(lo_act | hi_act):
        t, leg, row = r // 20, (r % 20) // 5, r % 5
        if t == 9 and row == 4 and lo_act[r]: continue
        extra[(t, "fz_min" if (row == 4 and lo_act[r]) else "fz_max" if row == 4 else "cone")] += 1; cnt += 1
    stance = [l for l in range(4) if st.planned_contacts[i][l]]
    for l in stance:
        if not lo_act[9 * 20 + l * 5 + 4]: missing += 1
    nextra.append(cnt)
nextra = np.array(nextra)
print("envs with NO extra active rows (1 round):", np.mean(nextra == 0), " guess rows that are not active:", missing)
print("extra active rows per env by (step, kind):")
for k, v in sorted(extra.items()): print("  ", k, round(v / n, 3))
print("hist of #extra:", np.bincount(nextra)[:12])
