"""Experiment: which stance leg has fz_min active at step h-2, and can inputs predict it? (oracle only)"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(REPO, "robot-gym_b200")); sys.path.insert(0, REPO)
import numpy as np
from robot_gym.model.robots.descriptions import GHOST
from robot_gym.util import synthetic
from oracle import convex_mpc as cm
ctrl = GHOST.GetCtrlConstants(); st = synthetic.make_states(4096, GHOST); mp = cm.MpcParams(horizon=10)
rows = []
for i in range(400):
    c = st.planned_contacts[i]
    if c.sum() != 2: continue
    qp = cm.build_qp(mp, st.com_velocity_body[i].astype(np.float64), st.base_rpy[i].astype(np.float64), st.base_rpy_rate[i].astype(np.float64),
                     c, st.foot_positions_base[i].astype(np.float64), [0, 0, ctrl.MPC_BODY_HEIGHT],
                     [st.command[i, 0], st.command[i, 1], 0.0], [0, 0, 0], [0, 0, float(st.command[i, 2])])
    x, info = cm.solve_qp(qp.p_mat, qp.q_vec, qp.c_mat, qp.lb, qp.ub)
    cx = qp.c_mat @ x; tol = 1e-7 * qp.ub.max()
    legs = [l for l in range(4) if c[l]]
    a8 = [bool(cx[8 * 20 + l * 5 + 4] < qp.lb[8 * 20 + l * 5 + 4] + tol) for l in legs]
    # unconstrained-ish proxy: fz at t=7 of each leg at the optimum, and inputs
    f7 = [x[7 * 12 + 3 * l + 2] for l in legs]
    feet = st.foot_positions_base[i].reshape(4, 3)
    rows.append((a8[0], a8[1], feet[legs[0], 0], feet[legs[1], 0], st.base_rpy[i][1], st.base_rpy_rate[i][1], st.com_velocity_body[i][0] - st.command[i][0],
                 st.base_rpy[i][0], st.base_rpy_rate[i][0], f7[0], f7[1], legs[0], legs[1]))
r = np.array(rows, dtype=float)
both, none, one = (r[:, 0] + r[:, 1] == 2).mean(), (r[:, 0] + r[:, 1] == 0).mean(), (r[:, 0] + r[:, 1] == 1).mean()
print("2-stance envs:", len(r), "t=8 fz_min active on both / none / exactly one leg:", both.round(3), none.round(3), one.round(3))
sel = r[:, 0] + r[:, 1] == 1
first = r[sel, 0] == 1          # the FIRST stance leg (front one of the diagonal pair) is the active one
print("exactly-one cases: active leg is the front leg of the pair:", first.mean().round(3))
for name, col in (("pitch", 4), ("pitch rate", 5), ("vx err", 6), ("roll", 7), ("roll rate", 8)):
    v = r[sel, col]
    print(f"   corr(front active, {name}) = {np.corrcoef(first.astype(float), v)[0,1]:+.2f}   sign rule accuracy {max(np.mean((v > 0) == first), np.mean((v < 0) == first)):.2f}")
print("   rule 'smaller fz at t=7 is the active one':", np.mean((r[sel, 9] < r[sel, 10]) == first).round(3))
