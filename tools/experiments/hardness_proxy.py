"""Experiment: can a cheap function of the INPUTS predict which envs the cold start gives up on
(many friction-cone rows violated by the unconstrained minimiser)?  (oracle only)"""
import os, sys, pickle
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(REPO, "robot-gym_b200")); sys.path.insert(0, REPO)
import numpy as np
from robot_gym.model.robots.descriptions import GHOST
from robot_gym.util import synthetic
from oracle import convex_mpc as cm

def auc(score, label):
    order = np.argsort(score); r = np.empty(len(score)); r[order] = np.arange(1, len(score) + 1)
    n1 = label.sum(); n0 = len(label) - n1
    return (r[label].sum() - n1 * (n1 + 1) / 2) / (n1 * n0)

n = 1200
ctrl = GHOST.GetCtrlConstants()
st = synthetic.make_states(4096, GHOST)
mp = cm.MpcParams(horizon=10)
cache = os.path.join(REPO, "gpurun_out", "hardness_labels.npy")
if os.path.exists(cache):
    ncone = np.load(cache)
else:
    ncone = np.zeros(n)
    for i in range(n):
        qp = cm.build_qp(mp, st.com_velocity_body[i].astype(np.float64), st.base_rpy[i].astype(np.float64),
                         st.base_rpy_rate[i].astype(np.float64), st.planned_contacts[i], st.foot_positions_base[i].astype(np.float64),
                         [0, 0, ctrl.MPC_BODY_HEIGHT], [st.command[i, 0], st.command[i, 1], 0.0], [0, 0, 0], [0, 0, float(st.command[i, 2])])
        free = np.repeat(~np.all((qp.ub == qp.lb).reshape(-1, 5), axis=1), 3)
        if not free.any(): continue
        x = np.zeros(len(free)); x[free] = np.linalg.solve(qp.p_mat[np.ix_(free, free)], -qp.q_vec[free])
        cx = qp.c_mat @ x; tol = 1e-9 * qp.ub.max()
        viol = ((cx > qp.ub + tol) | (cx < qp.lb - tol)) & (qp.ub != qp.lb)
        ncone[i] = viol[np.arange(len(cx)) % 5 < 4].sum()
    np.save(cache, ncone)
hard = ncone > 16
print("hard fraction", hard.mean())
s = st.slice(0, n)
feat = {"|roll|": np.abs(s.base_rpy[:, 0]), "|pitch|": np.abs(s.base_rpy[:, 1]), "|vx-cmd|": np.abs(s.com_velocity_body[:, 0] - s.command[:, 0]),
        "|vy-cmd|": np.abs(s.com_velocity_body[:, 1] - s.command[:, 1]), "|vz|": np.abs(s.com_velocity_body[:, 2]),
        "|wx|": np.abs(s.base_rpy_rate[:, 0]), "|wy|": np.abs(s.base_rpy_rate[:, 1]), "|wz-cmd|": np.abs(s.base_rpy_rate[:, 2] - s.command[:, 2]),
        "n_stance": s.planned_contacts.sum(axis=1).astype(float)}
for k, v in feat.items(): print(f"  AUC {k:10s} {auc(v, hard):.3f}")
X = np.column_stack(list(feat.values()) + [np.ones(n)])
w, *_ = np.linalg.lstsq(X, hard.astype(float), rcond=None)
sc = X @ w
print("  AUC linear combo", round(auc(sc, hard), 3), "weights", dict(zip(list(feat) + ["1"], w.round(3))))
top = np.argsort(-sc)[: int(0.15 * n)]
print("  hard envs captured in the top 15 % by score:", hard[top].sum(), "of", hard.sum())
