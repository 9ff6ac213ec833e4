"""Experiment: success of cold-start PDAS vs number of rows violated by the unconstrained solution."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(REPO, "robot-gym_b200")); sys.path.insert(0, REPO)
import numpy as np
from robot_gym.model.robots.descriptions import GHOST, with_gait
from robot_gym.util import synthetic
from oracle import convex_mpc as cm
from tools.experiments.active_set_stats import pdas

def main(n, gait=None, horizon=10):
    desc = GHOST if gait is None else with_gait(GHOST, gait)
    ctrl = desc.GetCtrlConstants()
    st = synthetic.make_states(4096, desc)
    mp = cm.MpcParams(horizon=horizon)
    out = []
    for i in range(n):
        qp = cm.build_qp(mp, st.com_velocity_body[i].astype(np.float64), st.base_rpy[i].astype(np.float64),
                         st.base_rpy_rate[i].astype(np.float64), st.planned_contacts[i],
                         st.foot_positions_base[i].astype(np.float64), [0, 0, ctrl.MPC_BODY_HEIGHT],
                         [st.command[i, 0], st.command[i, 1], 0.0], [0, 0, 0], [0, 0, float(st.command[i, 2])])
        nblk = qp.p_mat.shape[0] // 3
        free = np.array([not np.all(qp.ub[5*b:5*b+5] == qp.lb[5*b:5*b+5]) for b in range(nblk)])
        fidx = np.flatnonzero(np.repeat(free, 3)); ridx = np.flatnonzero(np.repeat(free, 5))
        if len(fidx) == 0: continue
        pm = qp.p_mat[np.ix_(fidx, fidx)]; qv = qp.q_vec[fidx]; cmx = qp.c_mat[np.ix_(ridx, fidx)]
        lo, hi = qp.lb[ridx], qp.ub[ridx]
        x0 = np.linalg.solve(pm, -qv); cx = cmx @ x0
        nviol = int(np.count_nonzero((cx > hi + 1e-9 * hi.max()) | (cx < lo - 1e-9 * hi.max())))
        r, xp, side = pdas(pm, qv, cmx, lo, hi, np.zeros(len(hi), dtype=np.int64), max_rounds=12)
        out.append((nviol, r, int(np.count_nonzero(side))))
    out = np.array(out)
    print(f"--- {gait or 'trot'} h={horizon}")
    for lo_, hi_ in [(0, 0), (1, 4), (5, 8), (9, 16), (17, 32), (33, 1000)]:
        sel = (out[:, 0] >= lo_) & (out[:, 0] <= hi_)
        if not sel.any(): continue
        r = out[sel, 1]
        print(f"  nviol {lo_:3d}-{hi_:4d}: n={sel.sum():4d}  ok<=2 {np.mean((r>0)&(r<=2)):.2f} ok<=3 {np.mean((r>0)&(r<=3)):.2f} ok<=4 {np.mean((r>0)&(r<=4)):.2f} ok<=6 {np.mean((r>0)&(r<=6)):.2f} fail12 {np.mean(r<0):.2f}")

if __name__ == "__main__":
    main(300); main(200, "pace"); main(200, "bound"); main(100, "walk"); main(100, "stand"); main(100, None, 20)
