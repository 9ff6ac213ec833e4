"""Experiment: does the TYPE of row violated by the unconstrained minimiser (friction cone vs fz bound)
predict whether the cold-start active-set iteration converges?  (oracle only)"""
import os, sys, pickle
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(REPO, "robot-gym_b200")); sys.path.insert(0, REPO)
import numpy as np
from robot_gym.model.robots.descriptions import GHOST, with_gait
from robot_gym.util import synthetic
from oracle import convex_mpc as cm
from tools.experiments.pdas_policy2 import pdas_trace


def collect(n, gait, h=10):
    desc = GHOST if gait is None else with_gait(GHOST, gait)
    ctrl = desc.GetCtrlConstants()
    st = synthetic.make_states(4096, desc)
    mp = cm.MpcParams(horizon=h)
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
        pm, qv, cmx, lo, hi = qp.p_mat[np.ix_(fidx, fidx)], qp.q_vec[fidx], qp.c_mat[np.ix_(ridx, fidx)], qp.lb[ridx], qp.ub[ridx]
        x0 = np.linalg.solve(pm, -qv); cx = cmx @ x0
        tol = 1e-9 * hi.max()
        viol = (cx > hi + tol) | (cx < lo - tol)
        kind = np.arange(len(hi)) % 5
        seq, ok = pdas_trace(pm, qv, cmx, lo, hi)
        out.append((int(viol[kind < 4].sum()), int(viol[kind == 4].sum()), seq, ok))
    return out


if __name__ == "__main__":
    cache = os.path.join(REPO, "gpurun_out", "pdas_traces3.pkl")
    if os.path.exists(cache):
        data = pickle.load(open(cache, "rb"))
    else:
        data = {g or "trot": collect(n, g) for g, n in ((None, 300), ("pace", 200), ("bound", 200), ("walk", 120))}
        pickle.dump(data, open(cache, "wb"))
    for g, tr in data.items():
        print(f"--- {g}")
        cone = np.array([t[0] for t in tr]); fz = np.array([t[1] for t in tr])
        rounds = np.array([len(t[2]) if t[3] else 99 for t in tr])
        for lo_, hi_ in [(0, 0), (1, 2), (3, 4), (5, 8), (9, 16), (17, 99)]:
            sel = (cone >= lo_) & (cone <= hi_)
            if sel.any():
                print(f"   cone rows violated {lo_:2d}-{hi_:2d}: n={sel.sum():3d} mean fz viol {fz[sel].mean():5.1f}  ok<=4 {np.mean(rounds[sel] <= 4):.2f} ok<=6 {np.mean(rounds[sel] <= 6):.2f} ok<=8 {np.mean(rounds[sel] <= 8):.2f}  mean rounds if ok {rounds[sel][rounds[sel] < 99].mean() if (rounds[sel] < 99).any() else float('nan'):.1f}")
