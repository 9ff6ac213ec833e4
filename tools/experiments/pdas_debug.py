import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(REPO, "robot-gym_b200")); sys.path.insert(0, REPO)
import numpy as np
from robot_gym.model.robots.descriptions import GHOST
from robot_gym.util import synthetic
from oracle import convex_mpc as cm
from tools.experiments.active_set_stats import pdas
desc = GHOST; ctrl = desc.GetCtrlConstants()
st = synthetic.make_states(4096, desc); mp = cm.MpcParams(horizon=10)
for i in range(300):
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
    r, xp, side = pdas(pm, qv, cmx, lo, hi, np.zeros(len(hi), dtype=np.int64))
    xref, info = cm.solve_qp(qp.p_mat, qp.q_vec, qp.c_mat, qp.lb, qp.ub)
    err = np.abs(xp - xref[fidx]).max() / max(1.0, np.abs(xref).max())
    if err > 1e-6:
        obj = lambda x: 0.5 * x @ pm @ x + qv @ x
        xr = xref[fidx]
        cert_p = cm.kkt_certificate(pm, qv, cmx, lo, hi, xp); cert_r = cm.kkt_certificate(pm, qv, cmx, lo, hi, xr)
        print(i, "rounds", r, "err", err, "polished", info.get("polished"), "obj pdas-ref", obj(xp) - obj(xr),
              "\n  pdas cert", {k: v for k, v in cert_p.items() if k != 'y'}, "\n  ref cert", {k: v for k, v in cert_r.items() if k != 'y'},
              "\n  nact pdas", np.count_nonzero(side))
