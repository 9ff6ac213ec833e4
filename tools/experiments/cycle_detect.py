"""Experiment (oracle only): when does the one-row-per-block active-set iteration first revisit a state (bound / pace,
h = 10)?  Result (200 bound envs): 87 % settle within 12 rounds, and of the 26 that do not NONE repeats an active set
within 12 rounds (first repeats at round 12-20, period 4 or 8): a hash of the set cannot hand a cycling env over early.
    python tools/experiments/cycle_detect.py [n_env] [gait]"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(REPO, "robot-gym_b200")); sys.path.insert(0, REPO)
import numpy as np
from robot_gym.model.robots.descriptions import GHOST, with_gait
from robot_gym.util import synthetic
from oracle import convex_mpc as cm

def run(pm, qv, cmx, lo, hi, side0, max_rounds=24):
    side = side0.copy()
    feas_tol = 1e-9 * float(np.abs(hi).max()); qs = max(1.0, float(np.abs(qv).max()))
    seen = {side.tobytes(): 0}; first_rep = None; nchgs = []
    for rnd in range(1, max_rounds + 1):
        rows = np.flatnonzero(side)
        b_act = np.where(side[rows] > 0, hi[rows], lo[rows])
        xp, yp = cm._solve_equality_qp(pm, qv, cmx[rows], b_act)
        cxp = cmx @ xp
        vio = np.maximum(cxp - hi, lo - cxp); vio[rows] = 0.0
        add = np.flatnonzero(vio > feas_tol)
        drop = rows[(side[rows] * yp) < -1e-10 * qs]
        if len(add) == 0 and len(drop) == 0: return rnd, first_rep, nchgs
        keep = {}
        for r in add:
            b = r // 5
            if b not in keep or vio[r] > vio[keep[b]]: keep[b] = r
        add = np.array(sorted(keep.values()), dtype=int)
        nchgs.append(len(add) + len(drop))
        for r in add: side[r] = 1 if cxp[r] > hi[r] else -1
        side[drop] = 0
        k = side.tobytes()
        if k in seen and first_rep is None: first_rep = (rnd, rnd - seen[k])
        seen.setdefault(k, rnd)
    return -1, first_rep, nchgs

def main(n, gait):
    desc = with_gait(GHOST, gait); ctrl = desc.GetCtrlConstants()
    st = synthetic.make_states(4096, desc, schedule_ctrl=ctrl)
    mp = cm.MpcParams(horizon=10)
    out = []
    for i in range(n):
        qp = cm.build_qp(mp, st.com_velocity_body[i].astype(np.float64), st.base_rpy[i].astype(np.float64), st.base_rpy_rate[i].astype(np.float64),
                         st.planned_contacts[i], st.foot_positions_base[i].astype(np.float64), [0, 0, ctrl.MPC_BODY_HEIGHT],
                         [st.command[i, 0], st.command[i, 1], 0.0], [0, 0, 0], [0, 0, float(st.command[i, 2])])
        nblk = qp.p_mat.shape[0] // 3
        free = np.array([not np.all(qp.ub[5*b:5*b+5] == qp.lb[5*b:5*b+5]) for b in range(nblk)])
        fidx = np.flatnonzero(np.repeat(free, 3)); ridx = np.flatnonzero(np.repeat(free, 5))
        if len(fidx) == 0: continue
        pm, qv, cmx, lo, hi = qp.p_mat[np.ix_(fidx, fidx)], qp.q_vec[fidx], qp.c_mat[np.ix_(ridx, fidx)], qp.lb[ridx], qp.ub[ridx]
        side0 = np.zeros(len(hi), dtype=np.int64)
        nleg = int(free[-4:].sum())
        side0[-5 * nleg:][4::5] = -1
        out.append(run(pm, qv, cmx, lo, hi, side0))
    r = np.array([o[0] for o in out])
    print(f"--- {gait}: {len(out)} envs; converged <=12: {np.mean((r>0)&(r<=12)):.3f}  <=24: {np.mean(r>0):.3f}; rounds hist (ok): {np.bincount(r[r>0])}")
    bad = [o for o in out if o[0] < 0 or o[0] > 12]
    print(f"   {len(bad)} not converged within 12; first repeat (round, period): {[o[1] for o in bad]}")
    ok_rep = [o for o in out if 0 < o[0] <= 12 and o[1] is not None]
    print(f"   converged-within-12 envs that nevertheless revisited a state: {len(ok_rep)} {[ (o[0], o[1]) for o in ok_rep][:10]}")
    print("   nchg traces of the bad ones:"); 
    for o in bad[:12]: print("     ", o[2][:16])

if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 200, sys.argv[2] if len(sys.argv) > 2 else "bound")
