"""Experiment (oracle only): does copying rows along a leg's time axis save the cascade rounds of the long envs?

The long trot envs unload one leg over most of the horizon and its blocks walk to a corner of their pyramid one row
per round (tools/experiments/cascade_trace.py).  Variants of the kernel's round (one violated row per block) that also
hand a held / newly admitted row to the neighbouring steps of the same leg, to every step of the leg that already
holds a row, ...; a row that was dropped once is never handed on again (without that guard 6-22 % of the envs cycle).
    python tools/experiments/row_propagation.py [n_env] [gait]
Result (400 trot envs): mean rounds 1.73 -> 1.81-2.35, the envs with >= 5 rounds 1 -> 15-50: every wrongly copied row
costs the round it was meant to save, and more."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(REPO, "robot-gym_b200")); sys.path.insert(0, REPO)
import numpy as np
from robot_gym.model.robots.descriptions import GHOST, with_gait
from robot_gym.util import synthetic
from oracle import convex_mpc as cm

def rounds_of(pm, qv, cmx, lo, hi, side0, blk, variant, max_rounds=14):
    side = side0.copy()
    feas_tol = 1e-9 * float(np.abs(hi).max()); qs = max(1.0, float(np.abs(qv).max()))
    pos = {tl: k for k, tl in enumerate(blk)}
    banned = set()
    for rnd in range(1, max_rounds + 1):
        rows = np.flatnonzero(side)
        xp, yp = cm._solve_equality_qp(pm, qv, cmx[rows], np.where(side[rows] > 0, hi[rows], lo[rows]))
        cxp = cmx @ xp
        vio = np.maximum(cxp - hi, lo - cxp); vio[rows] = 0.0
        add = np.flatnonzero(vio > feas_tol)
        drop = rows[(side[rows] * yp) < -1e-10 * qs]
        if len(add) == 0 and len(drop) == 0: return rnd
        keep = {}
        for r in add:
            if r // 5 not in keep or vio[r] > vio[keep[r // 5]]: keep[r // 5] = r
        add = np.array(sorted(keep.values()), dtype=int)
        new = side.copy()
        for r in add: new[r] = 1 if cxp[r] > hi[r] else -1
        new[drop] = 0
        if variant > 0:
            base = new.copy()
            dropped_now = set(int(d) for d in drop); banned |= dropped_now
            for k, (t, l) in enumerate(blk):
                for dt in ((-1, 1) if variant in (1, 2, 4) else range(-9, 10)):
                    if dt == 0: continue
                    k2 = pos.get((t + dt, l))
                    if k2 is None: continue
                    if variant in (1, 3, 4) and not np.any(base[5*k2:5*k2+5]): continue   # neighbour must be in play
                    for j in range(5):
                        s = base[5*k + j]
                        if s == 0: continue
                        if variant == 4 and not (side[5*k+j] == 0): continue   # only rows newly added this round propagate
                        if new[5*k2 + j] != 0: continue
                        if (5*k2 + j) in banned: continue
                        if j < 4 and new[5*k2 + (j ^ 1)] != 0: continue
                        if np.count_nonzero(new[5*k2:5*k2+5]) >= 3: continue
                        new[5*k2 + j] = s
        side = new
    return -1

def main(n, gait):
    desc = GHOST if gait == "trot" else with_gait(GHOST, gait); ctrl = desc.GetCtrlConstants()
    st = synthetic.make_states(4096, desc, schedule_ctrl=ctrl)
    mp = cm.MpcParams(horizon=10)
    res = {v: [] for v in (0, 1, 2, 3, 4)}
    for i in range(n):
        qp = cm.build_qp(mp, st.com_velocity_body[i].astype(np.float64), st.base_rpy[i].astype(np.float64), st.base_rpy_rate[i].astype(np.float64),
                         st.planned_contacts[i], st.foot_positions_base[i].astype(np.float64), [0, 0, ctrl.MPC_BODY_HEIGHT],
                         [st.command[i, 0], st.command[i, 1], 0.0], [0, 0, 0], [0, 0, float(st.command[i, 2])])
        nblk = qp.p_mat.shape[0] // 3
        free = np.array([not np.all(qp.ub[5*b:5*b+5] == qp.lb[5*b:5*b+5]) for b in range(nblk)])
        fidx = np.flatnonzero(np.repeat(free, 3)); ridx = np.flatnonzero(np.repeat(free, 5))
        blk = [(b // 4, b % 4) for b in range(nblk) if free[b]]
        if not blk: continue
        pm, qv, cmx, lo, hi = qp.p_mat[np.ix_(fidx, fidx)], qp.q_vec[fidx], qp.c_mat[np.ix_(ridx, fidx)], qp.lb[ridx], qp.ub[ridx]
        side0 = np.zeros(len(hi), dtype=np.int64)
        n4 = int(free[:4].sum())
        for k, (t, _) in enumerate(blk):
            if t >= 10 - (2 if n4 == 4 else 1): side0[5 * k + 4] = -1
        for v in res: res[v].append(rounds_of(pm, qv, cmx, lo, hi, side0, blk, v))
    names = {0: "one row per block (kernel)", 1: "+ held rows -> t+-1 same leg, in-play only", 2: "+ held rows -> t+-1 same leg, any", 3: "+ union over all in-play steps of the leg", 4: "+ NEW rows -> t+-1 in-play"}
    print(f"--- {gait} {n}")
    for v, r in res.items():
        r = np.array(r); ok = r > 0
        print(f"  {names[v]:46s} ok {ok.mean():.3f} mean {r[ok].mean():.3f} hist {np.bincount(r[ok])}")

if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 400, sys.argv[2] if len(sys.argv) > 2 else "trot")
