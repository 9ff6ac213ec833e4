"""Experiment: do damped variants of the primal-dual active-set iteration converge where the plain one cycles?
(oracle only; emulates the CUDA rounds incl. the last-step guess)"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(REPO, "robot-gym_b200")); sys.path.insert(0, REPO)
import numpy as np
from robot_gym.model.robots.descriptions import GHOST, with_gait
from robot_gym.util import synthetic
from oracle import convex_mpc as cm


def run(pm, qv, cmx, lo, hi, variant, side0, max_rounds=24):
    side = side0.copy()
    feas_tol = 1e-9 * float(np.abs(hi).max())
    qs = max(1.0, float(np.abs(qv).max()))
    for rnd in range(1, max_rounds + 1):
        rows = np.flatnonzero(side)
        b_act = np.where(side[rows] > 0, hi[rows], lo[rows])
        xp, yp = cm._solve_equality_qp(pm, qv, cmx[rows], b_act)
        cxp = cmx @ xp
        vio = np.maximum(cxp - hi, lo - cxp); vio[rows] = 0.0
        add = np.flatnonzero(vio > feas_tol)
        drop = rows[(side[rows] * yp) < -1e-10 * qs]
        if len(add) == 0 and len(drop) == 0:
            held_bad = np.maximum(cxp - hi, lo - cxp)[rows]
            if len(rows) and held_bad.max() > 1e-6: return -1     # dependent rows: not a real convergence
            return rnd
        if variant == 1 and len(add) > 4:                         # at most the 4 most violated rows
            add = add[np.argsort(-vio[add])[:4]]
        if variant == 2 and len(add):                             # one (the most violated) row per block
            keep = {}
            for r in add:
                b = r // 5
                if b not in keep or vio[r] > vio[keep[b]]: keep[b] = r
            add = np.array(sorted(keep.values()))
        if variant == 3 and len(add) > 1:                         # half of them
            add = add[np.argsort(-vio[add])[:max(1, len(add) // 2)]]
        if variant in (4, 5) and len(add):
            keep = {}
            for r in add:
                b = r // 5
                if b not in keep or vio[r] > vio[keep[b]]: keep[b] = r
            add = np.array(sorted(keep.values()))
        if variant == 6 and len(add):                             # the most violated row per (block, direction group)
            keep = {}
            for r in add:
                key = (r // 5, min((r % 5) // 2, 2))             # rows 0,1 -> x cone; 2,3 -> y cone; 4 -> fz
                if key not in keep or vio[r] > vio[keep[key]]: keep[key] = r
            add = np.array(sorted(keep.values()))
        if variant == 4 and len(drop) > 1:                        # ... and drop only the worst multiplier per block
            ysg = dict(zip(rows, side[rows] * yp)); keep = {}
            for r in drop:
                b = r // 5
                if b not in keep or ysg[r] < ysg[keep[b]]: keep[b] = r
            drop = np.array(sorted(keep.values()))
        if variant == 5 and len(add) and len(drop):               # ... and never add and drop in the same block in one round
            dblocks = set(int(r) // 5 for r in drop)
            add = np.array([r for r in add if int(r) // 5 not in dblocks], dtype=int)
        for r in add: side[r] = 1 if cxp[r] > hi[r] else -1
        side[drop] = 0
    return -1


def main(n, gait):
    desc = with_gait(GHOST, gait); ctrl = desc.GetCtrlConstants()
    st = synthetic.make_states(4096, desc, schedule_ctrl=ctrl)
    mp = cm.MpcParams(horizon=10)
    res = {v: [] for v in (0, 2, 6)}
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
        nleg = int(free[:4].sum()) if nblk >= 4 else 0
        side0[-5 * nleg:][4::5] = -1                              # fz >= fz_min at the last step, every stance leg
        for v in res: res[v].append(run(pm, qv, cmx, lo, hi, v, side0))
    names = {0: "plain", 2: "one row per block", 6: "one row per block and direction"}
    print(f"--- {gait}: {n} envs")
    for v, r in res.items():
        r = np.array(r); ok = r > 0
        print(f"   {names[v]:26s} converged<=24: {ok.mean():.3f}  <=8: {np.mean(ok & (r <= 8)):.3f}  <=12: {np.mean(ok & (r <= 12)):.3f}  <=16: {np.mean(ok & (r <= 16)):.3f}  mean rounds (ok) {r[ok].mean():.2f}")
    r0 = np.array(res[0]); hard = ~((r0 > 0) & (r0 <= 5))
    for v in (2, 6):
        r = np.array(res[v]); print(f"   of the {hard.sum()} envs plain does not settle in 5 rounds, '{names[v]}' settles {np.sum(hard & (r > 0) & (r <= 10))} within 10")


if __name__ == "__main__":
    main(150, "bound"); main(150, "pace"); main(200, "trot"); main(100, "walk")
