"""Experiment: which rows of a (step, leg) block should enter the active set together?  (oracle only)
The CUDA rounds add ONE violated row per block per round; envs where a leg unloads (fz wants to go below fz_min
over the whole horizon, dragging its friction rows along) then need 3-4 rounds for one block (z, then x, then y).
Variants emulated here, all with the last-step guess:
   2  one (most violated) row per block                              (the CUDA kernel in round 1)
   6  most violated row per (block, direction group x / y / z)
   7  the rows active at the EUCLIDEAN PROJECTION of the block's force onto its truncated friction pyramid
   8  like 7, but only used when the fz >= fz_min row of the block is among the violated rows (else rule 2)

    python tools/experiments/pdas_block_rules.py [n] [gait]
"""
import os, sys, itertools, multiprocessing as mp_
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(REPO, "robot-gym_b200")); sys.path.insert(0, REPO)
import numpy as np
from robot_gym.model.robots.descriptions import GHOST, with_gait
from robot_gym.util import synthetic
from oracle import convex_mpc as cm

GAIT = sys.argv[2] if len(sys.argv) > 2 else "trot"
DESC = GHOST if GAIT == "trot" else with_gait(GHOST, GAIT)
CTRL = DESC.GetCtrlConstants()
ST = synthetic.make_states(4096, DESC, schedule_ctrl=CTRL)
MP = cm.MpcParams(horizon=10)
MU = 0.45


def project_rows(f, fzmin, fzmax):
    """Rows (0..4 = pyramid rows, sign +1 upper / -1 lower) active at the Euclidean projection of f onto
    {|fx| <= mu fz, |fy| <= mu fz, fzmin <= fz <= fzmax}: enumerate faces (<= 3 rows), keep the closest feasible point."""
    g = np.array([[-1, 0, MU], [1, 0, MU], [0, -1, MU], [0, 1, MU], [0, 0, 1.0]])
    cand = [(0, -1), (1, -1), (2, -1), (3, -1), (4, -1), (4, +1)]      # cone rows at their lower bound 0, fz at min / max
    best, best_set = None, ()
    for k in range(0, 4):
        for combo in itertools.combinations(cand, k):
            if len({c[0] for c in combo}) < k: continue
            if k:
                a = np.array([g[c[0]] for c in combo]); b = np.array([0.0 if c[0] < 4 else (fzmin if c[1] < 0 else fzmax) for c in combo])
                if np.linalg.matrix_rank(a) < k: continue
                p = f - a.T @ np.linalg.solve(a @ a.T, a @ f - b)
            else:
                p = f.copy()
            c5 = g @ p
            if np.all(c5[:4] >= -1e-9) and fzmin - 1e-9 <= c5[4] <= fzmax + 1e-9:
                d = np.sum((p - f) ** 2)
                if best is None or d < best - 1e-12: best, best_set = d, combo
    return best_set


def run(pm, qv, cmx, lo, hi, variant, side0, max_rounds=24):
    side = side0.copy()
    feas_tol = 1e-9 * float(np.abs(hi).max())
    qs = max(1.0, float(np.abs(qv).max()))
    fzmin, fzmax = float(lo[4]), float(hi[4])
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
            if len(rows) and held_bad.max() > 1e-6: return -1
            return rnd
        new_side = {}
        blocks = sorted({int(r) // 5 for r in add})
        for b in blocks:
            radd = [r for r in add if r // 5 == b]
            zviol = any(r % 5 == 4 and cxp[r] < lo[r] for r in radd)
            if variant == 2 or (variant == 8 and not zviol):
                r = max(radd, key=lambda r: vio[r]); new_side[r] = 1 if cxp[r] > hi[r] else -1
            elif variant == 6:
                keep = {}
                for r in radd:
                    key = min((r % 5) // 2, 2)
                    if key not in keep or vio[r] > vio[keep[key]]: keep[key] = r
                for r in keep.values(): new_side[r] = 1 if cxp[r] > hi[r] else -1
            else:
                f = xp[3 * b:3 * b + 3]
                for (row, sgn) in project_rows(f, fzmin, fzmax):
                    r = 5 * b + row
                    if side[r] == 0: new_side[r] = sgn
                # rows of the block held so far that the projection does not keep are released
                if variant == 7:
                    keepers = {5 * b + row for (row, _) in project_rows(f, fzmin, fzmax)}
                    for r in range(5 * b, 5 * b + 5):
                        if side[r] != 0 and r not in keepers: side[r] = 0
        for r, s in new_side.items(): side[r] = s
        side[drop] = 0
    return -1


def one(i):
    st = ST
    qp = cm.build_qp(MP, st.com_velocity_body[i].astype(np.float64), st.base_rpy[i].astype(np.float64), st.base_rpy_rate[i].astype(np.float64),
                     st.planned_contacts[i], st.foot_positions_base[i].astype(np.float64), [0, 0, CTRL.MPC_BODY_HEIGHT],
                     [st.command[i, 0], st.command[i, 1], 0.0], [0, 0, 0], [0, 0, float(st.command[i, 2])])
    nblk = qp.p_mat.shape[0] // 3
    free = np.array([not np.all(qp.ub[5*b:5*b+5] == qp.lb[5*b:5*b+5]) for b in range(nblk)])
    fidx = np.flatnonzero(np.repeat(free, 3)); ridx = np.flatnonzero(np.repeat(free, 5))
    if len(fidx) == 0: return None
    pm, qv, cmx, lo, hi = qp.p_mat[np.ix_(fidx, fidx)], qp.q_vec[fidx], qp.c_mat[np.ix_(ridx, fidx)], qp.lb[ridx], qp.ub[ridx]
    side0 = np.zeros(len(hi), dtype=np.int64)
    nleg = int(free[:4].sum())
    side0[-5 * nleg:][4::5] = -1
    return [run(pm, qv, cmx, lo, hi, v, side0) for v in (2, 6, 7, 8)]


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    with mp_.Pool(os.cpu_count()) as pool:
        res = [r for r in pool.map(one, range(n), chunksize=8) if r is not None]
    res = np.array(res)
    print(f"--- {GAIT}: {len(res)} envs")
    for k, name in enumerate(("2 one row per block", "6 per block+direction", "7 projection (replace)", "8 projection when fz_min violated")):
        r = res[:, k]; ok = r > 0
        print(f"  {name:36s} converged {ok.mean():.3f}  mean rounds {r[ok].mean():.3f}  hist {np.bincount(r[ok], minlength=8)[:13]}")
