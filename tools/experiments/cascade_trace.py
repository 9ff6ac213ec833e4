"""Experiment (oracle only): what are the rounds 3+ of the long trot envs made of?

Emulates the kernel's cold-start rounds (last-step guess, one row per block per round) on the bench batch and
classifies every row admitted from round 3 on: does it join a block that already holds a row (the block is walking
towards a vertex of its pyramid: face -> edge -> corner, one row per round), a block of a LEG that already holds
rows at other steps, or a fresh leg?   python tools/experiments/cascade_trace.py [n_env] [gait]"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(REPO, "robot-gym_b200")); sys.path.insert(0, REPO)
import numpy as np
from robot_gym.model.robots.descriptions import GHOST, with_gait
from robot_gym.util import synthetic
from oracle import convex_mpc as cm


def rounds_of(pm, qv, cmx, lo, hi, side0, max_rounds=14):
    side = side0.copy(); hist = []
    feas_tol = 1e-9 * float(np.abs(hi).max()); qs = max(1.0, float(np.abs(qv).max()))
    for rnd in range(1, max_rounds + 1):
        rows = np.flatnonzero(side)
        xp, yp = cm._solve_equality_qp(pm, qv, cmx[rows], np.where(side[rows] > 0, hi[rows], lo[rows]))
        cxp = cmx @ xp
        vio = np.maximum(cxp - hi, lo - cxp); vio[rows] = 0.0
        add = np.flatnonzero(vio > feas_tol)
        drop = rows[(side[rows] * yp) < -1e-10 * qs]
        if len(add) == 0 and len(drop) == 0: return rnd, hist
        keep = {}
        for r in add:
            if r // 5 not in keep or vio[r] > vio[keep[r // 5]]: keep[r // 5] = r
        add = np.array(sorted(keep.values()), dtype=int)
        hist.append((side.copy(), add, drop))
        for r in add: side[r] = 1 if cxp[r] > hi[r] else -1
        side[drop] = 0
    return -1, hist


def main(n, gait):
    desc = GHOST if gait == "trot" else with_gait(GHOST, gait); ctrl = desc.GetCtrlConstants()
    st = synthetic.make_states(4096, desc, schedule_ctrl=ctrl)
    mp = cm.MpcParams(horizon=10)
    same_block = same_leg = fresh = 0; rounds = []; legs_at_vertex = []
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
        r, hist = rounds_of(pm, qv, cmx, lo, hi, side0)
        rounds.append(r)
        if r < 4: continue
        for side, add, _ in hist[2:]:
            held_blocks = set(np.flatnonzero(side) // 5)
            held_legs = set(blk[b][1] for b in held_blocks if blk[b][0] < 8)
            for a in add:
                if a // 5 in held_blocks: same_block += 1
                elif blk[a // 5][1] in held_legs: same_leg += 1
                else: fresh += 1
        final = hist[-1][0]
        nrows = np.add.reduceat(np.abs(final), np.arange(0, len(final), 5))
        legs_at_vertex.append(len(set(blk[b][1] for b in np.flatnonzero(nrows >= 2))))
    rounds = np.array(rounds)
    tot = max(1, same_block + same_leg + fresh)
    print(f"{gait}: {len(rounds)} envs, rounds hist {np.bincount(rounds[rounds > 0])}, mean {rounds[rounds > 0].mean():.2f}")
    print(f"   rows admitted in rounds 3+ of the >= 4-round envs: {same_block} ({100*same_block/tot:.0f} %) join a block that already holds a row, "
          f"{same_leg} ({100*same_leg/tot:.0f} %) a new step of a leg that holds rows elsewhere, {fresh} ({100*fresh/tot:.0f} %) a fresh leg")
    print(f"   legs with blocks holding >= 2 rows at the end, per long env: {np.bincount(legs_at_vertex)}")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 768, sys.argv[2] if len(sys.argv) > 2 else "trot")
