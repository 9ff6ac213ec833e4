"""Experiment: evaluate cold-start stopping rules on recorded PDAS change sequences (oracle only).
Cost model: a PDAS round = 1 unit, the interior-point fallback = 12 units (5-7 iterations of
factor + 2 solves, then 1-2 polish rounds)."""
import os, sys, pickle
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(REPO, "robot-gym_b200")); sys.path.insert(0, REPO)
import numpy as np
from robot_gym.model.robots.descriptions import GHOST, with_gait
from robot_gym.util import synthetic
from oracle import convex_mpc as cm


def pdas_trace(pm, qv, cmx, lo, hi, max_rounds=16):
    """Mimics the CUDA polish: <= 3 independent rows per block are held, the rest dropped."""
    side = np.zeros(len(hi), dtype=np.int64)
    feas_tol = 1e-9 * float(np.abs(hi).max())
    seq = []
    for rnd in range(max_rounds):
        rows = np.flatnonzero(side)
        b_act = np.where(side[rows] > 0, hi[rows], lo[rows])
        xp, yp = cm._solve_equality_qp(pm, qv, cmx[rows], b_act)
        cxp = cmx @ xp
        new = side.copy()
        viol = ((cxp - hi > feas_tol) | (lo - cxp > feas_tol)) & (side == 0)
        new[(cxp - hi > feas_tol) & (side == 0)] = 1
        new[(lo - cxp > feas_tol) & (side == 0)] = -1
        # rows held but infeasible (dependent rows the least squares could not satisfy) count as changes too
        held_bad = (side != 0) & ((cxp - hi > feas_tol) | (lo - cxp > feas_tol))
        wrong = (side[rows] * yp) < -1e-10 * max(1.0, float(np.abs(qv).max()))
        new[rows[wrong]] = 0
        nchg = int(np.count_nonzero(new != side)) + int(np.count_nonzero(held_bad))
        seq.append(nchg)
        if nchg == 0:
            return seq, True
        side = new
    return seq, False


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
        out.append(pdas_trace(qp.p_mat[np.ix_(fidx, fidx)], qp.q_vec[fidx], qp.c_mat[np.ix_(ridx, fidx)], qp.lb[ridx], qp.ub[ridx]))
    return out


def cost(seq, ok, max_rounds, max_viol, need_decrease, ipm=12.0):
    """rounds spent + fallback if the rule gives up before convergence."""
    for r, n in enumerate(seq):
        if n == 0: return r + 1, False
        if r + 1 >= max_rounds: return r + 1 + ipm, True
        if r == 0 and n > max_viol: return 1 + ipm, True
        if need_decrease and r >= 1 and n >= seq[r - 1] and n > need_decrease: return r + 1 + ipm, True
    return len(seq) + ipm, True


if __name__ == "__main__":
    cache = os.path.join(REPO, "gpurun_out", "pdas_traces.pkl")
    if os.path.exists(cache):
        data = pickle.load(open(cache, "rb"))
    else:
        data = {g or "trot": collect(n, g) for g, n in ((None, 400), ("pace", 200), ("bound", 200), ("walk", 150), ("stand", 100))}
        os.makedirs(os.path.dirname(cache), exist_ok=True)
        pickle.dump(data, open(cache, "wb"))
    rules = [("current 4 rounds, viol<=16", 4, 16, 0), ("6 rounds, viol<=16", 6, 16, 0), ("6 rounds, viol<=32", 6, 32, 0), ("8 rounds, viol<=32", 8, 32, 0),
             ("8 rounds, viol<=32, stop if no decrease (>2)", 8, 32, 2), ("10 rounds, viol<=48, stop if no decrease (>2)", 10, 48, 2),
             ("10 rounds, any viol, stop if no decrease (>4)", 10, 10**6, 4), ("no cold start", 0, -1, 0)]
    for g, traces in data.items():
        print(f"--- {g}: {len(traces)} envs")
        for name, mr, mv, nd in rules:
            if mr == 0:
                print(f"   {name:50s} mean cost 12.00  fallback 1.00  max 12"); continue
            cs = [cost(s, ok, mr, mv, nd) for s, ok in traces]
            c = np.array([x[0] for x in cs]); fb = np.mean([x[1] for x in cs])
            print(f"   {name:50s} mean cost {c.mean():5.2f}  fallback {fb:.2f}  max {c.max():.0f}  p99 {np.percentile(c, 99):.0f}")
