"""Experiment: can a cheap function of the INPUTS predict how many active-set rounds the cold start needs?
(oracle only; emulates the CUDA rounds -- last-step guess, one row per block per round -- in numpy.)
Feeds the longest-first ordering of the solve kernel: the launch tail at 4096 envs is set by 4-6-round envs
that start in the last wave (profiles/r01_timeline_4096.log).

    python tools/experiments/round_predictor.py [n_envs] [gait]
"""
import os, sys, multiprocessing as mp_
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(REPO, "robot-gym_b200")); sys.path.insert(0, REPO); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
from robot_gym.model.robots.descriptions import GHOST, with_gait
from robot_gym.util import synthetic
from oracle import convex_mpc as cm
from pdas_variants import run

GAIT = sys.argv[2] if len(sys.argv) > 2 else "trot"
DESC = GHOST if GAIT == "trot" else with_gait(GHOST, GAIT)
CTRL = DESC.GetCtrlConstants()
ST = synthetic.make_states(4096, DESC, schedule_ctrl=CTRL)
MP = cm.MpcParams(horizon=10)


def one(i):
    st = ST
    qp = cm.build_qp(MP, st.com_velocity_body[i].astype(np.float64), st.base_rpy[i].astype(np.float64), st.base_rpy_rate[i].astype(np.float64),
                     st.planned_contacts[i], st.foot_positions_base[i].astype(np.float64), [0, 0, CTRL.MPC_BODY_HEIGHT],
                     [st.command[i, 0], st.command[i, 1], 0.0], [0, 0, 0], [0, 0, float(st.command[i, 2])])
    nblk = qp.p_mat.shape[0] // 3
    free = np.array([not np.all(qp.ub[5*b:5*b+5] == qp.lb[5*b:5*b+5]) for b in range(nblk)])
    fidx = np.flatnonzero(np.repeat(free, 3)); ridx = np.flatnonzero(np.repeat(free, 5))
    if len(fidx) == 0:
        return (i, 0, 0, 0.0, 0.0)
    pm, qv, cmx, lo, hi = qp.p_mat[np.ix_(fidx, fidx)], qp.q_vec[fidx], qp.c_mat[np.ix_(ridx, fidx)], qp.lb[ridx], qp.ub[ridx]
    side0 = np.zeros(len(hi), dtype=np.int64)
    nleg = int(free[:4].sum())
    side0[-5 * nleg:][4::5] = -1
    r = run(pm, qv, cmx, lo, hi, 2, side0)
    # a cheap proxy of the gradient scale: |q|_inf and the unconstrained-in-fz "demand" sum
    return (i, r, nleg, float(np.abs(qv).max()), float(qp.com_z))


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    with mp_.Pool(os.cpu_count()) as pool:
        rows = pool.map(one, range(n), chunksize=16)
    rows = np.array(rows, dtype=float)
    np.save(os.path.join(REPO, "gpurun_out", f"round_labels_{GAIT}.npy"), rows)
    r = rows[:, 1]
    print("rounds histogram:", np.bincount(np.clip(r.astype(int), 0, 12)), " unconverged:", np.sum(r < 0))
    s = ST.slice(0, n)
    feat = {"n_stance": rows[:, 2], "|q|inf": rows[:, 3], "|roll|": np.abs(s.base_rpy[:, 0]), "|pitch|": np.abs(s.base_rpy[:, 1]),
            "|vx-cmd|": np.abs(s.com_velocity_body[:, 0] - s.command[:, 0]), "|vy-cmd|": np.abs(s.com_velocity_body[:, 1] - s.command[:, 1]),
            "|vz|": np.abs(s.com_velocity_body[:, 2]), "vz": s.com_velocity_body[:, 2], "|wx|": np.abs(s.base_rpy_rate[:, 0]), "|wy|": np.abs(s.base_rpy_rate[:, 1]),
            "|wz-cmd|": np.abs(s.base_rpy_rate[:, 2] - s.command[:, 2]), "com_z-h": rows[:, 4] - CTRL.MPC_BODY_HEIGHT}
    for k, v in feat.items():
        print(f"  corr(rounds, {k:10s}) = {np.corrcoef(r, v)[0, 1]:+.3f}   mean by rounds: " +
              " ".join(f"{v[r == q].mean():+.3f}" for q in range(1, 6) if np.any(r == q)))
    X = np.column_stack(list(feat.values()) + [np.ones(n)])
    w, *_ = np.linalg.lstsq(X, r, rcond=None)
    pred = X @ w
    print("  linear fit R^2:", 1 - np.var(r - pred) / np.var(r))
    hard = r >= 3
    top = np.argsort(-pred)[: int(hard.mean() * n * 1.5)]
    print(f"  envs with >= 3 rounds: {hard.sum()}; captured in the top {len(top)} by the linear score: {hard[top].sum()}")
    for ns in (2, 3, 4):
        m = rows[:, 2] == ns
        if m.any(): print(f"  n_stance {ns}: {m.sum()} envs, rounds hist {np.bincount(r[m].astype(int).clip(0, 8))}")
