"""Experiment (oracle only): can a "rows still moving" rule hand the bound / pace envs that will not settle to the
interior point before they have burnt the cold-start budget?  For rules "after round R the number of moving rows is
still >= f x its first value": how many failing envs they catch, the rounds saved, and the converging envs sent away.
Result (250 envs each): bound -- R = 5, f = 0.7 catches 22 of 30 failing envs and saves 154 rounds but sends 8 converging
envs to the interior point; pace -- every rule sends away 10-50 converging envs for 2 failing ones.  Not built.
    python tools/experiments/stall_rule.py [n_env]"""
import os, sys
import numpy as np
import importlib.util
spec = importlib.util.spec_from_file_location("cd", os.path.join(os.path.dirname(os.path.abspath(__file__)), "cycle_detect.py")); cd = importlib.util.module_from_spec(spec); spec.loader.exec_module(cd)
from robot_gym.model.robots.descriptions import GHOST, with_gait
from robot_gym.util import synthetic
from oracle import convex_mpc as cm
def collect(n, gait):
    desc = with_gait(GHOST, gait); ctrl = desc.GetCtrlConstants()
    st = synthetic.make_states(4096, desc, schedule_ctrl=ctrl)
    mp = cm.MpcParams(horizon=10); out = []
    for i in range(n):
        qp = cm.build_qp(mp, st.com_velocity_body[i].astype(np.float64), st.base_rpy[i].astype(np.float64), st.base_rpy_rate[i].astype(np.float64),
                         st.planned_contacts[i], st.foot_positions_base[i].astype(np.float64), [0, 0, ctrl.MPC_BODY_HEIGHT],
                         [st.command[i, 0], st.command[i, 1], 0.0], [0, 0, 0], [0, 0, float(st.command[i, 2])])
        nblk = qp.p_mat.shape[0] // 3
        free = np.array([not np.all(qp.ub[5*b:5*b+5] == qp.lb[5*b:5*b+5]) for b in range(nblk)])
        fidx = np.flatnonzero(np.repeat(free, 3)); ridx = np.flatnonzero(np.repeat(free, 5))
        if len(fidx) == 0: continue
        pm, qv, cmx, lo, hi = qp.p_mat[np.ix_(fidx, fidx)], qp.q_vec[fidx], qp.c_mat[np.ix_(ridx, fidx)], qp.lb[ridx], qp.ub[ridx]
        side0 = np.zeros(len(hi), dtype=np.int64); nleg = int(free[-4:].sum()); side0[-5 * nleg:][4::5] = -1
        out.append(cd.run(pm, qv, cmx, lo, hi, side0, max_rounds=12))
    return out
for gait in ("bound", "pace"):
    out = collect(int(sys.argv[1]) if len(sys.argv) > 1 else 250, gait)
    conv = np.array([o[0] for o in out]); 
    print(f"--- {gait}: converged<=12 {np.mean(conv>0):.3f}")
    for R in (3, 4, 5, 6):
        for frac in (0.5, 0.7, 0.9):
            flagged = np.array([len(o[2]) > R and o[2][R] >= frac * o[2][0] for o in out])   # still running after R+1 rounds and nchg not shrunk
            fail = conv < 0
            wasted_if_flag = sum((12 - (R + 1)) for o, f in zip(out, flagged) if f and o[0] < 0)
            lost = [o[0] for o, f in zip(out, flagged) if f and o[0] > 0]
            print(f"   rule round>={R+1}, nchg >= {frac} nchg0: flags {flagged.sum()} of which truly failing {np.sum(flagged & fail)} / {fail.sum()} failing; rounds saved {wasted_if_flag}; converging envs sent away {len(lost)} (they'd need {lost})")
