"""GPU robustness sweep (development helper): polished fraction / iteration statistics / worst error vs
the C oracle over contact schedules, horizons and weight sets."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "robot-gym_b200")); sys.path.insert(0, REPO)
import numpy as np, torch
from robot_gym import cuda as rg
from robot_gym.model.robots.descriptions import GHOST, with_gait
from robot_gym.util import synthetic
from oracle import convex_mpc as cm, c_oracle

W2 = (5, 5, 0.2, 0, 0, 10, 0., 0., 1., 1., 1., 0., 0)
for sched, horizon, weights, n in [("trot", 10, None, 16384), ("pace", 10, None, 8192), ("bound", 10, None, 8192), ("walk", 10, None, 8192),
                                   ("trot", 5, None, 8192), ("trot", 20, None, 2048), ("bound", 20, None, 1024), ("trot", 10, W2, 8192), ("pace", 5, W2, 4096)]:
    desc = with_gait(GHOST, sched); ctrl = desc.GetCtrlConstants()
    st = synthetic.make_states(n, desc, schedule_ctrl=ctrl, seed=77)
    p = rg.default_mpc_params(ctrl.MPC_BODY_MASS, ctrl.MPC_BODY_INERTIA, ctrl.MPC_BODY_HEIGHT, horizon)
    mp = cm.MpcParams(horizon=horizon)
    if weights:
        for i, w in enumerate(weights): p.weights[i] = w
        mp.weights = weights
    ws = rg.MpcWorkspace(p)
    t = lambda a: torch.from_numpy(a).cuda()
    f, hf, info = rg.mpc_build_solve(ws, t(st.com_velocity_body), t(st.base_rpy), t(st.base_rpy_rate), t(st.planned_contacts), t(st.foot_positions_base), t(st.command))
    info = info.cpu().numpy(); f = f.cpu().numpy()
    m = min(n, 2048)
    ref, _, _ = c_oracle.solve_batch(mp, st.slice(0, m), ctrl.MPC_BODY_HEIGHT, n_threads=os.cpu_count())
    err = np.abs(f[:m] - ref).max(axis=1) / np.maximum(1, np.abs(ref).max(axis=1))
    unp = np.flatnonzero((info[:, 2] & 1) == 0)
    print(f"{sched:6s} h={horizon:2d} w={'2' if weights else '1'} n={n}: iters mean {info[:,0].mean():.2f} max {info[:,0].max()} polish mean {info[:,1].mean():.2f} max {info[:,1].max()} "
          f"unpolished {len(unp)} numeric {int(((info[:,2]&8)!=0).sum())} | err vs C oracle (first {m}): p50 {np.median(err):.1e} max {err.max():.1e} (#>1e-4: {(err>1e-4).sum()})")
    for i in unp[:5]:
        o = cm.compute_contact_forces(mp, st.com_velocity_body[i].astype(np.float64), st.base_rpy[i].astype(np.float64), st.base_rpy_rate[i].astype(np.float64), st.planned_contacts[i], st.foot_positions_base[i].astype(np.float64), [0,0,ctrl.MPC_BODY_HEIGHT],[st.command[i,0],st.command[i,1],0.0],[0,0,0],[0,0,float(st.command[i,2])])
        print("     unpolished env", i, info[i], "err vs numpy oracle %.2e" % (np.abs(f[i]-o[:12]).max()/max(1,np.abs(o[:12]).max())))
