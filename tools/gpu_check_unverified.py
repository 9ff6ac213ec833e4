"""GPU result vs the numpy oracle on the envs a sweep left unverified (development helper).
python tools/gpu_check_unverified.py <schedule> <horizon> <seed> <n>"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "robot-gym_b200")); sys.path.insert(0, REPO)
import numpy as np, torch
from robot_gym import cuda as rg
from robot_gym.model.robots.descriptions import GHOST, with_gait
from robot_gym.util import synthetic
from oracle import convex_mpc as cm

sched, horizon, seed, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
desc = with_gait(GHOST, sched); ctrl = desc.GetCtrlConstants()
st = synthetic.make_states(n, desc, schedule_ctrl=ctrl, seed=seed)
p = rg.default_mpc_params(ctrl.MPC_BODY_MASS, ctrl.MPC_BODY_INERTIA, ctrl.MPC_BODY_HEIGHT, horizon)
ws = rg.MpcWorkspace(p)
t = lambda a: torch.from_numpy(a).cuda()
f, hf, info = rg.mpc_build_solve(ws, t(st.com_velocity_body), t(st.base_rpy), t(st.base_rpy_rate), t(st.planned_contacts), t(st.foot_positions_base), t(st.command), want_horizon=True)
info = info.cpu().numpy(); hf = hf.cpu().numpy()
mp = cm.MpcParams(horizon=horizon)
for i in np.flatnonzero((info[:, 2] & 1) == 0):
    o, oinfo = cm.compute_contact_forces(mp, st.com_velocity_body[i].astype(np.float64), st.base_rpy[i].astype(np.float64), st.base_rpy_rate[i].astype(np.float64),
                                         st.planned_contacts[i], st.foot_positions_base[i].astype(np.float64), [0, 0, ctrl.MPC_BODY_HEIGHT],
                                         [st.command[i, 0], st.command[i, 1], 0.0], [0, 0, 0], [0, 0, float(st.command[i, 2])], return_info=True)
    g = hf[i].reshape(-1)
    print(f"env {i} info {info[i]}: GPU vs numpy oracle (polished={oinfo.get('polished')}, ipm iters {oinfo['iters']}) rel err "
          f"first step {np.abs(g[:12]-o[:12]).max()/max(1,np.abs(o[:12]).max()):.2e} horizon {np.abs(g-o).max()/max(1,np.abs(o).max()):.2e}")
