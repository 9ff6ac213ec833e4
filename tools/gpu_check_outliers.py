"""GPU vs BOTH oracles on the envs where GPU and C oracle disagree most (development helper)."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "robot-gym_b200")); sys.path.insert(0, REPO)
import numpy as np, torch
from robot_gym import cuda as rg
from robot_gym.model.robots.descriptions import GHOST, with_gait
from robot_gym.util import synthetic
from oracle import convex_mpc as cm, c_oracle

sched, horizon, n, m = sys.argv[1] if len(sys.argv) > 1 else "bound", 10, 8192, 2048
desc = with_gait(GHOST, sched); ctrl = desc.GetCtrlConstants()
st = synthetic.make_states(n, desc, schedule_ctrl=ctrl, seed=77)
p = rg.default_mpc_params(ctrl.MPC_BODY_MASS, ctrl.MPC_BODY_INERTIA, ctrl.MPC_BODY_HEIGHT, horizon)
mp = cm.MpcParams(horizon=horizon)
ws = rg.MpcWorkspace(p)
t = lambda a: torch.from_numpy(a).cuda()
f, hf, info = rg.mpc_build_solve(ws, t(st.com_velocity_body), t(st.base_rpy), t(st.base_rpy_rate), t(st.planned_contacts), t(st.foot_positions_base), t(st.command), want_horizon=True)
info = info.cpu().numpy(); f = f.cpu().numpy(); hf = hf.cpu().numpy()
ref, _, _ = c_oracle.solve_batch(mp, st.slice(0, m), ctrl.MPC_BODY_HEIGHT, n_threads=os.cpu_count())
err = np.abs(f[:m] - ref).max(axis=1) / np.maximum(1, np.abs(ref).max(axis=1))
for i in np.argsort(err)[::-1][:6]:
    o, oinfo = cm.compute_contact_forces(mp, st.com_velocity_body[i].astype(np.float64), st.base_rpy[i].astype(np.float64), st.base_rpy_rate[i].astype(np.float64),
                                  st.planned_contacts[i], st.foot_positions_base[i].astype(np.float64), [0, 0, ctrl.MPC_BODY_HEIGHT],
                                  [st.command[i, 0], st.command[i, 1], 0.0], [0, 0, 0], [0, 0, float(st.command[i, 2])], return_info=True)
    sc = max(1, np.abs(o[:12]).max())
    qp = oinfo["qp"]
    obj = lambda x: 0.5 * x @ qp.p_mat @ x + qp.q_vec @ x
    xg = -hf[i].reshape(-1).astype(np.float64)
    cert = cm.kkt_certificate(qp.p_mat, qp.q_vec, qp.c_mat, qp.lb, qp.ub, xg)
    print(f"env {i}: info {info[i]} | GPU vs C oracle {err[i]:.2e} | GPU vs numpy oracle (polished={oinfo.get('polished')}) {np.abs(f[i]-o[:12]).max()/sc:.2e} | C vs numpy {np.abs(ref[i]-o[:12]).max()/sc:.2e}"
          f" | obj(GPU f32 horizon)-obj(numpy) {obj(xg)-obj(oinfo['x']):.3e} | GPU primal viol {cert['primal']:.2e}")
