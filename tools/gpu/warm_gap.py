"""Warm-vs-cold agreement of verified solves on one (gait, horizon, seed) batch; prints the worst envs (GPU)."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(REPO, "robot-gym_b200")); sys.path.insert(0, REPO)
import numpy as np, torch
from robot_gym import cuda as rg
from robot_gym.model.robots.descriptions import GHOST, with_gait
from robot_gym.util import synthetic
from oracle import c_oracle, convex_mpc
sched, horizon, seed, m = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]) if len(sys.argv) > 4 else 65536
desc = with_gait(GHOST, sched); ctrl = desc.GetCtrlConstants()
p = rg.default_mpc_params(ctrl.MPC_BODY_MASS, ctrl.MPC_BODY_INERTIA, ctrl.MPC_BODY_HEIGHT, horizon)
ws = rg.MpcWorkspace(p)
st = synthetic.make_states(m, desc, schedule_ctrl=ctrl, seed=seed)
t = lambda a: torch.from_numpy(a).cuda()
seedbuf = rg.new_active_set(m, horizon)
rg.mpc_build_solve(ws, t(st.com_velocity_body), t(st.base_rpy), t(st.base_rpy_rate), t(st.planned_contacts), t(st.foot_positions_base), t(st.command), active_set=seedbuf)
rng = np.random.default_rng(seed)
v2 = (st.com_velocity_body + rng.normal(0, 0.02, st.com_velocity_body.shape)).astype(np.float32)
rpy2 = (st.base_rpy + rng.normal(0, 0.005, st.base_rpy.shape) * np.array([1, 1, 0])).astype(np.float32)
fw, _, infow = rg.mpc_build_solve(ws, t(v2), t(rpy2), t(st.base_rpy_rate), t(st.planned_contacts), t(st.foot_positions_base), t(st.command), active_set=seedbuf)
fc, _, infoc = rg.mpc_build_solve(ws, t(v2), t(rpy2), t(st.base_rpy_rate), t(st.planned_contacts), t(st.foot_positions_base), t(st.command))
gap = ((fw - fc).abs().max(dim=1).values / fc.abs().max(dim=1).values.clamp(min=1.0)).cpu().numpy()
worst = np.argsort(-gap)[:5]
print("lib", os.environ.get("RG_CUDA_LIB", "default"), "worst gaps", gap[worst], "envs", worst)
import dataclasses
st2 = dataclasses.replace(st, com_velocity_body=v2, base_rpy=rpy2) if dataclasses.is_dataclass(st) else st
for e in worst[:3]:
    sl = st2.slice(int(e), int(e) + 1)
    ref, _, _ = c_oracle.solve_batch(convex_mpc.MpcParams(horizon=horizon), sl, ctrl.MPC_BODY_HEIGHT, n_threads=1)
    sc = max(1.0, np.abs(ref).max())
    print(f"  env {e}: warm err vs C oracle {np.abs(fw[e].cpu().numpy() - ref[0]).max() / sc:.2e}  cold err {np.abs(fc[e].cpu().numpy() - ref[0]).max() / sc:.2e}  info warm {infow[e].tolist()} cold {infoc[e].tolist()}")
