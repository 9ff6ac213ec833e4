"""A/B helper (GPU): solves/s at a given horizon for trot / bound.  python tools/perf_h.py <horizon> [n_env]"""
import os, sys, statistics
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "robot-gym_b200")); sys.path.insert(0, REPO)
import numpy as np, torch
from robot_gym import cuda as rg
from robot_gym.model.robots.descriptions import GHOST, with_gait
from robot_gym.util import synthetic
h = int(sys.argv[1]); n = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
for gait in ("trot", "bound"):
    desc = with_gait(GHOST, gait); ctrl = desc.GetCtrlConstants()
    p = rg.default_mpc_params(ctrl.MPC_BODY_MASS, ctrl.MPC_BODY_INERTIA, ctrl.MPC_BODY_HEIGHT, h)
    ws = rg.MpcWorkspace(p)
    st = synthetic.make_states(n, desc, schedule_ctrl=ctrl)
    t = lambda a: torch.from_numpy(a).cuda()
    args = (t(st.com_velocity_body), t(st.base_rpy), t(st.base_rpy_rate), t(st.planned_contacts), t(st.foot_positions_base), t(st.command))
    f = torch.empty((n, 12), dtype=torch.float32, device="cuda"); info = torch.empty((n, 4), dtype=torch.int32, device="cuda")
    for _ in range(3): rg.mpc_build_solve(ws, *args, contact_forces=f, solve_info=info)
    torch.cuda.synchronize(); ms = []
    for _ in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); rg.mpc_build_solve(ws, *args, contact_forces=f, solve_info=info); b.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
    med = statistics.median(ms)
    print(f"h={h} {gait:6s} n={n}: {med:8.3f} ms {n/med*1e3:12,.0f} solves/s polished {float(((info[:,2]&1)!=0).float().mean()):.5f}")
