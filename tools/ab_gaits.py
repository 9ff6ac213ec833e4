"""A/B helper (GPU): solves/s, iteration statistics and C-oracle agreement per gait schedule for the library
selected by RG_CUDA_LIB.   python tools/ab_gaits.py [n_env]"""
import os, sys, statistics
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "robot-gym_b200")); sys.path.insert(0, REPO)
import numpy as np, torch
from robot_gym import cuda as rg
from robot_gym.model.robots.descriptions import GHOST, with_gait
from robot_gym.util import synthetic
from oracle import c_oracle, convex_mpc

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    for gait in ("trot", "pace", "bound", "walk"):
        desc = GHOST if gait == "trot" else with_gait(GHOST, gait)
        ctrl = desc.GetCtrlConstants()
        p = rg.default_mpc_params(ctrl.MPC_BODY_MASS, ctrl.MPC_BODY_INERTIA, ctrl.MPC_BODY_HEIGHT, 10)
        for kv in filter(None, os.environ.get("RG_PERF_PARAMS", "").split(",")):
            k, v = kv.split("="); setattr(p, k, type(getattr(p, k))(float(v)))
        ws = rg.MpcWorkspace(p)
        st = synthetic.make_states(n, desc)
        t = lambda a: torch.from_numpy(a).cuda()
        args = (t(st.com_velocity_body), t(st.base_rpy), t(st.base_rpy_rate), t(st.planned_contacts), t(st.foot_positions_base), t(st.command))
        f = torch.empty((n, 12), dtype=torch.float32, device="cuda"); info = torch.empty((n, 4), dtype=torch.int32, device="cuda")
        for _ in range(3): rg.mpc_build_solve(ws, *args, contact_forces=f, solve_info=info)
        torch.cuda.synchronize()
        ms = []
        for _ in range(7):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); rg.mpc_build_solve(ws, *args, contact_forces=f, solve_info=info); b.record(); torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        med = statistics.median(ms)
        inf = info.cpu().numpy(); fo = f.cpu().numpy()
        m = 256
        ref, _, _ = c_oracle.solve_batch(convex_mpc.MpcParams(horizon=10), st.slice(0, m), ctrl.MPC_BODY_HEIGHT, n_threads=os.cpu_count())
        err = (np.abs(fo[:m] - ref).max(axis=1) / np.maximum(1, np.abs(ref).max(axis=1))).max()
        hard = inf[:, 0] > 0
        print(f"{gait:6s} n={n}: {med:7.3f} ms {n/med*1e3:11,.0f} solves/s | cold {np.mean((inf[:,2]&16)!=0):.3f} | ipm iters (when used) mean {inf[hard,0].mean() if hard.any() else 0:.2f} max {inf[:,0].max()} | rounds mean {inf[:,1].mean():.2f} max {inf[:,1].max()} | polished {np.mean((inf[:,2]&1)!=0):.5f} | err {err:.1e}")

if __name__ == "__main__":
    main()
