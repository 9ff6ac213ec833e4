"""GPU perf + correctness spot check of rg_mpc_build_solve (development helper).

    python tools/perf_mpc.py [n_env ...]      (default 4096 65536)
Prints solves/s (median of 10 event-timed launches) and the worst relative error vs the C oracle
on the first 512 envs."""
import os, sys, statistics
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "robot-gym_b200")); sys.path.insert(0, REPO)
import numpy as np, torch
from robot_gym import cuda as rg
from robot_gym.model.robots.descriptions import GHOST, with_gait
from robot_gym.util import synthetic
from oracle import c_oracle, convex_mpc

def main():
    sizes = [int(a) for a in sys.argv[1:]] or [4096, 65536]
    ctrl = GHOST.GetCtrlConstants()
    gait = os.environ.get("RG_PERF_GAIT", "trot")
    desc = GHOST if gait == "trot" else with_gait(GHOST, gait)
    for horizon in [int(h) for h in os.environ.get("RG_PERF_H", "10").split(",")]:
        p = rg.default_mpc_params(ctrl.MPC_BODY_MASS, ctrl.MPC_BODY_INERTIA, ctrl.MPC_BODY_HEIGHT, horizon)
        for kv in filter(None, os.environ.get("RG_PERF_PARAMS", "").split(",")):     # e.g. ipm_tol=1e-4,cold_start_rounds=0
            k, v = kv.split("=")
            setattr(p, k, type(getattr(p, k))(float(v)))
        ws = rg.MpcWorkspace(p, max_envs=max(sizes))
        for n in sizes:
            for all_stance in (False, True)[:1 if os.environ.get("RG_PERF_NO_ALLSTANCE") else 2]:
                st = synthetic.make_states(n, desc, schedule_ctrl=desc.GetCtrlConstants(), all_stance=all_stance)
                t = lambda a: torch.from_numpy(a).cuda()
                args = (t(st.com_velocity_body), t(st.base_rpy), t(st.base_rpy_rate), t(st.planned_contacts), t(st.foot_positions_base), t(st.command))
                f = torch.empty((n, 12), dtype=torch.float32, device="cuda"); info = torch.empty((n, 4), dtype=torch.int32, device="cuda")
                for _ in range(3): rg.mpc_build_solve(ws, *args, contact_forces=f, solve_info=info)
                torch.cuda.synchronize()
                ms = []
                for _ in range(10):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(); rg.mpc_build_solve(ws, *args, contact_forces=f, solve_info=info); b.record(); torch.cuda.synchronize()
                    ms.append(a.elapsed_time(b))
                med = statistics.median(ms)
                inf = info.cpu().numpy(); fo = f.cpu().numpy()
                m = min(n, 512)
                ref, _, _ = c_oracle.solve_batch(convex_mpc.MpcParams(horizon=horizon), st.slice(0, m), ctrl.MPC_BODY_HEIGHT, n_threads=os.cpu_count())
                err = (np.abs(fo[:m] - ref).max(axis=1) / np.maximum(1, np.abs(ref).max(axis=1))).max()
                print(f"{gait} h={horizon} n={n} all_stance={all_stance}: {med:.3f} ms  {n/med*1e3:,.0f} solves/s | iters {inf[:,0].mean():.2f} polish {inf[:,1].mean():.2f} "
                      f"polished {np.mean((inf[:,2]&1)!=0):.4f} cold {np.mean((inf[:,2]&16)!=0):.3f} | worst rel err vs C oracle {err:.2e}")

if __name__ == "__main__":
    main()
