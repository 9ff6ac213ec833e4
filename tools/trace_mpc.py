"""Single-env phase timing of mpc_solve_kernel (development helper, GPU).

Builds a -DRG_DEBUG_TRACE copy of the library into ab/ and prints the cycles thread 0 of one env
spends per phase, solo (1 env) and under load (the same env inside a 65536-env launch).
    python tools/trace_mpc.py [env_index ...]
"""
import os, sys, ctypes
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.makedirs(os.path.join(REPO, "ab"), exist_ok=True)
os.environ["RG_CUDA_LIB"] = os.path.join(REPO, "ab", "librg_trace.so")
os.environ["RG_DEBUG_TRACE"] = "1"
sys.path.insert(0, os.path.join(REPO, "robot-gym_b200")); sys.path.insert(0, REPO)
import numpy as np, torch
from robot_gym import cuda as rg
from robot_gym.model.robots.descriptions import GHOST, with_gait
from robot_gym.util import synthetic

PHASES = {8: "setup (inputs, tables, K^-1)", 40: "  setup: input loads, sincos", 41: "  setup: K2, lever arms, I_w^-1", 42: "  setup: Q_t inverses",
          43: "  setup: K^-1 angular (rank-h sums)", 44: "  setup: g~ + barrier", 0: "ipm: loop head", 1: "ipm: apply_p + residual", 2: "ipm: block parts",
          3: "ipm: factor_psi", 4: "ipm: rhs", 5: "ipm: woodbury", 6: "ipm: tail/step", 20: "as: basis + gap",
          12: "factor: n-blocks", 10: "factor: psi build", 11: "factor: cholesky", 21: "as: rhs", 22: "as: woodbury solve",
          23: "as: apply_p", 24: "as: verify + reduce", 7: "exit", 9: "TOTAL",
          30: "chol: phase 1 (last row)", 33: "chol: barrier 1", 31: "chol: phase 2 (last row)", 32: "chol: barrier 2 + loop"}


def main():
    envs = [int(a) for a in sys.argv[1:]] or [0, 1, 2, 3]
    find_unpolished = os.environ.get("RG_TRACE_FIND_UNPOLISHED") == "1"
    lib = rg.load()
    lib.rg_debug_set_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
    ctrl = GHOST.GetCtrlConstants()
    p = rg.default_mpc_params(ctrl.MPC_BODY_MASS, ctrl.MPC_BODY_INERTIA, ctrl.MPC_BODY_HEIGHT, int(os.environ.get("RG_TRACE_H", "10")))
    ws = rg.MpcWorkspace(p)
    n = int(os.environ.get("RG_TRACE_N", "65536"))     # batch the env index refers to (make_states(n) draws depend on n)
    gait = os.environ.get("RG_TRACE_GAIT", "trot")
    desc = GHOST if gait == "trot" else with_gait(GHOST, gait)
    st = synthetic.make_states(n, desc, schedule_ctrl=desc.GetCtrlConstants(), seed=int(os.environ.get("RG_TRACE_SEED", str(synthetic.SEED))))
    t = lambda a: torch.from_numpy(a).cuda()
    full = (t(st.com_velocity_body), t(st.base_rpy), t(st.base_rpy_rate), t(st.planned_contacts), t(st.foot_positions_base), t(st.command))
    trace = torch.zeros(1024, dtype=torch.float64, device="cuda")
    if find_unpolished:
        _, _, info_all = rg.mpc_build_solve(ws, *full)
        ia = info_all.cpu().numpy()
        envs = list(np.flatnonzero((ia[:, 2] & 1) == 0)[:3]) + (list(np.argsort(-ia[:, 0])[:2]) if os.environ.get("RG_TRACE_SLOWEST") else [])
        print("unpolished / slowest envs:", envs, ia[envs])
    for e in envs:
        for label, lo, cnt, idx in (("solo", e, 1, 0), ("loaded", 0, n, e))[:1 if os.environ.get("RG_TRACE_SOLO") else 2]:
            args = tuple(a[lo:lo + cnt].contiguous() for a in full)
            lib.rg_debug_set_trace(None, -1)
            for _ in range(2): f, _, info = rg.mpc_build_solve(ws, *args)
            trace.zero_(); torch.cuda.synchronize()
            lib.rg_debug_set_trace(trace.data_ptr(), idx)
            f, _, info = rg.mpc_build_solve(ws, *args)
            torch.cuda.synchronize()
            lib.rg_debug_set_trace(None, -1)
            tr = trace.cpu().numpy(); inf = info.cpu().numpy()[idx]
            tot = tr[909]
            print(f"env {e} [{label}] iters {inf[0]} rounds {inf[1]} status {inf[2]} nact {inf[3]}: total {tot:.0f} cyc = {tot/1.965e3:.1f} us")
            for r in range(64):
                if tr[4 * r] == 0 and tr[4 * r + 1] == 0 and r > 0: break
                kind = {-1.0: "as round (after ipm)", -2.0: "as round (cold)"}.get(tr[4 * r], "ipm iter")
                print(f"    trace[{r}] {kind}: {tr[4*r]:.3e} {tr[4*r+1]:.3e} {tr[4*r+2]:.3e} {tr[4*r+3]:.3e}")
            for k, name in PHASES.items():
                if tr[900 + k] > 0 and k != 9: print(f"    {name:32s} {tr[900+k]:9.0f} cyc  {100*tr[900+k]/tot:5.1f} %")


if __name__ == "__main__":
    main()
