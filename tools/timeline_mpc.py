"""Per-CTA timeline of one mpc_solve_kernel launch (development helper, GPU, -DRG_DEBUG_TRACE -DRG_DEBUG_TIMELINE_ONLY build in ab/).
Prints when the waves start, how long cold-start / interior-point envs take under load and what the tail is."""
import os, sys, ctypes
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.makedirs(os.path.join(REPO, "ab"), exist_ok=True)
os.environ["RG_CUDA_LIB"] = os.path.join(REPO, "ab", "librg_timeline.so")
os.environ["RG_DEBUG_TRACE"] = "2"      # start / end stamps only: the phase timers of mode 1 slow the kernel by a third
sys.path.insert(0, os.path.join(REPO, "robot-gym_b200")); sys.path.insert(0, REPO)
import numpy as np, torch
from robot_gym import cuda as rg
from robot_gym.cuda import build as _build
_build.build()                           # an RG_CUDA_LIB override is loaded as it is: rebuild the traced library when the sources moved
from robot_gym.model.robots.descriptions import GHOST
from robot_gym.util import synthetic


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    lib = rg.load()
    lib.rg_debug_set_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
    ctrl = GHOST.GetCtrlConstants()
    p = rg.default_mpc_params(ctrl.MPC_BODY_MASS, ctrl.MPC_BODY_INERTIA, ctrl.MPC_BODY_HEIGHT, 10)
    ws = rg.MpcWorkspace(p)
    st = synthetic.make_states(n, GHOST)
    t = lambda a: torch.from_numpy(a).cuda()
    args = (t(st.com_velocity_body), t(st.base_rpy), t(st.base_rpy_rate), t(st.planned_contacts), t(st.foot_positions_base), t(st.command))
    trace = torch.zeros(1024 + 4 * n, dtype=torch.float64, device="cuda")
    for _ in range(3): rg.mpc_build_solve(ws, *args)
    torch.cuda.synchronize()
    lib.rg_debug_set_trace(trace.data_ptr(), -2)
    f, _, info = rg.mpc_build_solve(ws, *args)
    torch.cuda.synchronize()
    lib.rg_debug_set_trace(None, -1)
    tr = trace.cpu().numpy()[1024:].reshape(n, 4); inf = info.cpu().numpy()
    t0 = tr[:, 0].min()
    start, end = (tr[:, 0] - t0) * 1e-3, (tr[:, 1] - t0) * 1e-3      # us
    dur = end - start
    cold = (inf[:, 2] & 16) != 0
    print(f"n={n}: kernel span {end.max():.1f} us; CTA durations: cold mean {dur[cold].mean():.1f} p50 {np.median(dur[cold]):.1f} p99 {np.percentile(dur[cold], 99):.1f} max {dur[cold].max():.1f} us"
          + (f" | interior-point envs ({(~cold).sum()}): mean {dur[~cold].mean():.1f} max {dur[~cold].max():.1f} us" if (~cold).any() else " | no interior-point envs"))
    order = np.argsort(end)[::-1][:12]
    print("   last CTAs to finish: (env, start us, end us, ipm iters, rounds, status)")
    for e in order: print(f"     {e:5d} {start[e]:8.1f} {end[e]:8.1f}  {inf[e,0]:2d} {inf[e,1]:2d} {inf[e,2]:3d}")
    for q in (0.25, 0.5, 0.75, 0.9, 0.99): print(f"   {int(q*100)} % of envs done by {np.quantile(end, q):.1f} us; last start {start.max():.1f} us")
    rounds = inf[:, 1]
    if os.environ.get("RG_TIMELINE_SAVE"):      # raw per-env stamps for offline scheduling experiments
        np.savez(os.environ["RG_TIMELINE_SAVE"], start=start, end=end, rounds=rounds, smid=tr[:, 2], n_stance=(st.planned_contacts != 0).sum(axis=1))
    for r in range(1, 16):
        sel = cold & (rounds == r)
        if sel.any(): print(f"   cold, {r} rounds: {sel.sum():5d} envs, mean duration {dur[sel].mean():.1f} us")


if __name__ == "__main__":
    main()
