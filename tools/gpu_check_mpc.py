"""Quick GPU bring-up check: rg_mpc_build_solve vs the numpy oracle on seeded synthetic states."""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "robot-gym_b200")); sys.path.insert(0, REPO)
import numpy as np, torch
from robot_gym import cuda as rg
from robot_gym.model.robots.descriptions import GHOST
from robot_gym.util import synthetic
from oracle import convex_mpc as cm

def run(n, horizon, all_stance, n_check, weights=None):
    ctrl = GHOST.GetCtrlConstants()
    p = rg.default_mpc_params(ctrl.MPC_BODY_MASS, ctrl.MPC_BODY_INERTIA, ctrl.MPC_BODY_HEIGHT, horizon)
    if weights is not None:
        for i, w in enumerate(weights): p.weights[i] = w
    ws = rg.MpcWorkspace(p)
    st = synthetic.make_states(n, GHOST, all_stance=all_stance)
    dev = "cuda"
    t = lambda a: torch.from_numpy(a).to(dev)
    args = (t(st.com_velocity_body), t(st.base_rpy), t(st.base_rpy_rate), t(st.planned_contacts), t(st.foot_positions_base), t(st.command))
    f, hf, info = rg.mpc_build_solve(ws, *args, want_horizon=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f, hf, info = rg.mpc_build_solve(ws, *args, want_horizon=True); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    f = f.cpu().numpy(); info = info.cpu().numpy(); hf = hf.cpu().numpy()
    print(f"n={n} h={horizon} all_stance={all_stance}: {ms:.3f} ms -> {n/ms*1e3:.0f} solves/s; iters mean {info[:,0].mean():.2f} max {info[:,0].max()}, "
          f"polish mean {info[:,1].mean():.2f} max {info[:,1].max()}, status hist {np.unique(info[:,2], return_counts=True)}, nan {np.isnan(f).sum()}")
    mp = cm.MpcParams(horizon=horizon)
    if weights is not None: mp.weights = tuple(weights)
    worst = 0.0; worst_h = 0.0
    t0 = time.time()
    for i in range(n_check):
        ref = cm.compute_contact_forces(mp, st.com_velocity_body[i].astype(np.float64), st.base_rpy[i].astype(np.float64),
            st.base_rpy_rate[i].astype(np.float64), st.planned_contacts[i], st.foot_positions_base[i].astype(np.float64),
            [0, 0, ctrl.MPC_BODY_HEIGHT], [st.command[i,0], st.command[i,1], 0.0], [0,0,0], [0,0,float(st.command[i,2])])
        scale = max(1.0, np.abs(ref[:12]).max())
        err = np.abs(f[i] - ref[:12]).max() / scale
        errh = np.abs(hf[i].reshape(-1) - ref).max() / max(1.0, np.abs(ref).max())
        if err > 1e-5: print("  env", i, "contacts", st.planned_contacts[i], "err", err, "info", info[i], "\n   gpu", f[i], "\n   ref", ref[:12])
        worst = max(worst, err); worst_h = max(worst_h, errh)
    print(f"  checked {n_check} envs vs oracle in {time.time()-t0:.1f}s: worst rel err first-step {worst:.3e}, full horizon {worst_h:.3e}")

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), rg.load().rg_version().decode())
    run(512, 10, False, 64)
    run(512, 10, True, 48)
    run(256, 5, False, 32)
    run(256, 20, False, 16)
    run(256, 10, False, 32, weights=(5,5,0.2,0,0,10,0.,0.,1.,1.,1.,0.,0))
    run(4096, 10, False, 0)
    run(4096, 10, True, 0)
    run(65536, 10, False, 0)
