"""Large robustness sweep (GPU): unverified solves / numeric flags / worst KKT violation over many seeds,
schedules and horizons.  Every solve is checked on the GPU output itself: swing forces exactly zero, fz and
friction-pyramid bounds within 1e-6 * fz_max, status bits.  python tools/robustness_big.py [n_env] [n_seeds]"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "robot-gym_b200")); sys.path.insert(0, REPO)
import numpy as np, torch
from robot_gym import cuda as rg
from robot_gym.model.robots.descriptions import GHOST, with_gait
from robot_gym.util import synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 3
tot = dict(envs=0, unverified=0, numeric=0)
for sched, horizon in [("trot", 10), ("pace", 10), ("bound", 10), ("walk", 10), ("stand", 10), ("trot", 5), ("bound", 5), ("trot", 20), ("walk", 20)]:
    desc = with_gait(GHOST, sched); ctrl = desc.GetCtrlConstants()
    p = rg.default_mpc_params(ctrl.MPC_BODY_MASS, ctrl.MPC_BODY_INERTIA, ctrl.MPC_BODY_HEIGHT, horizon)
    ws = rg.MpcWorkspace(p)
    m = n if horizon < 20 else n // 4
    for seed in range(100, 100 + seeds):
        st = synthetic.make_states(m, desc, schedule_ctrl=ctrl, seed=seed)
        t = lambda a: torch.from_numpy(a).cuda()
        f, hf, info = rg.mpc_build_solve(ws, t(st.com_velocity_body), t(st.base_rpy), t(st.base_rpy_rate), t(st.planned_contacts),
                                         t(st.foot_positions_base), t(st.command), want_horizon=True)
        status = info[:, 2]
        ok = ((status & 1) != 0)
        contacts = t(st.planned_contacts).bool()                                  # [m,4]
        fr = -hf.view(m, horizon, 4, 3).double()                                  # QP variables (solver returns the negation)
        swing_max = fr[~contacts[:, None, :].expand(m, horizon, 4)].abs().max().item() if (~contacts).any() else 0.0
        fz = fr[..., 2]; mu = 0.45
        stance = contacts[:, None, :].expand(m, horizon, 4)
        viol = torch.zeros((), dtype=torch.float64, device="cuda")
        viol = torch.maximum(viol, (p.fz_min - fz)[stance].max()); viol = torch.maximum(viol, (fz - p.fz_max)[stance].max())
        for comp in (0, 1):
            viol = torch.maximum(viol, (fr[..., comp].abs() - mu * fz)[stance].max())
        torch.cuda.synchronize()
        unv = int((~ok).sum()); num = int(((status & 8) != 0).sum())
        tot["envs"] += m; tot["unverified"] += unv; tot["numeric"] += num
        # warm-started re-solve of a neighbouring problem (the next control step): seeds from this solve
        seedbuf = rg.new_active_set(m, horizon)
        rg.mpc_build_solve(ws, t(st.com_velocity_body), t(st.base_rpy), t(st.base_rpy_rate), t(st.planned_contacts),
                           t(st.foot_positions_base), t(st.command), active_set=seedbuf)
        rng = np.random.default_rng(seed)
        v2 = (st.com_velocity_body + rng.normal(0, 0.02, st.com_velocity_body.shape)).astype(np.float32)
        rpy2 = (st.base_rpy + rng.normal(0, 0.005, st.base_rpy.shape) * np.array([1, 1, 0])).astype(np.float32)
        fw, _, infow = rg.mpc_build_solve(ws, t(v2), t(rpy2), t(st.base_rpy_rate), t(st.planned_contacts), t(st.foot_positions_base), t(st.command), active_set=seedbuf)
        fc, _, infoc = rg.mpc_build_solve(ws, t(v2), t(rpy2), t(st.base_rpy_rate), t(st.planned_contacts), t(st.foot_positions_base), t(st.command))
        torch.cuda.synchronize()
        warm_unv = int(((infow[:, 2] & 1) == 0).sum())
        warm_gap = ((fw - fc).abs().max(dim=1).values / fc.abs().max(dim=1).values.clamp(min=1.0)).max().item()
        tot["unverified"] += warm_unv
        tot.setdefault("warm_gap_max", 0.0); tot["warm_gap_max"] = max(tot["warm_gap_max"], warm_gap)
        tot.setdefault("warm_rounds", []).append((float(infow[:, 1].float().mean()), float(infoc[:, 1].float().mean())))
        print(f"{sched:6s} h={horizon:2d} seed={seed} n={m}: unverified {unv} (warm {warm_unv}, warm-vs-cold gap {warm_gap:.1e}, rounds warm {infow[:,1].float().mean():.2f} / cold {infoc[:,1].float().mean():.2f}) numeric {num} | swing |f| max {swing_max:.1e} | worst bound/cone violation {viol.item() / p.fz_max:.1e} x fz_max"
              f" | ipm iters max {int(info[:,0].max())} rounds max {int(info[:,1].max())}")
wr = tot.pop("warm_rounds"); print("TOTAL", tot, "mean rounds warm/cold", np.mean([a for a, b in wr]).round(2), np.mean([b for a, b in wr]).round(2))
