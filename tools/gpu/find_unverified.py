"""Which envs of the sharded 65536-env trot batch end unverified, and does the single-kernel route verify them? (GPU helper)"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(REPO, "robot-gym_b200")); sys.path.insert(0, REPO)
import numpy as np, torch
from robot_gym import cuda as rg
from robot_gym.model.robots.descriptions import GHOST, with_gait
from robot_gym.util import synthetic
ctrl = GHOST.GetCtrlConstants()
for gait in ("trot", "pace"):
    desc = with_gait(GHOST, gait)
    st = synthetic.make_states_sharded(0, 65536, desc, schedule_ctrl=desc.GetCtrlConstants())
    t = lambda a: torch.from_numpy(a).cuda()
    ins = [t(getattr(st, k)) for k in ("com_velocity_body", "base_rpy", "base_rpy_rate", "planned_contacts", "foot_positions_base", "command")]
    for two in (1, 0):
        p = rg.default_mpc_params(ctrl.MPC_BODY_MASS, ctrl.MPC_BODY_INERTIA, ctrl.MPC_BODY_HEIGHT, 10)
        p.two_kernel_solve = two
        ws = rg.MpcWorkspace(p, max_envs=65536)
        f, _, info = rg.mpc_build_solve(ws, *ins)
        torch.cuda.synchronize()
        i = info.cpu().numpy()
        bad = np.flatnonzero((i[:, 2] & 5) == 0)
        print(gait, "two_kernel", two, "unverified envs:", bad, i[bad])
        if two == 1: f_two = f.clone()
        else:
            for e in bad_two:
                print("   env", e, "single-kernel info", i[e], "force gap", float((f[e] - f_two[e]).abs().max()))
        if two == 1: bad_two = bad
