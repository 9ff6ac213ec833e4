"""Small workload for compute-sanitizer (racecheck / memcheck): MPC solve at h = 5/10/20 + a control step."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "robot-gym_b200")); sys.path.insert(0, REPO)
import numpy as np, torch
from robot_gym import cuda as rg
from robot_gym.controllers.mpc.batched_mpc_controller import BatchedMPCController
from robot_gym.model.robots.descriptions import GHOST, with_gait
from robot_gym.model.robots.synthetic_robot import SyntheticRobotBatch
from robot_gym.util import synthetic
ctrl = GHOST.GetCtrlConstants()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
for h in (10, 5, 20):
    p = rg.default_mpc_params(ctrl.MPC_BODY_MASS, ctrl.MPC_BODY_INERTIA, ctrl.MPC_BODY_HEIGHT, h)
    ws = rg.MpcWorkspace(p)
    st = synthetic.make_states(n, GHOST, seed=h)
    t = lambda a: torch.from_numpy(a).cuda()
    f, hf, info = rg.mpc_build_solve(ws, t(st.com_velocity_body), t(st.base_rpy), t(st.base_rpy_rate), t(st.planned_contacts), t(st.foot_positions_base), t(st.command), want_horizon=True)
    torch.cuda.synchronize()
    print("h", h, "ok", np.isfinite(f.cpu().numpy()).all(), info.cpu().numpy()[:, 2].min())
# a bound schedule sends ~15 % of the envs through the interior-point fallback (warm start, escalation ladder)
pace = with_gait(GHOST, "bound")
p = rg.default_mpc_params(ctrl.MPC_BODY_MASS, ctrl.MPC_BODY_INERTIA, ctrl.MPC_BODY_HEIGHT, 10)
ws = rg.MpcWorkspace(p)
st = synthetic.make_states(n, pace, seed=5)
t = lambda a: torch.from_numpy(a).cuda()
f, hf, info = rg.mpc_build_solve(ws, t(st.com_velocity_body), t(st.base_rpy), t(st.base_rpy_rate), t(st.planned_contacts), t(st.foot_positions_base), t(st.command))
torch.cuda.synchronize()
inf = info.cpu().numpy()
print("bound ok", np.isfinite(f.cpu().numpy()).all(), "interior-point envs", int((inf[:, 0] > 0).sum()), "of", n)
# the two-kernel path (lean kernel + fallback queue) needs more envs than fit one wave
big = int(os.environ.get("RG_SANITIZE_BIG", "1600"))
ws = rg.MpcWorkspace(p, max_envs=big)
st = synthetic.make_states(big, pace, seed=6)
f, hf, info = rg.mpc_build_solve(ws, t(st.com_velocity_body), t(st.base_rpy), t(st.base_rpy_rate), t(st.planned_contacts), t(st.foot_positions_base), t(st.command))
torch.cuda.synchronize()
inf = info.cpu().numpy()
print("two-kernel bound ok", np.isfinite(f.cpu().numpy()).all(), "queued envs", int(((inf[:, 2] & 16) == 0).sum()), "of", big)
st = synthetic.make_states(n, GHOST, seed=3)
robot = SyntheticRobotBatch(GHOST, st)
ctl = BatchedMPCController(robot, robot.GetTimeSinceReset)
for _ in range(2): ctl.step()
torch.cuda.synchronize()
print("control step ok")
