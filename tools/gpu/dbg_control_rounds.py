"""Why does the cold control step need more active-set rounds than the plain solve bench? (development helper, GPU)"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(REPO, "robot-gym_b200")); sys.path.insert(0, REPO)
import numpy as np, torch
from robot_gym import cuda as rg
from robot_gym.controllers.mpc.batched_mpc_controller import BatchedMPCController
from robot_gym.model.robots.descriptions import GHOST
from robot_gym.model.robots.synthetic_robot import SyntheticRobotBatch
from robot_gym.util import synthetic

n = 4096
st = synthetic.make_states_sharded(0, n, GHOST)
robot = SyntheticRobotBatch(GHOST, st, device="cuda")
ctl = BatchedMPCController(robot, robot.GetTimeSinceReset, warm_start=False)
ctl.command.copy_(torch.from_numpy(st.command).cuda())
for k in range(25):
    ctl.step()
    if k in (0, 1, 2, 5, 19, 20, 24):
        torch.cuda.synchronize()
        info = ctl.solve_info.cpu().numpy()
        print(f"step {k}: rounds mean {info[:,1].mean():.3f} hist {np.bincount(info[:,1], minlength=8)[:10]} ipm>0 {np.mean(info[:,0]>0):.4f} "
              f"|v_est| {ctl.com_velocity_body.abs().mean().item():.3f} n_stance hist {np.bincount(ctl.mpc_contact_state.sum(dim=1).cpu().numpy(), minlength=5)}")
# the same QPs through the plain entry point
ctrl = GHOST.GetCtrlConstants()
p = rg.default_mpc_params(ctrl.MPC_BODY_MASS, ctrl.MPC_BODY_INERTIA, ctrl.MPC_BODY_HEIGHT, 10)
ws = rg.MpcWorkspace(p, max_envs=n)
rpy = robot.base_rpy.clone(); rpy[:, 2] = 0
f, _, info = rg.mpc_build_solve(ws, ctl.com_velocity_body, rpy, robot.base_rpy_rate, ctl.mpc_contact_state, robot.foot_positions_base.view(n, 12), ctl.command)
torch.cuda.synchronize()
i = info.cpu().numpy()
print("plain solve on the controller's inputs: rounds mean", i[:, 1].mean(), "forces equal:", torch.equal(f, ctl.contact_forces))
f2, _, info2 = rg.mpc_build_solve(ws, *[torch.from_numpy(getattr(st, k)).cuda() for k in ("com_velocity_body", "base_rpy", "base_rpy_rate", "planned_contacts", "foot_positions_base", "command")])
torch.cuda.synchronize()
i2 = info2.cpu().numpy()
print("plain solve on the bench inputs: rounds mean", i2[:, 1].mean(), "hist", np.bincount(i2[:, 1], minlength=8)[:10])
two = st.planned_contacts.sum(axis=1) == 2
for name, pair in (("FL+RR", (0, 1, 1, 0)), ("FR+RL", (1, 0, 0, 1))):
    m = np.all(st.planned_contacts == np.array(pair, dtype=np.uint8), axis=1)
    print(f"   bench envs with stance {name}: {m.sum()} rounds mean {i2[m, 1].mean():.3f}")
