"""Control-step latency (CUDA events, p50) over batch sizes, cold and warm-started: python tools/gpu/lat_step.py [n ...]"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(REPO, "robot-gym_b200")); sys.path.insert(0, REPO)
import torch
from robot_gym.controllers.mpc.batched_mpc_controller import BatchedMPCController
from robot_gym.model.robots.descriptions import GHOST
from robot_gym.model.robots.synthetic_robot import SyntheticRobotBatch
from robot_gym.util import synthetic

for n in [int(a) for a in sys.argv[1:]] or [1, 64, 1024, 4096, 65536]:
    for warm_start in (False, True):
        st = synthetic.make_states(n, GHOST, seed=1)
        robot = SyntheticRobotBatch(GHOST, st)
        ctl = BatchedMPCController(robot, robot.GetTimeSinceReset, squeeze_single=False, warm_start=warm_start)
        ctl.command.copy_(torch.from_numpy(st.command).cuda())
        reps = 300 if n <= 4096 else 40
        for _ in range(20): ctl.step()
        torch.cuda.synchronize()
        ms = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ctl.step(); b.record(); torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        ms.sort()
        print(f"n={n} warm_start={warm_start}: p50 {ms[len(ms)//2]*1e3:.1f} us  p99 {ms[int(0.99*len(ms))]*1e3:.1f} us")
