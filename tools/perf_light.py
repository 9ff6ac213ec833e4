"""HBM-bound kernels either side of the solve against the HBM roofline (development helper, GPU).

    python tools/perf_light.py [n_env]                 (default 2^20)
Event-timed medians of rg_state_from_sim, rg_hybrid_motor_torque and of the control step's prologue / epilogue
kernels (timed through rg_gait_step + ... is not possible from outside: the two fused kernels are taken from the
ncu launch list of THIS script instead, `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
-k regex:"step_|state_from_sim|hybrid" python tools/perf_light.py`), with the algorithmic bytes per env of each.
"""
import json, os, statistics, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "robot-gym_b200")); sys.path.insert(0, REPO)
import numpy as np, torch
from robot_gym import cuda as rg
from robot_gym.controllers.mpc.batched_mpc_controller import BatchedMPCController
from robot_gym.model.robots.descriptions import GHOST
from robot_gym.model.robots.sim_state_robot import SimStateRobotBatch
from robot_gym.model.robots.synthetic_robot import SyntheticRobotBatch
from robot_gym.util import synthetic


def time_ms(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        out.append(a.elapsed_time(b))
    return statistics.median(out)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    peak = 6558.7
    try: peak = float(json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception: pass
    rng = np.random.default_rng(3); dev = "cuda"
    f32 = lambda *shape: torch.from_numpy(rng.uniform(-0.3, 0.3, shape).astype(np.float32)).to(dev)
    quat = torch.from_numpy(synthetic.euler_to_quat_xyzw(rng.uniform(-0.5, 0.5, (n, 3))).astype(np.float32)).to(dev)
    robot = SimStateRobotBatch(GHOST, n, device=dev)
    t = torch.zeros(n, dtype=torch.float64, device=dev)
    v, w, joints = f32(n, 3), f32(n, 3), f32(n, 12)
    contacts = torch.ones((n, 4), dtype=torch.uint8, device=dev)
    rows = []
    def row(name, ms, b):
        rows.append({"kernel": name, "envs": n, "ms": ms, "algorithmic_bytes_per_env": b, "gb_per_s": b * n / ms / 1e6, "frac_of_hbm_peak": b * n / ms / 1e6 / peak})
    # in: quat 16 + w 12 + joints 48; out: rpy 12 + rate 12 + motor 48 + feet 48
    row("rg_state_from_sim", time_ms(lambda: robot.set_sim_state(t, quat, v, w, joints, contacts)), 196)
    action, q, qd = f32(n, 60), f32(n, 12), f32(n, 12)
    tau = torch.empty((n, 12), dtype=torch.float32, device=dev)
    lib = rg.load()
    call = lambda: rg.check(lib.rg_hybrid_motor_torque(n, rg._ptr(action, torch.float32, (60,)), rg._ptr(q, torch.float32, (12,)),
                                                        rg._ptr(qd, torch.float32, (12,)), rg._ptr(tau, torch.float32, (12,)), rg.current_stream_ptr()))
    row("rg_hybrid_motor_torque", time_ms(call), 240 + 48 + 48 + 48)
    del action, q, qd, tau, robot, quat, v, w, joints
    # two control steps so that the fused prologue / epilogue kernels show up in an ncu launch list of this script
    m = min(n, 1 << 18)
    st = synthetic.make_states(m, GHOST, seed=1)
    srobot = SyntheticRobotBatch(GHOST, st)
    ctl = BatchedMPCController(srobot, srobot.GetTimeSinceReset, squeeze_single=False)
    ms = time_ms(ctl.step, reps=3, warm=1)
    print(json.dumps({"control_step_envs": m, "ms": ms}))
    for r in rows: print(json.dumps(r))


if __name__ == "__main__":
    main()
