"""BASELINE.json configs 2-4 + latency table on one GPU (writes one JSON document).

    python tools/bench_configs.py [out.json]

config 2: 4096 envs MPC stance solve (h = 10), gait-derived and all-stance contacts
config 3: 65536 envs full control step (gait + estimator + swing + IK + MPC + pack)
config 4: horizon 5/10/20 x {trot, pace, bound, walk} contact schedules, 65536 envs, MPC solve
latency : p50 / p99 of one control step for N in {1, 4096, 65536} (CUDA events, 200 reps after 20 warm-ups)
pace / bound / walk are builder-defined schedules (only trot exists in the reference).
"""
import json, os, statistics, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "robot-gym_b200")); sys.path.insert(0, REPO)
import numpy as np, torch
from robot_gym import cuda as rg
from robot_gym.controllers.mpc.batched_mpc_controller import BatchedMPCController
from robot_gym.model.robots.descriptions import GHOST, with_gait
from robot_gym.model.robots.synthetic_robot import SyntheticRobotBatch
from robot_gym.util import synthetic


def time_ms(fn, reps, warm):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        out.append(a.elapsed_time(b))
    return out


def mpc_case(n, horizon, schedule, all_stance=False, reps=10):
    desc = with_gait(GHOST, schedule); ctrl = desc.GetCtrlConstants()
    p = rg.default_mpc_params(ctrl.MPC_BODY_MASS, ctrl.MPC_BODY_INERTIA, ctrl.MPC_BODY_HEIGHT, horizon)
    ws = rg.MpcWorkspace(p)
    st = synthetic.make_states(n, desc, schedule_ctrl=ctrl, all_stance=all_stance)
    t = lambda a: torch.from_numpy(a).cuda()
    args = (t(st.com_velocity_body), t(st.base_rpy), t(st.base_rpy_rate), t(st.planned_contacts), t(st.foot_positions_base), t(st.command))
    f = torch.empty((n, 12), dtype=torch.float32, device="cuda"); info = torch.empty((n, 4), dtype=torch.int32, device="cuda")
    ms = time_ms(lambda: rg.mpc_build_solve(ws, *args, contact_forces=f, solve_info=info), reps, 3)
    inf = info.cpu().numpy()
    med = statistics.median(ms)
    return {"envs": n, "horizon": horizon, "schedule": schedule, "all_stance": all_stance, "ms": med, "solves_per_s": n / med * 1e3,
            "ipm_iters_mean": float(inf[:, 0].mean()), "ipm_iters_max": int(inf[:, 0].max()), "polish_rounds_mean": float(inf[:, 1].mean()),
            "polished_fraction": float(np.mean((inf[:, 2] & 1) != 0)), "stance_legs_mean": float(st.planned_contacts.sum(axis=1).mean())}


def control_case(n, reps, warm):
    st = synthetic.make_states(n, GHOST)
    robot = SyntheticRobotBatch(GHOST, st)
    ctl = BatchedMPCController(robot, robot.GetTimeSinceReset, squeeze_single=False)
    ctl.command.copy_(torch.from_numpy(st.command).cuda())
    ms = sorted(time_ms(ctl.step, reps, warm))
    p50, p99 = ms[len(ms) // 2], ms[min(len(ms) - 1, int(0.99 * len(ms)))]
    return {"envs": n, "p50_ms": p50, "p99_ms": p99, "env_steps_per_s": n / p50 * 1e3, "launches_per_step": 3}


def main():
    out = {"gpu": torch.cuda.get_device_name(0), "library": rg.load().rg_version().decode()}
    out["config2_mpc_4096"] = [mpc_case(4096, 10, "trot"), mpc_case(4096, 10, "trot", all_stance=True)]
    out["config3_control_step_65536"] = control_case(65536, 10, 3)
    out["config4_horizon_x_schedule_65536"] = [mpc_case(65536 if h < 20 else 16384, h, s, reps=5) for h in (5, 10, 20) for s in ("trot", "pace", "bound", "walk")]
    out["latency_control_step"] = [control_case(n, 200 if n < 65536 else 30, 20 if n < 65536 else 3) for n in (1, 4096, 65536)]
    out["fma_peak_tflops"] = {"fp64": rg.measure_fma_peak(True), "fp32": rg.measure_fma_peak(False)}
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "gpurun_out", "configs.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as fh: json.dump(out, fh, indent=1)
    for k, v in out.items():
        print(k, json.dumps(v) if not isinstance(v, list) else "")
        if isinstance(v, list):
            for row in v: print("   ", json.dumps(row))


if __name__ == "__main__":
    main()
