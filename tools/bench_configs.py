"""BASELINE.json configs 2-4 + latency table on one GPU (writes one JSON document).

    python tools/bench_configs.py [out.json]

config 2: 4096 envs MPC stance solve (h = 10), gait-derived and all-stance contacts
config 3: 65536 envs full control step (gait + estimator + swing + IK + MPC + pack); cold = every QP from scratch,
          warm_start = seeded with the previous step's verified active set (static synthetic states: best case)
config 4: horizon 5/10/20 x {trot, pace, bound, walk} contact schedules, 65536 envs, MPC solve
config 5: the MPC solve at 2^20 envs on ONE GPU (the whole config-5 batch) and at 2^17 (one rank's share of it on 8 GPUs)
light   : rg_state_from_sim / rg_hybrid_motor_torque at 2^20 envs against the HBM roofline
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
from robot_gym.model.robots.sim_state_robot import SimStateRobotBatch
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


def control_case(n, reps, warm, warm_start=False):
    st = synthetic.make_states(n, GHOST)
    robot = SyntheticRobotBatch(GHOST, st)
    ctl = BatchedMPCController(robot, robot.GetTimeSinceReset, squeeze_single=False, warm_start=warm_start)
    ctl.command.copy_(torch.from_numpy(st.command).cuda())
    ms = sorted(time_ms(ctl.step, reps, warm))
    p50, p99 = ms[len(ms) // 2], ms[min(len(ms) - 1, int(0.99 * len(ms)))]
    return {"envs": n, "warm_start": warm_start, "p50_ms": p50, "p99_ms": p99, "env_steps_per_s": n / p50 * 1e3, "launches_per_step": 3,
            "active_set_rounds_mean": float(ctl.solve_info[:, rg.RG_INFO_POLISH_ROUNDS].to(torch.float64).mean())}


def light_kernel_cases(n):
    """HBM-bound kernels either side of the solve, at n envs: achieved GB/s of algorithmic bytes against the
    measured HBM peak (MEASURED_PEAKS.json when present)."""
    peak = 6558.7
    try:
        peak = float(json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    rng = np.random.default_rng(3)
    dev = "cuda"
    f32 = lambda *shape: torch.from_numpy(rng.uniform(-0.3, 0.3, shape).astype(np.float32)).to(dev)
    quat = torch.from_numpy(synthetic.euler_to_quat_xyzw(rng.uniform(-0.5, 0.5, (n, 3))).astype(np.float32)).to(dev)
    robot = SimStateRobotBatch(GHOST, n, device=dev)
    t = torch.zeros(n, dtype=torch.float64, device=dev)
    v, w, joints = f32(n, 3), f32(n, 3), f32(n, 12)
    contacts = torch.ones((n, 4), dtype=torch.uint8, device=dev)
    rows = []
    ms = statistics.median(time_ms(lambda: robot.set_sim_state(t, quat, v, w, joints, contacts), 20, 3))
    b = 196 * n     # in: quat 16 + w 12 + joints 48; out: rpy 12 + rate 12 + motor 48 + feet 48
    rows.append({"kernel": "rg_state_from_sim", "envs": n, "ms": ms, "algorithmic_bytes": b, "gb_per_s": b / ms / 1e6, "frac_of_hbm_peak": b / ms / 1e6 / peak})
    action, q, qd = f32(n, 60), f32(n, 12), f32(n, 12)
    tau = torch.empty((n, 12), dtype=torch.float32, device=dev)
    lib = rg.load()
    call = lambda: rg.check(lib.rg_hybrid_motor_torque(n, rg._ptr(action, torch.float32, (60,)), rg._ptr(q, torch.float32, (12,)),
                                                        rg._ptr(qd, torch.float32, (12,)), rg._ptr(tau, torch.float32, (12,)), rg.current_stream_ptr()))
    ms = statistics.median(time_ms(call, 20, 3))
    b = (240 + 48 + 48 + 48) * n
    rows.append({"kernel": "rg_hybrid_motor_torque", "envs": n, "ms": ms, "algorithmic_bytes": b, "gb_per_s": b / ms / 1e6, "frac_of_hbm_peak": b / ms / 1e6 / peak})
    return rows


def main():
    out = {"gpu": torch.cuda.get_device_name(0), "library": rg.load().rg_version().decode()}
    out["config2_mpc_4096"] = [mpc_case(4096, 10, "trot"), mpc_case(4096, 10, "trot", all_stance=True)]
    out["config3_control_step_65536"] = [control_case(65536, 10, 3), control_case(65536, 10, 3, warm_start=True)]
    out["config4_horizon_x_schedule_65536"] = [mpc_case(65536 if h < 20 else 16384, h, s, reps=5) for h in (5, 10, 20) for s in ("trot", "pace", "bound", "walk")]
    out["latency_control_step"] = [control_case(n, 200 if n < 65536 else 30, 20 if n < 65536 else 3, warm_start=w) for n in (1, 4096, 65536) for w in (False, True)]
    out["config5_single_gpu_share_of_2p20_envs"] = [mpc_case(1 << 20, 10, "trot", reps=3), mpc_case(1 << 17, 10, "trot", reps=5)]
    out["light_kernels_hbm"] = light_kernel_cases(1 << 20)
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
