"""BASELINE config[0] substitute: the reference run -- Ghost playground trot, PyBullet DIRECT, 1000
control steps on the CPU -- CANNOT be executed here (pybullet and motion_imitation are absent offline).  What runs
instead, clearly labelled: the in-repo ORACLE controller (oracle/locomotion.py: restated gait generator + estimator +
Raibert swing + IK + stance QP), not the reference stack, looped over 1000 control steps of ONE synthetic env (seeded
state trace, no physics), timed per control step on one host core -- next to BatchedMPCController with N = 1 on the same
trace (wall clock per get_action(), including the host read of the [60] command: the drop-in latency a playground loop
sees).  Both are compared with the 10 ms control period (core/sim_constants.py:7,11: ACTION_REPEAT x 1 ms).

    python tools/config1_substitute.py [--steps 1000] [--numpy-steps 100] [--out profiles/r02_config1_substitute.json]
"""
import argparse, json, os, statistics, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "robot-gym_b200")); sys.path.insert(0, REPO)
import numpy as np
from robot_gym.model.robots.descriptions import GHOST
from robot_gym.util import synthetic

BUDGET_MS = 10.0


def pct(v, q):
    v = sorted(v)
    return v[min(len(v) - 1, int(round(q * (len(v) - 1))))]


def run_oracle(seq, steps, solver, label):
    from oracle import kinematics, locomotion
    ctrl = GHOST.GetCtrlConstants()
    robot = kinematics.OracleRobot(GHOST)
    clock = {"t": 0.0}

    def load(k):
        s = seq[k]
        robot.set_state(base_velocity=s.base_velocity_world[0].astype(np.float64), base_orientation=s.base_orientation_xyzw[0].astype(np.float64),
                        base_rpy=s.base_rpy[0].astype(np.float64), base_rpy_rate=s.base_rpy_rate[0].astype(np.float64),
                        foot_positions=s.foot_positions_base[0].astype(np.float64), foot_contacts=s.foot_contacts[0],
                        motor_angles=s.motor_angles[0].astype(np.float64))
        clock["t"] = float(s.time_since_reset[0])
    load(0)
    ctl = locomotion.build_mpc_controller(robot, lambda: clock["t"], ctrl, mpc_solver=solver)
    ctl.reset()
    ms, actions = [], []
    for k in range(steps):
        load(k)
        t0 = time.perf_counter()
        locomotion.update_controller_params(ctl, ctrl, (0.2, 0.0, 0.1))
        ctl.update()
        a = ctl.get_action()
        ms.append((time.perf_counter() - t0) * 1e3)
        actions.append(a)
    return {"what": label, "steps": steps, "p50_ms": pct(ms, 0.5), "p99_ms": pct(ms, 0.99), "mean_ms": statistics.mean(ms),
            "fraction_of_steps_within_10ms": float(np.mean(np.array(ms) <= BUDGET_MS)), "cores": 1}, np.array(actions)


def run_gpu(seq, steps):
    import torch
    from robot_gym.controllers.mpc.batched_mpc_controller import BatchedMPCController
    from robot_gym.model.robots.synthetic_robot import SyntheticRobotBatch
    robot = SyntheticRobotBatch(GHOST, seq[0], device="cuda")
    out = {}
    actions = None
    for key, warm in (("cold", False), ("warm_start", True)):
        robot.load(seq[0])
        ctl = BatchedMPCController(robot, robot.GetTimeSinceReset, warm_start=warm)
        ms, acts = [], []
        for k in range(steps):
            robot.load(seq[k])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ctl.update_controller_params((0.2, 0.0, 0.1))
            a = ctl.get_action()                       # [60] float32 numpy: includes the device -> host read
            ms.append((time.perf_counter() - t0) * 1e3)
            acts.append(a)
        ms = ms[20:]                                    # first steps pay one-time CUDA initialisation
        out[key] = {"steps": len(ms), "p50_ms": pct(ms, 0.5), "p99_ms": pct(ms, 0.99), "mean_ms": statistics.mean(ms),
                    "fraction_of_steps_within_10ms": float(np.mean(np.array(ms) <= BUDGET_MS))}
        if not warm:
            actions = np.array(acts)
    return out, actions


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--numpy-steps", type=int, default=100)
    ap.add_argument("--out", default=os.path.join(REPO, "profiles", "r02_config1_substitute.json"))
    args = ap.parse_args()
    from oracle import c_oracle, convex_mpc
    seq = synthetic.make_state_sequence(1, args.steps, GHOST, seed=synthetic.SEED + 1)
    res = {"label": "SUBSTITUTE for BASELINE config[0]: in-repo oracle controller on a synthetic single-env trace, NOT the "
                    "reference stack (pybullet / motion_imitation unavailable offline); no physics in the loop",
           "control_period_budget_ms": BUDGET_MS, "trace": f"make_state_sequence(1, {args.steps}, GHOST): fresh random state per step, "
           "clock advancing 10 ticks of 1 ms per control step"}
    res["oracle_c_solver"], a_c = run_oracle(seq, args.steps, c_oracle.compute_contact_forces,
                                              "oracle/locomotion.py (python) with the C port of the stance QP (oracle/c/mpc_oracle.c)")
    if args.numpy_steps > 0:
        res["oracle_numpy_solver"], a_np = run_oracle(seq, args.numpy_steps, convex_mpc.compute_contact_forces,
                                                      "oracle/locomotion.py with the numpy stance QP (oracle/convex_mpc.py)")
        res["oracle_numpy_vs_c_max_action_diff"] = float(np.abs(a_np - a_c[:args.numpy_steps]).max())
    try:
        import torch
        if torch.cuda.is_available():
            res["batched_controller_n1"], a_gpu = run_gpu(seq, args.steps)
            ref = a_c.reshape(args.steps, 12, 5)
            got = a_gpu.reshape(args.steps, 12, 5)
            res["batched_vs_oracle"] = {"max_swing_joint_target_diff_rad": float(np.abs(got[:, :, 0] - ref[:, :, 0]).max()),
                                        "max_torque_rel_diff": float((np.abs(got[:, :, 4] - ref[:, :, 4]).max(axis=1) /
                                                                      np.maximum(1.0, np.abs(ref[:, :, 4]).max(axis=1))).max())}
            res["gpu"] = torch.cuda.get_device_name(0)
        else:
            res["batched_controller_n1"] = "no CUDA device on this machine"
    except Exception as exc:
        res["batched_controller_n1"] = f"failed: {exc!r}"
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(res, fh, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
