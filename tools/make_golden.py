"""Generates the golden fixtures under tests/golden/ from the reference tree (build container only).

  reference_constants.json : the per-robot constants of robot_gym/model/robots/{ghost,k3lso}/*.py,
                             obtained by IMPORTING the reference modules (with the absent
                             third-party `mpc_controller.gait_generator` stubbed by its LegState enum),
                             plus MOTOR_COMMAND layout constants of simple_motor.py and the time
                             constants of core/sim_constants.py.
  leg_chains.json          : 3-joint leg chains parsed from the reference URDFs
                             (tools/extract_leg_chains.py).

/root/reference does not exist on the GPU box; tests only read the committed JSON.
  mpc_oracle_golden.npz, control_step_oracle_golden.npz : ORACLE outputs on seeded synthetic inputs
                             (--oracle); they freeze the oracle, they are not reference outputs.

Run:  python tools/make_golden.py [--oracle | --oracle-only]
"""
import enum
import importlib
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, HERE)
import extract_leg_chains  # noqa: E402


class LegState(enum.Enum):   # motion_imitation mpc_controller/gait_generator.py (recalled values)
    SWING = 0
    STANCE = 1
    EARLY_CONTACT = 2
    LOSE_CONTACT = 3


def _stub_third_party():
    pkg = types.ModuleType("mpc_controller")
    gg = types.ModuleType("mpc_controller.gait_generator")
    gg.LegState = LegState
    pkg.gait_generator = gg
    sys.modules["mpc_controller"] = pkg
    sys.modules["mpc_controller.gait_generator"] = gg


def _jsonable(v):
    if isinstance(v, enum.Enum):
        return int(v.value)
    if isinstance(v, np.ndarray):
        return v.tolist()
    if isinstance(v, (np.floating, np.integer)):
        return v.item()
    if isinstance(v, (list, tuple)):
        return [_jsonable(x) for x in v]
    if isinstance(v, (int, float, str, bool)) or v is None:
        return v
    raise TypeError(type(v))


def _dump_module(mod, names):
    return {n: _jsonable(getattr(mod, n)) for n in names}


def main():
    os.makedirs(OUT, exist_ok=True)
    _stub_third_party()
    sys.path.insert(0, REF)
    consts = {}
    for robot in ("ghost", "k3lso"):
        base = f"robot_gym.model.robots.{robot}"
        ctrl = importlib.import_module(base + ".ctrl_constants")
        cst = importlib.import_module(base + ".constants")
        mot = importlib.import_module(base + ".motor_constants")
        marks = importlib.import_module(base + ".marks")
        consts[robot] = {
            "ctrl_constants": _dump_module(ctrl, [
                "MPC_BODY_MASS", "MPC_BODY_INERTIA", "MPC_BODY_HEIGHT", "MPC_VELOCITY_MULTIPLIER",
                "STANCE_DURATION_SECONDS", "DUTY_FACTOR", "INIT_PHASE_FULL_CYCLE", "INIT_LEG_STATE",
                "VX_OFFSET", "VY_OFFSET", "WZ_OFFSET"]),
            "constants": _dump_module(cst, [
                "NUM_LEG", "START_POS", "INIT_ORIENTATION", "INIT_MOTOR_ANGLES", "IDENTITY_ORIENTATION",
                "DEFAULT_HIP_POSITIONS", "HIP_JOINT_OFFSET", "UPPER_LEG_JOINT_OFFSET", "LOWER_LEG_JOINT_OFFSET"]),
            "motor_constants": _dump_module(mot, [
                "NUM_MOTORS", "MOTOR_OFFSET", "MOTOR_DIRECTION", "MOTOR_POSITION_GAINS", "MOTOR_VELOCITY_GAINS"]),
            "marks": {"motor_names": marks.MARK_PARAMS["1"]["motor_names"],
                      "urdf_name": marks.MARK_PARAMS["1"]["urdf_name"],
                      "num_motors": marks.MARK_PARAMS["1"]["num_motors"],
                      "num_legs": marks.MARK_PARAMS["1"]["num_legs"]},
        }
    sm = importlib.import_module("robot_gym.model.robots.simple_motor")
    consts["simple_motor"] = _dump_module(sm, [
        "MOTOR_CONTROL_POSITION", "MOTOR_CONTROL_TORQUE", "MOTOR_CONTROL_HYBRID", "MOTOR_COMMAND_DIMENSION",
        "POSITION_INDEX", "POSITION_GAIN_INDEX", "VELOCITY_INDEX", "VELOCITY_GAIN_INDEX", "TORQUE_INDEX"])
    sc = importlib.import_module("robot_gym.core.sim_constants")
    consts["sim_constants"] = _dump_module(sc, ["ACTION_REPEAT", "SIMULATION_TIME_STEP"])
    # known answers of the reference's own HYBRID motor model (numpy only, importable here)
    model = sm.RobotMotorModel(num_motors=12, kp=[220.0] * 12, kd=[1.0, 2.0, 2.0] * 4,
                               motor_control_mode=sm.MOTOR_CONTROL_HYBRID)
    rng = np.random.default_rng(7)
    cmds = rng.uniform(-1, 1, (8, 60))
    cmds[:, 1::5] = rng.uniform(0, 300, (8, 12))
    cmds[:, 3::5] = rng.uniform(0, 5, (8, 12))
    q = rng.uniform(-1, 1, (8, 12))
    qd = rng.uniform(-3, 3, (8, 12))
    tau = [model.convert_to_torque(cmds[i], q[i], qd[i], qd[i], sm.MOTOR_CONTROL_HYBRID)[0].tolist() for i in range(8)]
    consts["hybrid_motor_kat"] = {"commands": cmds.tolist(), "q": q.tolist(), "qd": qd.tolist(), "torque": tau}
    with open(os.path.join(OUT, "reference_constants.json"), "w") as fh:
        json.dump(consts, fh, indent=1, sort_keys=True)

    chains = {}
    for robot in ("ghost", "k3lso"):
        urdf = os.path.join(REF, "robot_gym", "util", "pybullet_data", consts[robot]["marks"]["urdf_name"])
        chains[robot] = extract_leg_chains.parse(urdf, consts[robot]["marks"]["motor_names"])
    with open(os.path.join(OUT, "leg_chains.json"), "w") as fh:
        json.dump(chains, fh, indent=1, sort_keys=True)
    print("wrote", os.listdir(OUT))


if __name__ == "__main__" and "--oracle-only" not in sys.argv:
    main()


# ---------------------------------------------------------------------------------------------
# Oracle outputs on seeded synthetic inputs (the oracle restates third-party code that cannot run
# here, so these are "oracle goldens": they freeze the oracle's answers so that the GPU tests on
# the GPU box and future oracle edits are both checked against the same committed numbers).
def make_oracle_goldens():
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "robot-gym_b200"))
    from oracle import convex_mpc, kinematics, locomotion
    from robot_gym.model.robots.descriptions import GHOST
    from robot_gym.util import synthetic

    ctrl = GHOST.GetCtrlConstants()
    out = {}
    for horizon, n in ((10, 96), (5, 24), (20, 12)):
        st = synthetic.make_states(n, GHOST, seed=synthetic.SEED + horizon)
        mp = convex_mpc.MpcParams(horizon=horizon)
        forces = np.zeros((n, horizon * 12))
        for i in range(n):
            forces[i] = convex_mpc.compute_contact_forces(
                mp, st.com_velocity_body[i].astype(np.float64), st.base_rpy[i].astype(np.float64),
                st.base_rpy_rate[i].astype(np.float64), st.planned_contacts[i],
                st.foot_positions_base[i].astype(np.float64), [0, 0, ctrl.MPC_BODY_HEIGHT],
                [float(st.command[i, 0]), float(st.command[i, 1]), 0.0], [0, 0, 0], [0, 0, float(st.command[i, 2])])
        out[f"mpc_h{horizon}_forces"] = forces
        out[f"mpc_h{horizon}_n"] = np.array(n)
    np.savez_compressed(os.path.join(OUT, "mpc_oracle_golden.npz"), **out)

    # full control step: 6 envs x 40 steps through the restated LocomotionController
    n_env, n_steps = 6, 40
    seq = synthetic.make_state_sequence(n_env, n_steps, GHOST)
    actions = np.zeros((n_steps, n_env, 60), dtype=np.float32)
    desired = np.zeros((n_steps, n_env, 4), dtype=np.int32)
    state = np.zeros((n_steps, n_env, 4), dtype=np.int32)
    phase = np.zeros((n_steps, n_env, 4), dtype=np.float64)
    forces = np.zeros((n_steps, n_env, 12), dtype=np.float64)
    vbody = np.zeros((n_steps, n_env, 3), dtype=np.float64)
    for e in range(n_env):
        robot = kinematics.OracleRobot(GHOST)
        clock = {"t": 0.0}
        def load(k):
            s = seq[k]
            robot.set_state(base_velocity=s.base_velocity_world[e].astype(np.float64),
                            base_orientation=s.base_orientation_xyzw[e].astype(np.float64),
                            base_rpy=s.base_rpy[e].astype(np.float64), base_rpy_rate=s.base_rpy_rate[e].astype(np.float64),
                            foot_positions=s.foot_positions_base[e].astype(np.float64), foot_contacts=s.foot_contacts[e],
                            motor_angles=s.motor_angles[e].astype(np.float64))
            clock["t"] = float(s.time_since_reset[e])
        load(0)
        ctl = locomotion.build_mpc_controller(robot, lambda: clock["t"], ctrl)
        ctl.reset()
        for k in range(n_steps):
            load(k)
            s = seq[k]
            locomotion.update_controller_params(ctl, ctrl, (float(s.command[e, 0] - np.float32(ctrl.VX_OFFSET)),
                                                            float(s.command[e, 1] - np.float32(ctrl.VY_OFFSET)),
                                                            float(s.command[e, 2] - np.float32(ctrl.WZ_OFFSET))))
            # use the float32 command the GPU path sees
            ctl.swing_leg_controller.desired_speed = [float(s.command[e, 0]), float(s.command[e, 1]), 0.0]
            ctl.swing_leg_controller.desired_twisting_speed = float(s.command[e, 2])
            ctl.stance_leg_controller.desired_speed = [float(s.command[e, 0]), float(s.command[e, 1]), 0.0]
            ctl.stance_leg_controller.desired_twisting_speed = float(s.command[e, 2])
            ctl.update()
            actions[k, e] = ctl.get_action()
            desired[k, e] = ctl.gait_generator.desired_leg_state
            state[k, e] = ctl.gait_generator.leg_state
            phase[k, e] = ctl.gait_generator.normalized_phase
            forces[k, e] = ctl.stance_leg_controller.last_contact_forces
            vbody[k, e] = ctl.state_estimator.com_velocity_body_frame
    np.savez_compressed(os.path.join(OUT, "control_step_oracle_golden.npz"), actions=actions, desired_leg_state=desired,
                        leg_state=state, normalized_phase=phase, contact_forces=forces, com_velocity_body=vbody,
                        n_env=np.array(n_env), n_steps=np.array(n_steps))
    print("oracle goldens written")


if __name__ == "__main__" and ("--oracle" in sys.argv or "--oracle-only" in sys.argv):
    make_oracle_goldens()
