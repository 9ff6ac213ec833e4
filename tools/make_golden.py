"""Generates the golden fixtures under tests/golden/ from the reference tree (build container only).

  reference_constants.json : the per-robot constants of robot_gym/model/robots/{ghost,k3lso}/*.py,
                             obtained by IMPORTING the reference modules (with the absent
                             third-party `mpc_controller.gait_generator` stubbed by its LegState enum),
                             plus MOTOR_COMMAND layout constants of simple_motor.py and the time
                             constants of core/sim_constants.py.
  leg_chains.json          : 3-joint leg chains parsed from the reference URDFs
                             (tools/extract_leg_chains.py).

/root/reference does not exist on the GPU box; tests only read the committed JSON.
Run:  python tools/make_golden.py
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


if __name__ == "__main__":
    main()
