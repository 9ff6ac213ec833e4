"""The robot parameters the MPC path reads are UNCHANGED from the reference.

Fixture: tests/golden/reference_constants.json, produced by tools/make_golden.py by importing
robot_gym/model/robots/{ghost,k3lso}/{ctrl_constants,constants,motor_constants,marks}.py,
robot_gym/model/robots/simple_motor.py and robot_gym/core/sim_constants.py from /root/reference.
"""
import numpy as np
import pytest

from robot_gym.controllers.mpc.leg_state import LegState
from robot_gym.model.robots import descriptions


@pytest.mark.parametrize("name", ["ghost", "k3lso"])
def test_ctrl_constants_match_reference(reference_constants, name):
    ref = reference_constants[name]["ctrl_constants"]
    ours = descriptions.ROBOTS[name].GetCtrlConstants()
    for key, value in ref.items():
        got = getattr(ours, key)
        if key == "INIT_LEG_STATE":
            got = [int(s) for s in got]
        np.testing.assert_array_equal(np.asarray(got, dtype=np.float64), np.asarray(value, dtype=np.float64), err_msg=key)


@pytest.mark.parametrize("name", ["ghost", "k3lso"])
def test_robot_and_motor_constants_match_reference(reference_constants, name):
    ref = reference_constants[name]
    d = descriptions.ROBOTS[name]
    for key in ("NUM_LEG", "INIT_MOTOR_ANGLES", "DEFAULT_HIP_POSITIONS", "IDENTITY_ORIENTATION"):
        np.testing.assert_array_equal(np.asarray(getattr(d.GetConstants(), key), dtype=np.float64),
                                      np.asarray(ref["constants"][key], dtype=np.float64), err_msg=key)
    for key in ("NUM_MOTORS", "MOTOR_OFFSET", "MOTOR_DIRECTION", "MOTOR_POSITION_GAINS", "MOTOR_VELOCITY_GAINS"):
        np.testing.assert_array_equal(np.asarray(getattr(d.GetMotorConstants(), key), dtype=np.float64),
                                      np.asarray(ref["motor_constants"][key], dtype=np.float64), err_msg=key)
    assert ref["marks"]["num_motors"] == 12 and ref["marks"]["num_legs"] == 4
    assert [c["joint_names"][0] for c in d.leg_chains] == ref["marks"]["motor_names"][0::3]


def test_leg_chains_match_urdf_fixture(golden_dir):
    import json, os
    with open(os.path.join(golden_dir, "leg_chains.json")) as fh:
        ref = json.load(fh)
    for name, chains in ref.items():
        ours = descriptions.ROBOTS[name].leg_chains
        for leg in range(4):
            for key in ("p", "r", "axis", "toe"):
                np.testing.assert_array_equal(np.asarray(ours[leg][key], dtype=np.float64),
                                              np.asarray(chains[leg][key], dtype=np.float64))


def test_hybrid_layout_and_enums(reference_constants):
    sm = reference_constants["simple_motor"]
    assert descriptions.MOTOR_CONTROL_HYBRID == sm["MOTOR_CONTROL_HYBRID"]
    assert descriptions.MOTOR_COMMAND_DIMENSION == sm["MOTOR_COMMAND_DIMENSION"] == 5
    assert (sm["POSITION_INDEX"], sm["POSITION_GAIN_INDEX"], sm["VELOCITY_INDEX"], sm["VELOCITY_GAIN_INDEX"],
            sm["TORQUE_INDEX"]) == (0, 1, 2, 3, 4)
    assert reference_constants["sim_constants"] == {"ACTION_REPEAT": 10, "SIMULATION_TIME_STEP": 0.001}
    assert [int(LegState.SWING), int(LegState.STANCE), int(LegState.EARLY_CONTACT), int(LegState.LOSE_CONTACT)] == [0, 1, 2, 3]


def test_gait_schedules_cover_config4():
    assert set(descriptions.GAIT_SCHEDULES) >= {"trot", "pace", "bound", "walk"}
    trot = descriptions.GAIT_SCHEDULES["trot"]
    ghost = descriptions.GHOST.GetCtrlConstants()
    assert trot["DUTY_FACTOR"] == ghost.DUTY_FACTOR and trot["INIT_PHASE_FULL_CYCLE"] == ghost.INIT_PHASE_FULL_CYCLE
    walk = descriptions.with_gait(descriptions.GHOST, "walk").GetCtrlConstants()
    assert walk.DUTY_FACTOR == [0.75] * 4 and walk.MPC_BODY_HEIGHT == ghost.MPC_BODY_HEIGHT
