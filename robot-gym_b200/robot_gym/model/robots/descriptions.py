"""Robot descriptions: the per-robot parameters the MPC path reads, values unchanged from the reference.

The reference keeps these in three python modules per robot, fetched through
``robot.GetCtrlConstants() / GetConstants() / GetMotorConstants()`` (ghost/ghost.py:7-30).  A
:class:`RobotDescription` exposes the same three namespaces with the same attribute names, so
``BatchedMPCController`` accepts either a reference ``Robot`` class or one of these.  Values are
checked against the reference modules by tests/test_constants.py (fixture made by
tools/make_golden.py).
"""
from __future__ import annotations

import types

import numpy as np

from robot_gym.controllers.mpc.leg_state import LegState
from robot_gym.model.robots.leg_chains import LEG_CHAINS

MOTOR_CONTROL_HYBRID = 3          # simple_motor.py:10
MOTOR_COMMAND_DIMENSION = 5       # simple_motor.py:14


def _ns(**kw):
    return types.SimpleNamespace(**kw)


class RobotDescription:
    def __init__(self, name, ctrl_constants, constants, motor_constants, leg_chains):
        self.name = name
        self._ctrl, self._const, self._motor = ctrl_constants, constants, motor_constants
        self.leg_chains = leg_chains

    # same accessor names as model/robots/ghost/ghost.py:7-30
    def GetCtrlConstants(self):
        return self._ctrl

    def GetConstants(self):
        return self._const

    def GetMotorConstants(self):
        return self._motor


def _trot_gait():
    # ghost/ctrl_constants.py:13,27-37 (k3lso identical)
    return dict(
        STANCE_DURATION_SECONDS=[0.3] * 4,
        DUTY_FACTOR=[0.6] * 4,
        INIT_PHASE_FULL_CYCLE=[0.9, 0, 0, 0.9],
        INIT_LEG_STATE=(LegState.SWING, LegState.STANCE, LegState.STANCE, LegState.SWING),
    )


GHOST = RobotDescription(
    "ghost",
    ctrl_constants=_ns(
        MPC_BODY_MASS=190 / 9.8,                                                  # ghost/ctrl_constants.py:8
        MPC_BODY_INERTIA=(0.07335, 0, 0, 0, 0.25068, 0, 0, 0, 0.25447),           # :9
        MPC_BODY_HEIGHT=0.42,                                                     # :10
        MPC_VELOCITY_MULTIPLIER=1.0,                                              # :11
        VX_OFFSET=0.0, VY_OFFSET=0.08, WZ_OFFSET=-0.025,                          # :39-41
        **_trot_gait()),
    constants=_ns(
        NUM_LEG=4,                                                                # ghost/constants.py:4
        INIT_MOTOR_ANGLES=np.array([0, 0.67, -1.25] * 4, dtype=np.float64),       # :8-17
        DEFAULT_HIP_POSITIONS=((0.22, -0.1, 0), (0.22, 0.1, 0), (-0.22, -0.1, 0), (-0.22, 0.1, 0)),   # :31-36
        IDENTITY_ORIENTATION=[0, 0, 0, 1]),
    motor_constants=_ns(
        NUM_MOTORS=12,                                                            # ghost/motor_constants.py:5
        MOTOR_OFFSET=np.zeros(12), MOTOR_DIRECTION=np.ones(12),                   # :9-11
        MOTOR_POSITION_GAINS=[220.0] * 12,                                        # :13
        MOTOR_VELOCITY_GAINS=np.array([1.0, 2.0, 2.0] * 4)),                      # :15
    leg_chains=LEG_CHAINS["ghost"],
)

K3LSO = RobotDescription(
    "k3lso",
    ctrl_constants=_ns(
        MPC_BODY_MASS=190 / 9.8,                                                  # k3lso/ctrl_constants.py:8
        MPC_BODY_INERTIA=(0.07335, 0, 0, 0, 0.25068, 0, 0, 0, 0.25447),           # :10
        MPC_BODY_HEIGHT=0.38,                                                     # :11
        MPC_VELOCITY_MULTIPLIER=1.0,
        VX_OFFSET=0.0, VY_OFFSET=0.0, WZ_OFFSET=0.0,                              # :39-41
        **_trot_gait()),
    constants=_ns(
        NUM_LEG=4,
        INIT_MOTOR_ANGLES=np.array([0, 0.67, -1.25, -0.0, 0.67, 1.25, 0, -0.67, -1.25, 0, -0.67, 1.25],
                                   dtype=np.float64),                             # k3lso/constants.py:12-18
        DEFAULT_HIP_POSITIONS=((0.22, -0.105, 0), (0.22, 0.105, 0), (-0.22, -0.105, 0), (-0.22, 0.105, 0)),  # :32-37
        IDENTITY_ORIENTATION=[0, 0, 0, 1]),
    motor_constants=_ns(
        NUM_MOTORS=12,
        MOTOR_OFFSET=np.zeros(12), MOTOR_DIRECTION=np.ones(12),
        MOTOR_POSITION_GAINS=[220.0] * 12,
        MOTOR_VELOCITY_GAINS=np.array([1.0, 2.0, 2.0] * 4)),
    leg_chains=LEG_CHAINS["k3lso"],
)

ROBOTS = {"ghost": GHOST, "k3lso": K3LSO}

# Contact schedules for BASELINE config 4.  Only "trot" exists in the reference
# (ghost/ctrl_constants.py:27-37); pace / bound / walk are BUILDER-DEFINED in the same
# parameterisation (leg order FR, FL, RR, RL) -- SURVEY.md section 8(d).
_S, _T = LegState.SWING, LegState.STANCE
GAIT_SCHEDULES = {
    "trot": dict(STANCE_DURATION_SECONDS=[0.3] * 4, DUTY_FACTOR=[0.6] * 4,
                 INIT_PHASE_FULL_CYCLE=[0.9, 0, 0, 0.9], INIT_LEG_STATE=(_S, _T, _T, _S)),
    "pace": dict(STANCE_DURATION_SECONDS=[0.3] * 4, DUTY_FACTOR=[0.6] * 4,
                 INIT_PHASE_FULL_CYCLE=[0.9, 0, 0.9, 0], INIT_LEG_STATE=(_S, _T, _S, _T)),
    "bound": dict(STANCE_DURATION_SECONDS=[0.3] * 4, DUTY_FACTOR=[0.6] * 4,
                  INIT_PHASE_FULL_CYCLE=[0.9, 0.9, 0, 0], INIT_LEG_STATE=(_S, _S, _T, _T)),
    "walk": dict(STANCE_DURATION_SECONDS=[0.3] * 4, DUTY_FACTOR=[0.75] * 4,
                 INIT_PHASE_FULL_CYCLE=[0, 0.5, 0.75, 0.25], INIT_LEG_STATE=(_T, _T, _T, _T)),
    "stand": dict(STANCE_DURATION_SECONDS=[0.3] * 4, DUTY_FACTOR=[1.0] * 4,      # ghost/ctrl_constants.py:16-25 (commented out)
                  INIT_PHASE_FULL_CYCLE=[0.0] * 4, INIT_LEG_STATE=(_T, _T, _T, _T)),
}


def with_gait(description: RobotDescription, schedule: str) -> RobotDescription:
    """Copy of ``description`` whose ctrl constants use the named contact schedule."""
    ctrl = types.SimpleNamespace(**vars(description.GetCtrlConstants()))
    for k, v in GAIT_SCHEDULES[schedule].items():
        setattr(ctrl, k, v)
    return RobotDescription(description.name, ctrl, description.GetConstants(),
                            description.GetMotorConstants(), description.leg_chains)
