"""``LegState`` of motion_imitation's ``mpc_controller.gait_generator`` (third-party, absent).

The reference imports it in robot_gym/model/robots/ghost/ctrl_constants.py:3 and uses
SWING / STANCE in INIT_LEG_STATE (:32-37).  Integer values are the C ABI's RG_LEG_* codes.
"""
import enum


class LegState(enum.IntEnum):
    SWING = 0
    STANCE = 1
    EARLY_CONTACT = 2     # swing leg that touched down before the planned time
    LOSE_CONTACT = 3      # stance leg that lost contact
