"""B200-native batched MPC locomotion controller behind robot-gym's ``Controller`` interface.

Only the per-control-step controller hot path of nicrusso7/robot-gym lives here (gait generator,
COM velocity estimator, Raibert swing controller, leg IK, convex-MPC stance QP, hybrid action
packing), batched over independent envs and executed by hand-written sm_100a CUDA kernels through
the C ABI in ``robot_gym.cuda`` (include/rg_cuda.h).  Module paths mirror the reference's
(``robot_gym.controllers``, ``robot_gym.model``, ``robot_gym.cuda``) so the files overlay onto a
robot-gym checkout; see INTEGRATION.md.
"""
__version__ = "0.1.0"
