"""The slice of ``Simulation`` (robot_gym/core/simulation.py) that the control loop touches, for N envs in one process.

``Simulation`` owns the PyBullet client, the robot and the controller, counts physics ticks
(``GetTimeSinceReset = step_counter * SIMULATION_TIME_STEP``, :141-142) and applies a control action for
``ACTION_REPEAT`` ticks (``ApplyStepAction``, :175-179; each tick = ``Robot.ApplyAction`` -> motor model ->
``stepSimulation``, :169-173).  ``BatchedSimulation`` keeps exactly that contract with tensors:

  * physics is a PLUGGABLE object (``BatchedPhysics``) -- PyBullet is not reimplemented (north_star); tests use the
    ``SyntheticPhysics`` joint integrator below;
  * the robot getters come from ``SimStateRobotBatch`` (one ``rg_state_from_sim`` launch per refresh);
  * the motor model is the sm_100a kernel behind ``rg_hybrid_motor_torque_ex`` (strength ratios and motor
    direction applied as ``Robot.ApplyAction`` does, robot.py:276-307); its FIRST evaluation of a control step is
    fused into the controller's epilogue (no [N,60] read-back), the other nine run once per tick.
"""
from __future__ import annotations

import ctypes

import torch

from robot_gym import cuda as rg
from robot_gym.core import sim_constants
from robot_gym.model.robots.sim_state_robot import SimStateRobotBatch


class BatchedPhysics:
    """What a batched rigid-body simulator has to expose (all CUDA tensors, env-major):

        base_orientation_xyzw [N,4] f32, base_velocity_world [N,3] f32, base_angular_velocity_world [N,3] f32,
        joint_angles [N,12] f32 (raw URDF angles), joint_velocities [N,12] f32, foot_contacts [N,4] u8

    and two methods: ``step(applied_motor_torques [N,12] f32, dt)`` -- one physics tick -- and
    ``reset(env_ids)`` -- put the given envs (LongTensor, None = all) back into their start pose."""

    num_envs: int

    def step(self, applied_motor_torques, dt):
        raise NotImplementedError

    def reset(self, env_ids=None):
        raise NotImplementedError


class SyntheticPhysics(BatchedPhysics):
    """NOT a physics engine: a deterministic stand-in for tests and demos.  Every joint is a damped inertia driven by
    the applied motor torque (plus a weak spring towards the start pose against drift); the ground is a one-sided
    constraint on each foot -- a tick that would push a foot below its standing height is undone for that leg
    (joints locked), which is what the feed-forward stance torques do all the time; the base is held level.  A foot
    is 'in contact' while its base-frame height is within ``contact_band`` of the standing height.  It closes the
    loop just enough to exercise the whole control path: swing legs lift (contacts switch off), PD torques pull the
    joints to their IK targets, stance legs stay planted, resets re-arm envs."""

    def __init__(self, description, num_envs, kinematics, device="cuda", inertia=0.02, damping=0.5, stiffness=5.0,
                 contact_band=0.004):
        self.description, self.num_envs, self.device = description, int(num_envs), torch.device(device)
        self._kin = kinematics
        mc, c = description.GetMotorConstants(), description.GetConstants()
        n, dev = self.num_envs, self.device
        f32 = dict(dtype=torch.float32, device=dev)
        self._direction = torch.tensor(list(mc.MOTOR_DIRECTION), **f32)
        self._offset = torch.tensor(list(mc.MOTOR_OFFSET), **f32)
        self._q0 = torch.tensor(list(c.INIT_MOTOR_ANGLES), **f32) * self._direction + self._offset      # URDF angles
        self.base_orientation_xyzw = torch.zeros((n, 4), **f32)
        self.base_velocity_world = torch.zeros((n, 3), **f32)
        self.base_angular_velocity_world = torch.zeros((n, 3), **f32)
        self.joint_angles = torch.zeros((n, 12), **f32)
        self.joint_velocities = torch.zeros((n, 12), **f32)
        self.foot_contacts = torch.ones((n, 4), dtype=torch.uint8, device=dev)
        self.forced_airborne = torch.zeros(n, dtype=torch.bool, device=dev)     # tests: envs that have "fallen"
        self._inertia, self._damping, self._stiffness, self._band = inertia, damping, stiffness, contact_band
        self.reset()
        feet = self._kin.ComputeFootPositionsInBaseFrame(self._motor_angles(self.joint_angles)).view(n, 4, 3)
        self._stand_z = feet[0, :, 2].clone()                                    # per leg (ghost is left/right asymmetric)

    def _motor_angles(self, joint_angles):
        return ((joint_angles - self._offset) * self._direction).contiguous()

    def reset(self, env_ids=None):
        sel = slice(None) if env_ids is None else env_ids
        self.base_orientation_xyzw[sel] = torch.tensor([0.0, 0.0, 0.0, 1.0], device=self.device)
        self.base_velocity_world[sel] = 0
        self.base_angular_velocity_world[sel] = 0
        self.joint_angles[sel] = self._q0
        self.joint_velocities[sel] = 0
        self.foot_contacts[sel] = 1
        self.forced_airborne[sel] = False

    def step(self, applied_motor_torques, dt):
        n = self.num_envs
        acc = (applied_motor_torques - self._damping * self.joint_velocities -
               self._stiffness * (self.joint_angles - self._q0)) / self._inertia
        qd = self.joint_velocities + dt * acc                     # semi-implicit Euler, tentatively
        q = self.joint_angles + dt * qd
        z = self._kin.ComputeFootPositionsInBaseFrame(self._motor_angles(q)).view(n, 4, 3)[:, :, 2]
        blocked = (z < self._stand_z).repeat_interleave(3, dim=1)  # the ground: this leg's tick is undone
        self.joint_angles.copy_(torch.where(blocked, self.joint_angles, q))
        self.joint_velocities.copy_(torch.where(blocked, torch.zeros_like(qd), qd))
        contact = (z <= self._stand_z + self._band) & ~self.forced_airborne[:, None]
        self.foot_contacts.copy_(contact.to(torch.uint8))


class BatchedSimulation:
    """N envs: tick counters, the state provider, the controller, ``ApplyStepAction``."""

    def __init__(self, description, physics: BatchedPhysics, controller_class, device="cuda", controller_kwargs=None,
                 fuse_motor_model=True):
        self.description = description
        self.device = torch.device(device)
        self.num_envs = int(physics.num_envs)
        self._physics = physics
        n, dev = self.num_envs, self.device
        self._step_counter = torch.zeros(n, dtype=torch.int64, device=dev)
        self._time = torch.zeros(n, dtype=torch.float64, device=dev)
        self._robot = SimStateRobotBatch(description, n, device=dev)
        self._robot.motor_velocities = torch.zeros((n, 12), dtype=torch.float32, device=dev)
        self._robot.GetMotorVelocities = lambda: self._robot.motor_velocities          # robot.py:256-264
        self._direction = torch.tensor(list(description.GetMotorConstants().MOTOR_DIRECTION), dtype=torch.float32, device=dev)
        self.motor_strength_ratios = torch.ones((n, 12), dtype=torch.float32, device=dev)   # simple_motor.py:52
        self.applied_motor_torques = torch.zeros((n, 12), dtype=torch.float32, device=dev)
        self.observed_motor_torques = torch.zeros((n, 12), dtype=torch.float32, device=dev)
        self._refresh_robot()
        # controller_class(robot, sim.GetTimeSinceReset): core/simulation.py:117
        self._controller_obj = controller_class(self._robot, self.GetTimeSinceReset, **(controller_kwargs or {}))
        self._fused = bool(fuse_motor_model) and hasattr(self._controller_obj, "attach_torque_consumer")
        if self._fused:
            self._controller_obj.attach_torque_consumer(self._robot.GetMotorVelocities, self.motor_strength_ratios,
                                                        self.applied_motor_torques)
        self.reset()

    # ---- the names Simulation exposes (core/simulation.py:60-100,141-142)
    @property
    def controller(self):
        return self._controller_obj

    @property
    def robot(self):
        return self._robot

    @property
    def physics(self):
        return self._physics

    @property
    def env_time_step(self):
        return sim_constants.ACTION_REPEAT * sim_constants.SIMULATION_TIME_STEP

    def GetTimeSinceReset(self):
        return self._time

    def set_strength_ratios(self, ratios):
        """``RobotMotorModel.set_strength_ratios`` (simple_motor.py:54-60) for every env: [12] or [N,12]."""
        self.motor_strength_ratios.copy_(torch.as_tensor(ratios, dtype=torch.float32, device=self.device).expand(self.num_envs, 12))

    def reset(self, env_ids=None):
        """``Simulation.reset`` (:123-127): tick counters to zero, controller re-armed -- for all envs or a subset."""
        sel = slice(None) if env_ids is None else env_ids
        self._physics.reset(env_ids)
        self._step_counter[sel] = 0
        self._time[sel] = 0
        self._refresh_robot()
        self._controller_obj.reset(env_ids)

    def _refresh_robot(self):
        p = self._physics
        self._robot.motor_velocities.copy_(p.joint_velocities * self._direction)
        self._robot.set_sim_state(self._time, p.base_orientation_xyzw, p.base_velocity_world,
                                  p.base_angular_velocity_world, p.joint_angles, p.foot_contacts)

    def ApplyStepAction(self, action):
        """``Simulation.ApplyStepAction`` (:175-179): ACTION_REPEAT x (Robot.ApplyAction -> physics tick)."""
        lib, n = rg.load(), self.num_envs
        P = lambda t: ctypes.c_void_p(t.data_ptr())
        ws = self._controller_obj._robot_ws.ptr
        dt = sim_constants.SIMULATION_TIME_STEP
        for tick in range(sim_constants.ACTION_REPEAT):
            if tick > 0 or not self._fused:
                # Robot.ApplyAction: PD observation -> motor model -> * strength -> * MOTOR_DIRECTION (robot.py:276-307)
                with torch.cuda.device(self.device):
                    rg.check(lib.rg_hybrid_motor_torque_ex(ws, n, P(action), P(self._robot.motor_angles),
                                                           P(self._robot.motor_velocities), P(self.motor_strength_ratios),
                                                           P(self.observed_motor_torques), P(self.applied_motor_torques),
                                                           rg.current_stream_ptr()))
            self._physics.step(self.applied_motor_torques, dt)
            self._step_counter += 1
            torch.mul(self._step_counter, dt, out=self._time)
            self._refresh_robot()
