"""Vectorised counterpart of ``RobotGymEnv`` (robot_gym/gym/robot_gym_env.py): N envs stepped in one process.

The reference steps ONE env per ``RobotGymEnv`` (``step``: update_controller_params -> get_action -> ApplyStepAction,
:117-129) and gets "batched envs" for PPO by running one OS process per env behind pipes
(agents/ppo/tools/wrappers.py:294-457).  ``BatchedRobotGymEnv`` keeps the same step contract with ``[N, ...]`` CUDA
tensors: ``step(actions[N,2|3]) -> (observation[N,D], reward[N], done[N], info)``; the falling test
(``is_falling``: no foot in contact, :155-165) is evaluated per env, and envs that are done are reset individually
(the reference rebuilds the whole simulation, playground.py:119-121).  Task logic -- rewards, observations beyond the
default proprioceptive vector, GoEnv's planner -- is out of scope: ``reward`` / ``get_observation`` are the override
points, as in the reference's abstract methods (:62-76)."""
from __future__ import annotations

import torch

from robot_gym.controllers.mpc.batched_kinematics import BatchedKinematics, robot_params_from_description
from robot_gym.core.batched_simulation import BatchedSimulation, SyntheticPhysics
from robot_gym import cuda as rg


class BatchedRobotGymEnv:
    metadata = {"render.modes": [], "video.frames_per_second": 100}          # robot_gym_env.py:16 (100 Hz control)

    def __init__(self, description, num_envs, controller_class=None, physics=None, device="cuda", auto_reset=True,
                 controller_kwargs=None, fuse_motor_model=True):
        """``physics``: a ``BatchedPhysics`` (default: the synthetic joint integrator); ``controller_class``: a key of
        ``robot_gym.util.cli.mapper.CONTROLLERS`` or a class (default ``'mpc_cuda'``)."""
        from robot_gym.util.cli import mapper
        if controller_class is None:
            controller_class = "mpc_cuda"
        if isinstance(controller_class, str):
            controller_class = mapper.CONTROLLERS[controller_class]
        self.device = torch.device(device)
        self.num_envs = int(num_envs)
        if physics is None:
            kin = BatchedKinematics(rg.RobotWorkspace(robot_params_from_description(description), device=self.device), self.device)
            physics = SyntheticPhysics(description, self.num_envs, kin, device=self.device)
        self._simulation = BatchedSimulation(description, physics, controller_class, device=self.device,
                                             controller_kwargs=controller_kwargs, fuse_motor_model=fuse_motor_model)
        self._auto_reset = bool(auto_reset)
        self.last_action = None
        self.episode_steps = torch.zeros(self.num_envs, dtype=torch.int64, device=self.device)

    @property
    def simulation(self):
        return self._simulation

    # ---- override points (abstract in the reference: robot_gym_env.py:62-76)
    def reward(self):
        return torch.zeros(self.num_envs, dtype=torch.float32, device=self.device)

    def get_observation(self):
        """Default proprioceptive observation [N, 37]: rpy, body-frame angular velocity, motor angles, motor
        velocities, body-frame COM velocity estimate, foot contacts."""
        r, c = self._simulation.robot, self._simulation.controller
        return torch.cat([r.GetBaseRollPitchYaw(), r.GetBaseRollPitchYawRate(), r.GetMotorAngles(), r.GetMotorVelocities(),
                          c.com_velocity_body, r.GetFootContacts().to(torch.float32)], dim=1)

    # ---- gym surface
    def reset(self, env_ids=None):
        self._simulation.reset(env_ids)
        self.episode_steps[slice(None) if env_ids is None else env_ids] = 0
        return self.get_observation()

    def step(self, action):
        """``RobotGymEnv.step`` (:117-129) for N envs.  ``action``: [N,2] (vx, wz) or [N,3] (vx, vy, wz) tensor, or a
        2-/3-tuple broadcast to every env (mpc_controller.py:83-88)."""
        sim = self._simulation
        sim.controller.update_controller_params(action)
        hybrid = sim.controller.get_action()
        if not isinstance(hybrid, torch.Tensor):                     # N == 1 drop-in mode returns the [60] numpy array
            hybrid = sim.controller.action
        self.last_action = hybrid
        sim.ApplyStepAction(hybrid)
        self.episode_steps += 1
        observation = self.get_observation()
        reward = self.reward()
        done, info = self.termination()
        if self._auto_reset:
            ids = torch.nonzero(done, as_tuple=False).flatten()
            if ids.numel():                                           # one tiny host sync per step: which envs fell
                info["reset_env_ids"] = ids
                self.reset(ids)
        return observation, reward, done, info

    def is_falling(self):
        """[N] bool: no foot in contact (robot_gym_env.py:155-165)."""
        return ~(self._simulation.robot.GetFootContacts() != 0).any(dim=1)

    def termination(self):
        return self.is_falling(), {}

    def close(self):
        pass
