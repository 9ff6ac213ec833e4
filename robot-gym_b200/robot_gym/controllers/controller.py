"""The controller interface every robot-gym controller implements.

Mirror of robot_gym/controllers/controller.py:4-28 (same method names, same constructor
arguments) so that ``BatchedMPCController`` registers beside ``MPCController`` in
robot_gym/util/cli/mapper.py:7-9 and is constructed the same way
(``controller_class(robot, sim.GetTimeSinceReset)``, robot_gym/core/simulation.py:117).

Batched semantics added by this repo (the reference drives exactly one robot per controller):

* ``robot`` may answer its getters with ``[N, ...]`` CUDA tensors (N independent envs);
* the clock may return a Python float (all envs share it) or an ``[N]`` float64 tensor;
* ``update_controller_params`` takes one command for all envs or an ``[N, 2|3]`` tensor;
* ``get_action`` returns ``[N, 60]`` (``[60]`` numpy for N == 1, what ``ApplyStepAction`` consumes);
* ``reset`` optionally takes the env indices to re-arm.

Which of the five hooks a subclass really needs is unchanged: ``Simulation`` calls
``update_controller_params`` / ``get_action`` / ``reset`` every run and the two UI hooks only when a
PyBullet GUI is attached (robot_gym/core/simulation.py:104-121, playground/playground.py:61-116).
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, Callable


class Controller(ABC):
    """Owns a robot handle and a clock; turns high-level commands into per-motor actions."""

    #: read by Simulation.build_world / ApplyStepAction (core/simulation.py:113,178); subclasses set it
    MOTOR_CONTROL_MODE: Any = None

    def __init__(self, robot: Any, get_time_since_reset: Callable[[], Any]):
        self._robot = robot
        self.get_time_since_reset = get_time_since_reset

    # ------------------------------------------------------------------ command / action
    @abstractmethod
    def update_controller_params(self, params: Any) -> None:
        """Set the high-level command.  MPC: ``(vx, wz)`` or ``(vx, vy, wz)`` before the per-robot
        offsets are added (mpc_controller.py:83-100)."""

    @abstractmethod
    def get_action(self) -> Any:
        """One control step: pull the robot state through its getters, return the motor command in the
        layout ``MOTOR_CONTROL_MODE`` announces."""

    @abstractmethod
    def reset(self, env_ids: Any = None) -> None:
        """Re-arm the controller at the current clock (all envs, or the given subset)."""

    # ------------------------------------------------------------------ GUI hooks (PyBullet debug sliders)
    @abstractmethod
    def setup_ui_params(self, pybullet_client: Any) -> Any:
        """Create the debug sliders of this controller; returns their handles."""

    @abstractmethod
    def read_ui_params(self, pybullet_client: Any, ui: Any) -> Any:
        """Read back the sliders created by ``setup_ui_params`` as a command for ``update_controller_params``."""
