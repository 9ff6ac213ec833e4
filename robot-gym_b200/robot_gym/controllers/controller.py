"""The controller interface every robot-gym controller implements.

Mirror of robot_gym/controllers/controller.py:4-28 (same method names, same constructor
arguments) so that ``BatchedMPCController`` registers beside ``MPCController`` in
robot_gym/util/cli/mapper.py:7-9 and is constructed the same way
(``controller_class(robot, sim.GetTimeSinceReset)``, robot_gym/core/simulation.py:117).
"""
import abc


class Controller(abc.ABC):
    """A controller owns a robot handle and a clock, and turns commands into motor actions."""

    def __init__(self, robot, get_time_since_reset):
        self._robot = robot
        self.get_time_since_reset = get_time_since_reset

    @abc.abstractmethod
    def update_controller_params(self, params):
        """Set the high-level command (for the MPC controller: (vx, wz) or (vx, vy, wz))."""

    @abc.abstractmethod
    def get_action(self):
        """One control step: read the robot state, return the motor command."""

    @abc.abstractmethod
    def setup_ui_params(self, pybullet_client):
        """Create debug sliders; returns their handles."""

    @abc.abstractmethod
    def read_ui_params(self, pybullet_client, ui):
        """Read the debug sliders created by setup_ui_params."""

    @abc.abstractmethod
    def reset(self):
        """Re-arm the controller at the current clock."""
