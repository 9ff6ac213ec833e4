"""Simulation timing constants the control loop depends on (robot_gym/core/sim_constants.py:7,11, values unchanged)."""
ACTION_REPEAT = 10                 # physics ticks per control step
SIMULATION_TIME_STEP = 0.001       # seconds per physics tick  ->  100 Hz controller
