"""Name -> class registries, the counterpart of robot_gym/util/cli/mapper.py:7-19.

The reference's ``CONTROLLERS`` has one entry, ``'mpc': mpc_controller.MPCController``; the batched controller
registers beside it as ``'mpc_cuda'`` (the one-line change a reference maintainer adds is shown in INTEGRATION.md)."""
from robot_gym.controllers.mpc.batched_mpc_controller import BatchedMPCController
from robot_gym.model.robots.descriptions import GHOST, K3LSO

CONTROLLERS = {
    "mpc_cuda": BatchedMPCController,
}

ROBOTS = {
    "ghost": GHOST,
    "k3lso": K3LSO,
}
