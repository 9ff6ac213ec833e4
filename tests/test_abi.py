"""The C-ABI shared library: loads, exports every symbol include/rg_cuda.h declares, and its
host-only entry points validate arguments -- no compute call is made (this runs without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from robot_gym import cuda as rg

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(REPO, "include", "rg_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(rg_lib):
    declared = _declared_functions()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(rg_lib, name), f"{name} declared in rg_cuda.h but not exported"
    assert sorted(rg.EXPORTED_SYMBOLS) == declared
    assert b"sm_100a" in rg_lib.rg_version()


def test_library_is_built_for_sm_100a_only(rg_lib):
    import shutil, subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "--list-elf", rg.library_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_struct_layouts_match_header_sizes(rg_lib):
    # sizes computed from the header's field lists (8-byte aligned doubles, 4-byte ints)
    assert ctypes.sizeof(rg.MpcParams) == 8 + 72 + 4 + 4 + 8 + 104 + 8 + 32 + 8 + 8 + 8 + 8 + 8 + 4 + 4 + 4 + 4 + 4 + 4
    assert ctypes.sizeof(rg.LegChain) == 8 * (9 + 27 + 9 + 3 + 2)
    assert ctypes.sizeof(rg.ControllerState) == 8 * 32
    assert ctypes.sizeof(rg.MpcIo) == 8 * 12 + 8


def test_default_params_follow_motion_imitation_defaults(rg_lib):
    p = rg.default_mpc_params(190 / 9.8, (0.07335, 0, 0, 0, 0.25068, 0, 0, 0, 0.25447), 0.42, 10)
    assert p.horizon == 10 and p.num_legs == 4 and p.dt == 0.025 and p.alpha == 1e-5
    assert list(p.weights) == [5, 5, 0.2, 0, 0, 10, 0.5, 0.5, 0.2, 0.2, 0.2, 0.1, 0]
    assert list(p.friction_coeffs) == [0.45] * 4
    assert p.fz_max == pytest.approx(1900.0) and p.fz_min == pytest.approx(19.0)
    assert p.desired_body_height == 0.42
    assert p.cold_start_rounds == 12 and p.cold_start_max_violations == 0 and p.max_polish_rounds == 3


def test_error_codes_and_messages(rg_lib):
    n = ctypes.c_size_t()
    assert rg_lib.rg_workspace_bytes(0, 10, 4, ctypes.byref(n)) == 0 and n.value >= 3000
    assert rg_lib.rg_workspace_bytes(0, 7, 4, ctypes.byref(n)) == -2          # RG_ERR_UNSUPPORTED
    assert b"horizon" in rg_lib.rg_last_error()
    assert rg_lib.rg_workspace_bytes(0, 10, 6, ctypes.byref(n)) == -2
    assert rg_lib.rg_workspace_bytes(0, 10, 4, None) == -1                    # RG_ERR_BAD_ARG
    assert rg_lib.rg_mpc_setup(None, None, 0, None) == -1
    p = rg.default_mpc_params(19.4, (0.07, 0, 0, 0, 0.25, 0, 0, 0, 0.25), 0.42, 10)
    dummy = ctypes.create_string_buffer(16)
    assert rg_lib.rg_mpc_setup(ctypes.byref(p), dummy, 16, None) == -4        # RG_ERR_WORKSPACE (too small)
    big = ctypes.create_string_buffer(131072)
    p.weights[9] = 0.0; p.weights[3] = 0.0                                    # x channel unpenalised
    assert rg_lib.rg_mpc_setup(ctypes.byref(p), big, 131072, None) == -3        # RG_ERR_SINGULAR
    assert b"unpenalised" in rg_lib.rg_last_error()
    p.weights[9] = 0.2; p.friction_coeffs[2] = 0.0
    assert rg_lib.rg_mpc_setup(ctypes.byref(p), big, 131072, None) == -1
    assert rg_lib.rg_mpc_build_solve(None, 4, *([None] * 10), None) == -1
    assert rg_lib.rg_gait_step(None, 4, None, None, None, None, None, None) == -1
    with pytest.raises(rg.RgCudaError) as err:
        rg.check(rg_lib.rg_leg_ik(None, 1, None, None, None, None))
    assert err.value.code == -1


def test_wrappers_refuse_cpu_tensors(rg_lib):
    import torch
    with pytest.raises(TypeError, match="CUDA tensors only"):
        rg._ptr(torch.zeros(4, 3), torch.float32, (3,))


def test_ik_calibration_is_host_only_and_picks_the_nominal_branch(rg_lib):
    from robot_gym.controllers.mpc.batched_kinematics import robot_params_from_description
    from robot_gym.model.robots.descriptions import GHOST, K3LSO
    for desc in (GHOST, K3LSO):
        p = robot_params_from_description(desc)
        assert all(abs(p.legs[l].ik_sign_knee) == 1.0 and abs(p.legs[l].ik_sign_hip) == 1.0 for l in range(4))
        assert p.velocity_window == 20 and p.foot_clearance == 0.01 and p.contact_detection_phase_threshold == 0.1
        assert [p.initial_leg_state[l] for l in range(4)] == [0, 1, 1, 0]
    # a pose the chain cannot reach from any branch is reported, not guessed
    p = robot_params_from_description(GHOST)
    p.legs[0].axis[1][0], p.legs[0].axis[1][1] = 1.0, 0.0       # upper axis parallel to the hip axis
    arr = (ctypes.c_double * 12)(*GHOST.GetConstants().INIT_MOTOR_ANGLES)
    assert rg_lib.rg_robot_calibrate_ik(ctypes.byref(p), arr) == -2
    assert b"closed-form IK unsupported" in rg_lib.rg_last_error()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from robot_gym.cuda import build
    monkeypatch.setattr(rg, "_lib", None)
    monkeypatch.setattr(build, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rg.load()
