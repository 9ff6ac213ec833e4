"""ctypes wrapper of the plain-C CPU oracle (oracle/c/mpc_oracle.c).  TEST INFRASTRUCTURE ONLY.

Used by the GPU parity tests at full size (4096 envs), by ``__graft_entry__.smoke()`` and by
bench.py's CPU-baseline / ``--impl reference`` legs.  PARITY UNPINNED -- see oracle/__init__.py.
kind = "port": there is no compilable reference source for this path, this is the in-repo
restatement of mpc_osqp's dense formulation with a dense interior-point solve.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libmpc_oracle.so")


class RgoParams(ctypes.Structure):
    _fields_ = [("mass", ctypes.c_double), ("inertia", ctypes.c_double * 9), ("num_legs", ctypes.c_int),
                ("horizon", ctypes.c_int), ("dt", ctypes.c_double), ("weights", ctypes.c_double * 13),
                ("alpha", ctypes.c_double), ("mu", ctypes.c_double * 4), ("gravity", ctypes.c_double),
                ("fz_max", ctypes.c_double), ("fz_min", ctypes.c_double)]


def build(force=False):
    src = os.path.join(_HERE, "c", "mpc_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", os.path.join(_HERE, "c"), "-B"], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return LIB_PATH


_lib = None


def load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(LIB_PATH)
        lib.rgo_batch.restype = ctypes.c_long
        lib.rgo_batch.argtypes = [ctypes.POINTER(RgoParams), ctypes.c_int] + [ctypes.c_void_p] * 6 + [ctypes.c_double] + \
                                 [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        lib.rgo_compute_contact_forces.restype = ctypes.c_int
        lib.rgo_compute_contact_forces.argtypes = [ctypes.POINTER(RgoParams)] + [ctypes.c_void_p] * 12
        lib.rgo_scratch_doubles.restype = ctypes.c_size_t
        lib.rgo_scratch_doubles.argtypes = [ctypes.c_int, ctypes.c_int]
        _lib = lib
    return _lib


def params_from(mpc_params) -> RgoParams:
    """From oracle.convex_mpc.MpcParams."""
    p = RgoParams()
    p.mass = mpc_params.mass
    for i, v in enumerate(mpc_params.inertia):
        p.inertia[i] = v
    p.num_legs, p.horizon, p.dt, p.alpha = mpc_params.num_legs, mpc_params.horizon, mpc_params.dt, mpc_params.alpha
    for i, v in enumerate(mpc_params.weights):
        p.weights[i] = v
    for i, v in enumerate(mpc_params.friction_coeffs):
        p.mu[i] = v
    p.gravity, p.fz_max, p.fz_min = mpc_params.gravity, mpc_params.fz_max, mpc_params.fz_min
    return p


def solve_batch(mpc_params, states, desired_height, n_threads=1, want_horizon=False):
    """First-step forces [N,12] (and [N,h,12] if asked) for SyntheticStates-like host arrays."""
    lib = load()
    p = params_from(mpc_params)
    n = len(states.base_rpy)
    out = np.zeros((n, 12), dtype=np.float32)
    hout = np.zeros((n, mpc_params.horizon, 12), dtype=np.float32) if want_horizon else None
    arrs = [np.ascontiguousarray(a) for a in (states.com_velocity_body, states.base_rpy, states.base_rpy_rate,
                                              states.planned_contacts, states.foot_positions_base, states.command)]
    assert arrs[3].dtype == np.uint8 and all(a.dtype == np.float32 for a in arrs[:3] + arrs[4:])
    iters = lib.rgo_batch(ctypes.byref(p), n, *[a.ctypes.data_as(ctypes.c_void_p) for a in arrs], float(desired_height),
                          out.ctypes.data_as(ctypes.c_void_p),
                          hout.ctypes.data_as(ctypes.c_void_p) if hout is not None else None, int(n_threads))
    return out, hout, int(iters)


def compute_contact_forces(params, com_velocity, rpy, angular_velocity, contacts, foot_positions_base,
                           desired_com_position, desired_com_velocity, desired_rpy, desired_angular_velocity,
                           com_position=None):
    """One env through the C port, float64 in and out: same signature and return value (the NEGATED solution,
    3*k*h doubles) as ``oracle.convex_mpc.compute_contact_forces``, about 100x faster.  Lets the restated
    LocomotionController (oracle/locomotion.py) be stepped over hundreds of envs in a test."""
    lib = load()
    p = params_from(params)
    d = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    cv, r3, w3, ft = d(com_velocity), d(rpy), d(angular_velocity), d(foot_positions_base).reshape(-1)
    ct = np.ascontiguousarray(np.asarray(contacts).astype(np.int32))
    dp, dv, dr, dw = d(desired_com_position), d(desired_com_velocity), d(desired_rpy), d(desired_angular_velocity)
    cp = d(com_position) if com_position is not None and len(com_position) == 3 else None
    out = np.zeros(3 * params.num_legs * params.horizon, dtype=np.float64)
    scratch = np.zeros(int(lib.rgo_scratch_doubles(params.horizon, params.num_legs)), dtype=np.float64)
    lib.rgo_compute_contact_forces(ctypes.byref(p), ptr(cv), ptr(r3), ptr(w3), ptr(ct), ptr(ft), ptr(dp), ptr(dv), ptr(dr),
                                   ptr(dw), ptr(cp) if cp is not None else None, ptr(out), ptr(scratch))
    return out
