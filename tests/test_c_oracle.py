"""The plain-C oracle (oracle/c/mpc_oracle.c: dense reference-style build, Pade matrix exponential,
dense interior point) pinned against the numpy oracle and its frozen goldens.  It is the checker of
the full-size GPU parity tests and the CPU baseline bench.py times."""
import os

import numpy as np
import pytest

from oracle import c_oracle, convex_mpc as cm
from robot_gym.model.robots.descriptions import GHOST, with_gait
from robot_gym.util import synthetic


@pytest.mark.parametrize("horizon", [10, 5, 20])
def test_c_oracle_matches_frozen_numpy_goldens(golden_dir, horizon):
    g = np.load(os.path.join(golden_dir, "mpc_oracle_golden.npz"))
    n = int(g[f"mpc_h{horizon}_n"])
    st = synthetic.make_states(n, GHOST, seed=synthetic.SEED + horizon)
    f, hf, iters = c_oracle.solve_batch(cm.MpcParams(horizon=horizon), st, 0.42, n_threads=2, want_horizon=True)
    ref = g[f"mpc_h{horizon}_forces"]
    err = max(np.abs(hf[i].reshape(-1) - ref[i]).max() / max(1.0, np.abs(ref[i]).max()) for i in range(n))
    assert err < 5e-5, err            # interior point without polish: ~1e-5 in the alpha-directions
    np.testing.assert_array_equal(f, hf[:, 0, :])
    assert 4 <= iters / n <= 25


def test_c_oracle_thread_count_does_not_change_results():
    st = synthetic.make_states(64, GHOST, seed=2)
    a, _, _ = c_oracle.solve_batch(cm.MpcParams(), st, 0.42, n_threads=1)
    b, _, _ = c_oracle.solve_batch(cm.MpcParams(), st, 0.42, n_threads=4)
    np.testing.assert_array_equal(a, b)


def test_c_oracle_swing_legs_zero_and_feasible():
    desc = with_gait(GHOST, "pace")
    st = synthetic.make_states(48, desc, schedule_ctrl=desc.GetCtrlConstants(), seed=9)
    p = cm.MpcParams()
    f, _, _ = c_oracle.solve_batch(p, st, 0.42, n_threads=2)
    grf = -f.reshape(-1, 4, 3).astype(np.float64)
    swing = st.planned_contacts == 0
    assert np.all(grf[swing] == 0)
    fz = grf[~swing][:, 2]
    assert fz.min() >= p.fz_min - 1e-3 and fz.max() <= p.fz_max + 1e-3
    assert np.all(np.abs(grf[~swing][:, :2]) <= 0.45 * fz[:, None] + 1e-3)


def test_c_oracle_matrix_exponential_agrees_with_closed_form_pipeline():
    """Same inputs through the numpy oracle (scipy expm) and the C port (own Pade expm)."""
    st = synthetic.make_states(6, GHOST, seed=5)
    p = cm.MpcParams()
    f, _, _ = c_oracle.solve_batch(p, st, 0.42)
    for i in range(6):
        ref = cm.compute_contact_forces(p, st.com_velocity_body[i].astype(np.float64), st.base_rpy[i].astype(np.float64),
                                        st.base_rpy_rate[i].astype(np.float64), st.planned_contacts[i],
                                        st.foot_positions_base[i].astype(np.float64), [0, 0, 0.42],
                                        [float(st.command[i, 0]), float(st.command[i, 1]), 0.0], [0, 0, 0],
                                        [0, 0, float(st.command[i, 2])])
        assert np.abs(f[i] - ref[:12]).max() < 5e-5 * max(1.0, np.abs(ref[:12]).max())
