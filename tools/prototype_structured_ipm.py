"""NumPy prototype of the algorithm the CUDA kernel ``rg_mpc_build_solve`` implements.

Design aid, not product and not oracle: it mirrors the kernel's data flow (closed-form
discretisation, Kronecker-structured Hessian, 6h x 6h Woodbury-reduced Newton system,
Mehrotra predictor-corrector) in float64 so the algebra can be checked against
``oracle/convex_mpc.py`` on a CPU before any CUDA is written.  See DESIGN.md section 3.

Run:  PYTHONPATH=. python tools/prototype_structured_ipm.py
"""
from __future__ import annotations

import math
import sys

import numpy as np
import scipy.linalg


def time_tables(h):
    """c1(j,k) = h - max(j,k);  c2(j,k) = sum_{i=max(j,k)+1..h} (i-j-1/2)(i-k-1/2)."""
    c1 = np.zeros((h, h))
    c2 = np.zeros((h, h))
    for j in range(h):
        for k in range(h):
            m = max(j, k)
            c1[j, k] = h - m
            c2[j, k] = sum((i - j - 0.5) * (i - k - 0.5) for i in range(m + 1, h + 1))
    return c1, c2


def build_structured(params, com_velocity, rpy, angular_velocity, contacts, feet_base,
                     des_pos, des_vel, des_rpy, des_w):
    h, dt, g = params.horizon, params.dt, params.gravity
    w = np.asarray(params.weights, dtype=np.float64)
    l_rho, l_nu = w[0:6], w[6:12]
    roll, pitch, yaw = rpy
    cr, sr, cp, sp, cy, sy = math.cos(roll), math.sin(roll), math.cos(pitch), math.sin(pitch), math.cos(yaw), math.sin(yaw)
    rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    r_feet = rx @ ry @ rz
    r_body = rz @ ry @ rx
    inv_i = r_body @ np.linalg.inv(np.asarray(params.inertia).reshape(3, 3)) @ r_body.T
    tmat = np.array([[cy / cp, sy / cp, 0], [-sy, cy, 0], [cy * sp / cp, sy * sp / cp, 1]])
    feet_w = (r_feet @ np.asarray(feet_base).reshape(-1, 3).T).T
    stance = [i for i, c in enumerate(contacts) if c]
    if stance:
        com_z = abs(sum(feet_w[i, 2] for i in stance) / len(stance))
    else:
        com_z = des_pos[2]
    # B_nu per stance leg: 6x3  = [ I_w^-1 [r]x ; I/m ]
    bnu = []
    for i in stance:
        r = feet_w[i]
        sk = np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]])
        bnu.append(np.vstack([inv_i @ sk, np.eye(3) / params.mass]))
    g6 = np.zeros((6, 6))
    g6[:3, :3] = tmat
    g6[3:, 3:] = np.eye(3)
    m6 = g6.T @ (l_rho[:, None] * g6)
    rho0 = np.array([roll, pitch, yaw, 0, 0, com_z])
    nu0 = np.array([*angular_velocity, *com_velocity])
    grav = np.array([0, 0, 0, 0, 0, -g])
    gt = np.zeros((h, 6))
    e_rho = np.zeros((h + 1, 6))
    e_nu = np.zeros((h + 1, 6))
    for i in range(1, h + 1):
        rho_ref = np.array([des_rpy[0], des_rpy[1], yaw + dt * i * des_w[2], dt * i * des_vel[0], dt * i * des_vel[1], des_pos[2]])
        nu_ref = np.array([0, 0, des_w[2], des_vel[0], des_vel[1], 0])
        e_rho[i] = rho0 + i * dt * (g6 @ nu0) + 0.5 * (i * dt) ** 2 * grav - rho_ref
        e_nu[i] = nu0 + i * dt * grav - nu_ref
    for j in range(h):
        for i in range(j + 1, h + 1):
            gt[j] += 2 * (dt * l_nu * e_nu[i] + dt * dt * (i - j - 0.5) * (g6.T @ (l_rho * e_rho[i])))
    return dict(stance=stance, bnu=bnu, m6=m6, l_nu=l_nu, gt=gt, com_z=com_z)


def solve_structured(params, st, max_iter=30, tol=1e-9, verbose=False):
    h, dt, alpha = params.horizon, params.dt, params.alpha
    stance, bnu, m6, l_nu, gt = st["stance"], st["bnu"], st["m6"], st["l_nu"], st["gt"]
    ns = len(stance)
    nleg = params.num_legs
    out = np.zeros((h, nleg, 3))
    if ns == 0:
        return out, 0
    c1, c2 = time_tables(h)
    kt = 2 * dt * dt * np.kron(c1, np.diag(l_nu)) + 2 * dt ** 4 * np.kron(c2, m6)        # 6h x 6h, index (t, c)
    kt_inv = np.linalg.inv(kt)
    mu_f = np.asarray(params.friction_coeffs, dtype=np.float64)
    gblk = np.array([[-1, 0, mu_f[0]], [1, 0, mu_f[1]], [0, -1, mu_f[2]], [0, 1, mu_f[3]], [0, 0, 1]])
    gfull = np.vstack([gblk, -gblk])                         # 10 x 3 : G f <= hv
    fzmax, fzmin = params.fz_max, params.fz_min
    big_u = (mu_f[0] + 1) * fzmax
    hv = np.array([big_u, big_u, big_u, big_u, fzmax, 0, 0, 0, 0, -fzmin])
    bmat = np.stack(bnu)                                      # ns x 6 x 3
    q = np.einsum("lcd,tc->tld", bmat, gt)                    # h x ns x 3

    def apply_p(u):
        a = np.einsum("lcd,tld->tc", bmat, u).reshape(-1)     # W u  (6h)
        ka = (kt @ a).reshape(h, 6)
        return 2 * alpha * u + np.einsum("lcd,tc->tld", bmat, ka)

    u = np.zeros((h, ns, 3))
    u[:, :, 2] = math.sqrt(fzmin * fzmax)
    s = hv[None, None, :] - np.einsum("rd,tld->tlr", gfull, u)
    qscale = max(1.0, np.abs(q).max())
    lam = 0.1 * qscale / s
    m = s.size
    it = 0
    for it in range(1, max_iter + 1):
        r_d = apply_p(u) + q + np.einsum("rd,tlr->tld", gfull, lam)
        r_p = np.einsum("rd,tld->tlr", gfull, u) + s - hv
        mu = float((s * lam).sum()) / m
        res = max(np.abs(r_d).max() / qscale, np.abs(r_p).max() / fzmax, mu / qscale)
        if verbose:
            print(it, res, mu)
        if res < tol:
            break
        d = lam / s
        ebar = np.einsum("rd,tlr,re->tlde", gfull, d, gfull) + 2 * alpha * np.eye(3)
        lc = np.linalg.cholesky(ebar)                                  # 3x3 lower factors
        lc_inv = np.linalg.inv(lc)                                     # triangular: forward substitution in the kernel
        zmat = np.einsum("tlde,lce->tldc", lc_inv, bmat)               # L^-1 B^T : 3 x 6 per block
        nt = np.einsum("tldc,tldf->tcf", zmat, zmat)                   # h x 6 x 6 (Gram form: PSD by construction)
        psi = kt_inv.copy()
        for t in range(h):
            psi[6 * t:6 * t + 6, 6 * t:6 * t + 6] += nt[t]
        cho = scipy.linalg.cho_factor(psi)

        def phi_solve(b):
            yh = np.einsum("tlde,tle->tld", lc_inv, b)                 # L^-1 b
            tt = np.einsum("tldc,tld->tc", zmat, yh).reshape(-1)       # W Ebar^-1 b
            v = scipy.linalg.cho_solve(cho, tt).reshape(h, 6)
            return np.einsum("tled,tle->tld", lc_inv, yh - np.einsum("tldc,tc->tld", zmat, v))

        def newton(r_c):
            rhs = -r_d - np.einsum("rd,tlr->tld", gfull, (-r_c + lam * r_p) / s)
            dx = phi_solve(rhs)
            ds = -r_p - np.einsum("rd,tld->tlr", gfull, dx)
            dl = (-r_c - lam * ds) / s
            return dx, ds, dl

        def max_step(v, dv):
            neg = dv < 0
            return min(1.0, float(np.min(-v[neg] / dv[neg]))) if np.any(neg) else 1.0

        dx_a, ds_a, dl_a = newton(s * lam)
        a_aff = min(max_step(s, ds_a), max_step(lam, dl_a))
        mu_aff = float(((s + a_aff * ds_a) * (lam + a_aff * dl_a)).sum()) / m
        sigma = (mu_aff / mu) ** 3
        dx, ds, dl = newton(s * lam + ds_a * dl_a - sigma * mu)
        a = min(1.0, 0.99 * min(max_step(s, ds), max_step(lam, dl)))
        u, s, lam = u + a * dx, s + a * ds, lam + a * dl
    for li, leg in enumerate(stance):
        out[:, leg, :] = u[:, li, :]
    return out, it


if __name__ == "__main__":
    sys.path.insert(0, ".")
    from oracle import convex_mpc as cm

    p = cm.MpcParams()
    feet = np.array([[0.2, -0.15, -0.42], [0.2, 0.15, -0.42], [-0.2, -0.15, -0.42], [-0.2, 0.15, -0.42]])
    rng = np.random.default_rng(1)
    worst, its = 0.0, []
    for i in range(60):
        rpy = np.array([rng.uniform(-.2, .2), rng.uniform(-.2, .2), rng.uniform(-.3, .3) if i % 5 == 0 else 0.0])
        v, w = rng.uniform(-.5, .5, 3), rng.uniform(-.5, .5, 3)
        ft = feet + rng.uniform(-.05, .05, (4, 3))
        contacts = [[1, 1, 1, 1], [1, 0, 0, 1], [0, 1, 1, 0], [1, 1, 1, 0], [0, 0, 1, 0]][i % 5]
        args = (v, rpy, w, contacts, ft.ravel(), [0, 0, 0.42], [rng.uniform(0, .35), 0.08, 0], [0, 0, 0], [0, 0, rng.uniform(-.4, .4)])
        f_ref = cm.compute_contact_forces(p, *args)
        st = build_structured(p, *args)
        sol, it = solve_structured(p, st)
        f = -sol.reshape(-1)
        err = np.abs(f - f_ref).max() / max(1.0, np.abs(f_ref).max())
        worst = max(worst, err)
        its.append(it)
    print("max rel err vs oracle", worst, "iters mean/max", np.mean(its), max(its))
