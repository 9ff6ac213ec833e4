"""Prototype (numpy): solve  Psi v = b,  Psi = K^-1 + blkdiag_t(D_t)  WITHOUT forming K^-1 or a 6h x 6h matrix.

K = c1 (x) K1 + c2 (x) K2 is the Hessian, in acceleration space, of the tracking cost
    1/2 a^T K a = sum_{i=1..h} sigma_i^T Q1 sigma_i + pi_i^T Q2 pi_i,   Q1 = K1/2, Q2 = K2/2,
    sigma_i = sum_{j<i} a_j,   pi_i = sum_{j<i} (i-j-1/2) a_j
i.e. of the linear system  x_{i+1} = Phi x_i + Gam a_i,  x = (pi; sigma),  Phi = [[I,I],[0,I]],  Gam = [I/2; I],  x_0 = 0.
With a = K^-1 v the equation reads  a_t = b_t - D_t v_t,  v_t = Gam^T lam_{t+1},  lam_i = 2Q x_i + Phi^T lam_{i+1}:
a two-point boundary value problem that a backward Riccati sweep (lam_i = P_i x_i + p_i) solves in O(h) 6x6 / 12x12
operations.  This file checks the recursion against the dense solve; rg_mpc.cu implements it (riccati_factor / riccati_solve).
"""
import numpy as np

def tables(h):
    j = np.arange(h)
    m = np.maximum.outer(j, j)
    c1 = (h - m).astype(float)
    c2 = np.zeros((h, h))
    for a in range(h):
        for b in range(h):
            i = np.arange(max(a, b) + 1, h + 1)
            c2[a, b] = np.sum((i - a - 0.5) * (i - b - 0.5))
    return c1, c2

def riccati_factor(h, q1, q2, d):
    """d: [h,6,6] PSD.  Returns per stage t (0..h-1): PG_{t+1} (12x6), S_{t+1}, J_t, N_t."""
    q = np.zeros((12, 12)); q[:6, :6] = 2 * q2; q[6:, 6:] = 2 * q1
    phi = np.block([[np.eye(6), np.eye(6)], [np.zeros((6, 6)), np.eye(6)]])
    gam = np.vstack([0.5 * np.eye(6), np.eye(6)])
    p = q.copy()                                   # P_h
    out = [None] * h
    for t in range(h - 1, -1, -1):
        pg = p @ gam                               # = P[:, :6]/2 + P[:, 6:]
        s = gam.T @ pg
        m = np.eye(6) + s @ d[t]
        jm = np.linalg.inv(m)
        n = d[t] @ jm
        out[t] = (pg, s, jm, n)
        if t > 0:
            pp = p - pg @ n @ pg.T
            pp = 0.5 * (pp + pp.T)
            p = q + phi.T @ pp @ phi
    return out

def riccati_solve(h, fac, d, b):
    phi = np.block([[np.eye(6), np.eye(6)], [np.zeros((6, 6)), np.eye(6)]])
    gam = np.vstack([0.5 * np.eye(6), np.eye(6)])
    pvec = np.zeros(12)                            # p_{t+1}, starting with p_h = 0
    r = np.zeros((h, 6))
    for t in range(h - 1, -1, -1):
        pg, s, jm, n = fac[t]
        r[t] = s @ b[t] + gam.T @ pvec
        w = b[t] - n @ r[t]
        pvec = phi.T @ (pvec + pg @ w)
    x = np.zeros(12)
    v = np.zeros((h, 6))
    for t in range(h):
        pg, s, jm, n = fac[t]
        v[t] = jm @ (pg.T @ (phi @ x) + r[t])
        a = b[t] - d[t] @ v[t]
        x = phi @ x + gam @ a
    return v

if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for h in (5, 10, 20):
        c1, c2 = tables(h)
        worst = 0
        for trial in range(20):
            dt = 0.025
            k1 = np.diag(2 * dt**2 * np.array([0.5, 0.5, 0.2, 0.2, 0.2, 0.1]))
            tm = rng.normal(size=(3, 3)) * 0.2 + np.eye(3)
            k2 = np.zeros((6, 6)); k2[:3, :3] = 2 * dt**4 * tm.T @ np.diag([5, 5, 0.2]) @ tm; k2[3:, 3:] = np.diag(2 * dt**4 * np.array([0, 0, 10.0]))
            kk = np.kron(c1, k1) + np.kron(c2, k2)
            d = np.zeros((h, 6, 6))
            for t in range(h):
                nfree = rng.integers(0, 7)          # rank-deficient D_t included (blocks with active rows)
                bz = rng.normal(size=(6, nfree)) * np.array([5, 5, 5, .05, .05, .05])[:, None]
                d[t] = bz @ bz.T / 2e-5
            b = rng.normal(size=(h, 6)) * 1e3
            psi = np.linalg.inv(kk) + np.kron(np.eye(h), np.ones((6, 6))) * 0
            for t in range(h): psi[6*t:6*t+6, 6*t:6*t+6] += d[t]
            v_ref = np.linalg.solve(psi, b.reshape(-1)).reshape(h, 6)
            fac = riccati_factor(h, k1 / 2, k2 / 2, d)
            v = riccati_solve(h, fac, d, b)
            err = np.abs(v - v_ref).max() / np.abs(v_ref).max()
            worst = max(worst, err)
        print(f"h={h}: worst relative error of the Riccati solve vs dense (K^-1 + D)^-1 b: {worst:.2e}  cond(psi) ~ {np.linalg.cond(psi):.1e}")
