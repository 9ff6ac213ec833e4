import os, sys, ctypes
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "robot-gym_b200")); sys.path.insert(0, REPO)
import numpy as np, torch
from robot_gym import cuda as rg
lib = rg.load()
for h in (10, 5, 20):
    n = 6*h
    rng = np.random.default_rng(h)
    M = rng.standard_normal((n, n)); A = M @ M.T + n*np.eye(n); b = rng.standard_normal(n)
    a_d = torch.from_numpy(A).cuda(); b_d = torch.from_numpy(b).cuda(); x_d = torch.zeros(n, dtype=torch.float64, device="cuda"); l_d = torch.zeros((n,n), dtype=torch.float64, device="cuda")
    p = ctypes.c_void_p
    rc = lib.rg_debug_chol_solve(h, p(a_d.data_ptr()), p(b_d.data_ptr()), p(x_d.data_ptr()), p(l_d.data_ptr()), None)
    torch.cuda.synchronize()
    L = np.linalg.cholesky(A); x = np.linalg.solve(A, b)
    lo = l_d.cpu().numpy(); xo = x_d.cpu().numpy()
    print("h", h, "rc", rc, "L err", np.abs(np.tril(lo)-L).max(), "x err", np.abs(xo-x).max())
    if np.abs(np.tril(lo)-L).max() > 1e-9:
        bad = np.argwhere(np.abs(np.tril(lo)-L) > 1e-9); print("   first bad L entries", bad[:10].tolist())
    else:
        # check sweeps separately
        y = np.linalg.solve(L, b); 
        print("   fwd+bwd ok?", np.abs(xo-x).max() < 1e-9)
