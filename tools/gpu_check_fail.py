"""GPU bring-up: find envs whose polish did not verify, compare them with the oracle, dump indices."""
import os, sys, json
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "robot-gym_b200")); sys.path.insert(0, REPO)
import numpy as np, torch
from robot_gym import cuda as rg
from robot_gym.model.robots.descriptions import GHOST
from robot_gym.util import synthetic
from oracle import convex_mpc as cm

def run(n, horizon, weights=None, tag=""):
    ctrl = GHOST.GetCtrlConstants()
    p = rg.default_mpc_params(ctrl.MPC_BODY_MASS, ctrl.MPC_BODY_INERTIA, ctrl.MPC_BODY_HEIGHT, horizon)
    if weights is not None:
        for i, w in enumerate(weights): p.weights[i] = w
    ws = rg.MpcWorkspace(p)
    st = synthetic.make_states(n, GHOST)
    t = lambda a: torch.from_numpy(a).to("cuda")
    args = (t(st.com_velocity_body), t(st.base_rpy), t(st.base_rpy_rate), t(st.planned_contacts), t(st.foot_positions_base), t(st.command))
    f, hf, info = rg.mpc_build_solve(ws, *args)
    f = f.cpu().numpy(); info = info.cpu().numpy()
    bad = np.flatnonzero((info[:, 2] & 1) == 0)
    print(tag, "n", n, "h", horizon, "unpolished", len(bad), "contact patterns:", np.unique(st.planned_contacts[bad], axis=0, return_counts=True))
    mp = cm.MpcParams(horizon=horizon)
    if weights is not None: mp.weights = tuple(weights)
    out = []
    for i in bad[:40]:
        ref, oinfo = cm.compute_contact_forces(mp, st.com_velocity_body[i].astype(np.float64), st.base_rpy[i].astype(np.float64),
            st.base_rpy_rate[i].astype(np.float64), st.planned_contacts[i], st.foot_positions_base[i].astype(np.float64),
            [0, 0, ctrl.MPC_BODY_HEIGHT], [st.command[i,0], st.command[i,1], 0.0], [0,0,0], [0,0,float(st.command[i,2])], return_info=True)
        err = np.abs(f[i] - ref[:12]).max() / max(1.0, np.abs(ref[:12]).max())
        print("  env", int(i), st.planned_contacts[i], "info", info[i], "err %.3e" % err, "oracle iters", oinfo["iters"], "polished", oinfo.get("polished"))
        out.append(dict(env=int(i), info=info[i].tolist(), err=float(err)))
    return out

res = {}
res["h10"] = run(65536, 10, tag="default")
res["h5"] = run(4096, 5, tag="h5")
res["w2"] = run(4096, 10, weights=(5,5,0.2,0,0,10,0.,0.,1.,1.,1.,0.,0), tag="weights2")
json.dump(res, open(os.path.join(REPO, "gpurun_out", "fail.json"), "w"))
