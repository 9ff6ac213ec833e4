#!/bin/bash
# ncu --set full capture of the lean kernel at h = 5 (one warp per env)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export RG_PERF_NO_ALLSTANCE=1
RG_PERF_H=5 ncu --set full --clock-control none --import-source on -k regex:mpc_solve -s 3 -c 1 -f -o gpurun_out/r02q_prof_h5 python tools/perf_mpc.py 65536 > gpurun_out/r02q_ncu_h5.log 2>&1
tail -3 gpurun_out/r02q_ncu_h5.log
