#!/bin/bash
# ncu --set full capture of the lean h = 10 kernel: tools/gpu/ncu10.sh <tag> [n_env]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export RG_PERF_NO_ALLSTANCE=1
ncu --set full --clock-control none --import-source on -k regex:mpc_solve -s 3 -c 1 -f -o gpurun_out/$1_prof_h10 python tools/perf_mpc.py ${2:-4096} > gpurun_out/$1_ncu_h10.log 2>&1
tail -2 gpurun_out/$1_ncu_h10.log
