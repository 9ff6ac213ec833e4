#!/bin/bash
# ncu --set full captures: lean kernel h = 20 (Riccati) and h = 10 (Cholesky)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export RG_PERF_NO_ALLSTANCE=1
RG_PERF_H=20 ncu --set full --clock-control none --import-source on -k regex:mpc_solve -s 3 -c 1 -f -o gpurun_out/r02h_prof_h20 python tools/perf_mpc.py 16384 > gpurun_out/r02h_ncu_h20.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mpc_solve -s 3 -c 1 -f -o gpurun_out/r02h_prof_h10_4096 python tools/perf_mpc.py 4096 > gpurun_out/r02h_ncu_h10.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
