#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mpc.py tests/test_gpu_parity_full.py -m gpu -q -x -k "riccati or frozen or config4 or kkt" > gpurun_out/r02i_pytest.log 2>&1
tail -3 gpurun_out/r02i_pytest.log
export RG_PERF_NO_ALLSTANCE=1
RG_PERF_H=20 timeout 300 python tools/perf_mpc.py 65536 > gpurun_out/r02i_perf_h20.log 2>&1
RG_PERF_H=20 RG_PERF_GAIT=bound timeout 300 python tools/perf_mpc.py 65536 >> gpurun_out/r02i_perf_h20.log 2>&1
cat gpurun_out/r02i_perf_h20.log
timeout 300 python tools/gpu/find_unverified.py > gpurun_out/r02i_unverified.log 2>&1
cat gpurun_out/r02i_unverified.log
