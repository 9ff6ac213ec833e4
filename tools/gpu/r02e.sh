#!/bin/bash
# Round-2 GPU pass E: Riccati linear algebra for h = 20 (dense Cholesky kept for h = 5, 10)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mpc.py -m gpu -q -x -k "riccati or frozen" > gpurun_out/r02e_pytest_first.log 2>&1
tail -3 gpurun_out/r02e_pytest_first.log
export RG_PERF_NO_ALLSTANCE=1
RG_PERF_H=20 timeout 300 python tools/perf_mpc.py 16384 65536 > gpurun_out/r02e_perf_h20.log 2>&1
RG_PERF_H=20 RG_PERF_GAIT=pace timeout 300 python tools/perf_mpc.py 65536 > gpurun_out/r02e_perf_h20_pace.log 2>&1
RG_PERF_H=20 RG_PERF_GAIT=bound timeout 300 python tools/perf_mpc.py 65536 > gpurun_out/r02e_perf_h20_bound.log 2>&1
timeout 300 python tools/perf_mpc.py 4096 65536 > gpurun_out/r02e_perf_h10.log 2>&1
cat gpurun_out/r02e_perf_*.log
RG_TRACE_H=20 RG_TRACE_SOLO=1 RG_TRACE_N=4096 timeout 300 python tools/trace_mpc.py 0 3 > gpurun_out/r02e_trace_h20.log 2>&1
cat gpurun_out/r02e_trace_h20.log
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02e_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02e_pytest.log
tail -12 gpurun_out/r02e_pytest.log
