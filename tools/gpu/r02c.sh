#!/bin/bash
# Round-2 GPU pass C: the Riccati linear algebra in the solve kernel -- self-test, full suite, A/B against the Cholesky build.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mpc.py -m gpu -q -x -k "riccati or frozen" > gpurun_out/r02c_pytest_first.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02c_pytest_first.log
tail -5 gpurun_out/r02c_pytest_first.log
export RG_PERF_NO_ALLSTANCE=1
timeout 300 python tools/perf_mpc.py 4096 65536 > gpurun_out/r02c_perf_riccati.log 2>&1
RG_CUDA_LIB=$PWD/ab/librg_chol.so timeout 300 python tools/perf_mpc.py 4096 65536 > gpurun_out/r02c_perf_chol.log 2>&1
RG_PERF_H=20 timeout 300 python tools/perf_mpc.py 16384 65536 > gpurun_out/r02c_perf_riccati_h20.log 2>&1
RG_PERF_H=5 timeout 300 python tools/perf_mpc.py 65536 > gpurun_out/r02c_perf_riccati_h5.log 2>&1
RG_PERF_GAIT=bound timeout 300 python tools/perf_mpc.py 65536 > gpurun_out/r02c_perf_riccati_bound.log 2>&1
cat gpurun_out/r02c_perf_*.log
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02c_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02c_pytest.log
tail -25 gpurun_out/r02c_pytest.log
timeout 300 python tools/gpu/dbg_control_rounds.py > gpurun_out/r02c_dbg_rounds.log 2>&1
cat gpurun_out/r02c_dbg_rounds.log
