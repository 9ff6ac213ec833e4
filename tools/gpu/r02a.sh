#!/bin/bash
# Round-2 GPU pass A: full GPU suite, Cholesky microbenchmark, A/B of the lean two-kernel solve, bench line.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/r02a_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02a_pytest.log
timeout 300 tools/microbench/chol_bench > gpurun_out/r02a_chol_bench.log 2>&1
export RG_PERF_NO_ALLSTANCE=1
timeout 300 python tools/perf_mpc.py 4096 65536 > gpurun_out/r02a_perf_default.log 2>&1
RG_PERF_PARAMS=two_kernel_solve=0 timeout 300 python tools/perf_mpc.py 4096 65536 > gpurun_out/r02a_perf_onekernel.log 2>&1
RG_CUDA_LIB=$PWD/ab/librg_mb9.so timeout 300 python tools/perf_mpc.py 4096 65536 > gpurun_out/r02a_perf_mb9.log 2>&1
RG_PERF_GAIT=bound timeout 300 python tools/perf_mpc.py 16384 > gpurun_out/r02a_perf_bound.log 2>&1
RG_PERF_GAIT=bound RG_PERF_PARAMS=two_kernel_solve=0 timeout 300 python tools/perf_mpc.py 16384 > gpurun_out/r02a_perf_bound_onekernel.log 2>&1
RG_PERF_H=20 timeout 300 python tools/perf_mpc.py 16384 > gpurun_out/r02a_perf_h20.log 2>&1
timeout 900 python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
tail -3 gpurun_out/r02a_pytest.log
cat gpurun_out/r02a_perf_*.log
