#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02k_pytest.log 2>&1
tail -3 gpurun_out/r02k_pytest.log
export RG_PERF_NO_ALLSTANCE=1
timeout 300 python tools/perf_mpc.py 4096 65536 > gpurun_out/r02k_perf.log 2>&1
RG_PERF_H=20 timeout 300 python tools/perf_mpc.py 65536 >> gpurun_out/r02k_perf.log 2>&1
RG_PERF_H=5 timeout 300 python tools/perf_mpc.py 65536 >> gpurun_out/r02k_perf.log 2>&1
RG_PERF_GAIT=bound timeout 300 python tools/perf_mpc.py 65536 >> gpurun_out/r02k_perf.log 2>&1
cat gpurun_out/r02k_perf.log
timeout 900 python bench.py > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err
cut -c1-300 gpurun_out/r02k_bench.json
