#!/bin/bash
# Round-2 GPU pass B: env-shim tests, control-step round count investigation, bound/pace after the fallback grid fix.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_env_shim.py tests/test_gpu_parity_full.py -m gpu -q -x --durations=5 > gpurun_out/r02b_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02b_pytest.log
timeout 300 python tools/gpu/dbg_control_rounds.py > gpurun_out/r02b_dbg_rounds.log 2>&1
export RG_PERF_NO_ALLSTANCE=1
RG_PERF_GAIT=bound timeout 300 python tools/perf_mpc.py 16384 65536 > gpurun_out/r02b_perf_bound.log 2>&1
RG_PERF_GAIT=pace timeout 300 python tools/perf_mpc.py 65536 > gpurun_out/r02b_perf_pace.log 2>&1
tail -15 gpurun_out/r02b_pytest.log; cat gpurun_out/r02b_dbg_rounds.log gpurun_out/r02b_perf_bound.log gpurun_out/r02b_perf_pace.log
