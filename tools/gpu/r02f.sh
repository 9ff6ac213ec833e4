#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_env_shim.py -m gpu -q -x > gpurun_out/r02f_pytest_shim.log 2>&1
tail -5 gpurun_out/r02f_pytest_shim.log
export RG_PERF_NO_ALLSTANCE=1
for lib in librg_h20mb5 librg_h20mb6; do
  RG_CUDA_LIB=$PWD/ab/$lib.so RG_PERF_H=20 timeout 300 python tools/perf_mpc.py 65536 > gpurun_out/r02f_perf_$lib.log 2>&1
  RG_CUDA_LIB=$PWD/ab/$lib.so RG_PERF_H=20 RG_PERF_GAIT=bound timeout 300 python tools/perf_mpc.py 65536 >> gpurun_out/r02f_perf_$lib.log 2>&1
  echo $lib; cat gpurun_out/r02f_perf_$lib.log
done
timeout 900 python tools/config1_substitute.py --steps 1000 --numpy-steps 60 --out gpurun_out/r02_config1_substitute.json > gpurun_out/r02f_config1.log 2>&1
tail -40 gpurun_out/r02f_config1.log
