#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export RG_PERF_NO_ALLSTANCE=1
for lib in default librg_h20mb6 librg_h20mb4; do
  if [ $lib = default ]; then unset RG_CUDA_LIB; else export RG_CUDA_LIB=$PWD/ab/$lib.so; fi
  RG_PERF_H=20 timeout 300 python tools/perf_mpc.py 65536 > gpurun_out/r02n_perf_$lib.log 2>&1
  RG_PERF_H=20 RG_PERF_GAIT=bound timeout 300 python tools/perf_mpc.py 65536 >> gpurun_out/r02n_perf_$lib.log 2>&1
  echo $lib; cat gpurun_out/r02n_perf_$lib.log
done
unset RG_CUDA_LIB
timeout 600 python -m pytest tests/test_gpu_mpc.py tests/test_gpu_parity_full.py -m gpu -q -x -k "frozen or config4 or riccati or parameter_variants" 2>&1 | tail -3
