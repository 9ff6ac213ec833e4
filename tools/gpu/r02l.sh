#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export RG_PERF_NO_ALLSTANCE=1
for lib in default librg_mb9 librg_mb9arr librg_mb8arr; do
  if [ $lib = default ]; then unset RG_CUDA_LIB; else export RG_CUDA_LIB=$PWD/ab/$lib.so; fi
  timeout 300 python tools/perf_mpc.py 4096 65536 > gpurun_out/r02l_perf_$lib.log 2>&1
  echo $lib; cat gpurun_out/r02l_perf_$lib.log
done
