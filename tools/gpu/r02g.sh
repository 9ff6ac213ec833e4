#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export RG_PERF_NO_ALLSTANCE=1
for lib in default librg_warp0 librg_h20mb6; do
  if [ $lib = default ]; then unset RG_CUDA_LIB; else export RG_CUDA_LIB=$PWD/ab/$lib.so; fi
  RG_PERF_H=20 timeout 300 python tools/perf_mpc.py 65536 > gpurun_out/r02g_perf_$lib.log 2>&1
  if [ $lib != librg_h20mb6 ]; then timeout 300 python tools/perf_mpc.py 4096 65536 >> gpurun_out/r02g_perf_$lib.log 2>&1; fi
  echo $lib; cat gpurun_out/r02g_perf_$lib.log
done
unset RG_CUDA_LIB
timeout 600 python -m pytest tests/test_gpu_env_shim.py tests/test_gpu_mpc.py -m gpu -q -x > gpurun_out/r02g_pytest.log 2>&1
tail -5 gpurun_out/r02g_pytest.log
