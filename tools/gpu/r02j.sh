#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mpc.py tests/test_gpu_parity_full.py -m gpu -q -x > gpurun_out/r02j_pytest.log 2>&1
tail -3 gpurun_out/r02j_pytest.log
export RG_PERF_NO_ALLSTANCE=1
for lib in default librg_mb9; do
  if [ $lib = default ]; then unset RG_CUDA_LIB; else export RG_CUDA_LIB=$PWD/ab/$lib.so; fi
  timeout 300 python tools/perf_mpc.py 4096 65536 > gpurun_out/r02j_perf_$lib.log 2>&1
  RG_PERF_H=20 timeout 300 python tools/perf_mpc.py 65536 >> gpurun_out/r02j_perf_$lib.log 2>&1
  RG_PERF_H=5 timeout 300 python tools/perf_mpc.py 65536 >> gpurun_out/r02j_perf_$lib.log 2>&1
  RG_PERF_GAIT=bound timeout 300 python tools/perf_mpc.py 65536 >> gpurun_out/r02j_perf_$lib.log 2>&1
  echo $lib; cat gpurun_out/r02j_perf_$lib.log
done
unset RG_CUDA_LIB
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:mpc_solve -s 3 -c 1 python tools/perf_mpc.py 4096 2>&1 | grep -E "dram__|gpu__time|mpc_solve" > gpurun_out/r02j_dram.log
cat gpurun_out/r02j_dram.log
