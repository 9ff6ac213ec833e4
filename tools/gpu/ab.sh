#!/bin/bash
# A/B perf of library builds: tools/gpu/ab.sh <tag> <lib1> <lib2> ...   (lib = "default" or a name under ab/ without .so)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
tag=$1; shift
if [ -z "$AB_ALLSTANCE" ]; then export RG_PERF_NO_ALLSTANCE=1; fi
for lib in "$@"; do
  if [ $lib = default ]; then unset RG_CUDA_LIB; else export RG_CUDA_LIB=$PWD/ab/$lib.so; fi
  out=gpurun_out/${tag}_perf_$lib.log
  RG_PERF_H=${AB_H:-10} RG_PERF_GAIT=${AB_GAIT:-trot} timeout 300 python tools/perf_mpc.py ${AB_SIZES:-4096 65536} > $out 2>&1
  echo "== $lib"; cat $out
done
