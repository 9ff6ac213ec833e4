#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
RG_TRACE_SOLO=1 RG_TRACE_N=4096 timeout 300 python tools/trace_mpc.py 0 3 > gpurun_out/r02d_trace.log 2>&1
cat gpurun_out/r02d_trace.log
