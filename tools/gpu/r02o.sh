#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
RG_TRACE_N=4096 timeout 600 python tools/trace_mpc.py 0 1 2 3 5 > gpurun_out/r02o_trace_h10.log 2>&1
tail -60 gpurun_out/r02o_trace_h10.log
