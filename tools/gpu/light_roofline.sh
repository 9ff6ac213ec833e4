#!/bin/bash
# ncu launch list of the light kernels for profiles/<tag>_light_kernels.md:  tools/gpu/light_roofline.sh <tag>
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; tag=${1:-r02}
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active
python tools/perf_light.py > gpurun_out/${tag}_light.log 2>&1; cat gpurun_out/${tag}_light.log
ncu --metrics $M --clock-control none -k regex:"step_|state_from_sim|hybrid" --csv --log-file gpurun_out/${tag}_light_ncu.csv python tools/perf_light.py > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:"step_" --csv --log-file gpurun_out/${tag}_light_ncu_4096.csv python tools/perf_light.py 4096 > /dev/null 2>&1
wc -l gpurun_out/${tag}_light_ncu.csv gpurun_out/${tag}_light_ncu_4096.csv
