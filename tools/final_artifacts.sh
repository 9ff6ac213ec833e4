#!/bin/bash
# Round-end measurement pass on the GPU box (run through gpurun from the repo root):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/final_artifacts.sh r02'
# Writes everything under gpurun_out/<tag>_*; tools/collect_profiles.py turns it into profiles/.
set -u
TAG=${1:-r02}
OUT=gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p $OUT
python -m pytest tests -m gpu -q --durations=10 > $OUT/${TAG}_pytest_gpu.log 2>&1; tail -1 $OUT/${TAG}_pytest_gpu.log
python bench.py > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; cut -c1-200 $OUT/${TAG}_bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference_n1.json 2> $OUT/${TAG}_bench_reference_n1.err; cut -c1-160 $OUT/${TAG}_bench_reference_n1.json
# launch list of the bench command (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 > $OUT/${TAG}_bench_under_ncu.log 2>&1
# full capture of the dominant kernel inside the bench command (4096 envs), at 65536 envs, and the h = 20 kernel
ncu --set full --clock-control none --import-source on -k regex:mpc_solve -s 3 -c 1 -f -o $OUT/${TAG}_prof_bench_mpc python bench.py --steps 2 --warmup 1 > /dev/null 2>&1
RG_PERF_NO_ALLSTANCE=1 ncu --set full --clock-control none --import-source on -k regex:mpc_solve -s 3 -c 1 -f -o $OUT/${TAG}_prof_mpc_65536 python tools/perf_mpc.py 65536 > /dev/null 2>&1
RG_PERF_NO_ALLSTANCE=1 RG_PERF_H=20 ncu --set full --clock-control none --import-source on -k regex:mpc_solve -s 3 -c 1 -f -o $OUT/${TAG}_prof_mpc_h20 python tools/perf_mpc.py 16384 > /dev/null 2>&1
python tools/timeline_mpc.py 4096 > $OUT/${TAG}_timeline_4096.log 2>&1
tools/microbench/chol_bench > $OUT/${TAG}_chol_bench.log 2>&1
python tools/config1_substitute.py --steps 1000 --numpy-steps 60 --out $OUT/${TAG}_config1_substitute.json > $OUT/${TAG}_config1.log 2>&1
compute-sanitizer --tool racecheck python tools/sanitize_run.py > $OUT/${TAG}_racecheck.log 2>&1; tail -1 $OUT/${TAG}_racecheck.log
compute-sanitizer --tool memcheck python tools/sanitize_run.py > $OUT/${TAG}_memcheck.log 2>&1; tail -1 $OUT/${TAG}_memcheck.log
for tool in synccheck initcheck; do compute-sanitizer --tool $tool python tools/sanitize_run.py > $OUT/${TAG}_$tool.log 2>&1; tail -1 $OUT/${TAG}_$tool.log; done
ls -la $OUT | grep ${TAG}_
