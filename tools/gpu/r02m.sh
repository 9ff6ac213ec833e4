#!/bin/bash
# two-rank run of the bench (NCCL), smoke(), and the suites touched last
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
echo "bench n2 exit $?"; cut -c1-400 gpurun_out/r02_bench_n2.json; tail -5 gpurun_out/r02_bench_n2.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_env_shim.py tests/test_gpu_mpc.py -m gpu -q 2>&1 | tail -3
