#!/bin/bash
# round-2 pass E (2 GPUs): the library's exchange (all-gather / all-reduce over peer memory) checked, then the bench at N = 2
mkdir -p gpurun_out
echo "== dist check"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/dist_check.py 18 > gpurun_out/e_dist.log 2>&1; echo "rc=$?"; grep -v "^W\|^\[W\|warn" gpurun_out/e_dist.log | tail -30
echo "== bench N=2"; python bench.py --steps 2 --warmup 1 --quick > /dev/null 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/e_bench2.log 2> gpurun_out/e_bench2.err; echo "rc=$?"; tail -c 5000 gpurun_out/e_bench2.log; tail -20 gpurun_out/e_bench2.err | grep -v "^W\|warn"
