#!/bin/bash
# multi-GPU sanity + scaling (run under gpurun --gpus N)
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --quick > /dev/null 2>&1     # graph cache
for n in $@; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 50 --warmup 5 --no-spgemm > gpurun_out/bench_n$n.log 2> gpurun_out/bench_n$n.err
  echo "N=$n rc=$?"; tail -1 gpurun_out/bench_n$n.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k: d[k] for k in ('value','ms_per_step','n_gpus')}, d['roofline']['kernel_ms'], d['e2e']['value'])" 2>&1 | tail -1
done
