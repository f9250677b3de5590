#!/bin/bash
# round-2 pass I (8 GPUs): exchange check at 8 ranks, then the bench at N = 8 and N = 4
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --quick > /dev/null 2>&1     # graph cache
echo "== dist check x8"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 tools/dist_check.py 18 > gpurun_out/i_dist8.log 2>&1; echo "rc=$?"; grep -E "allgather|allreduce|DIST_CHECK|Error|error" gpurun_out/i_dist8.log | tail -12
for n in 8 4; do
  echo "== bench N=$n"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2963$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/i_bench$n.log 2> gpurun_out/i_bench$n.err; echo "rc=$?"
  python - $n <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f'gpurun_out/i_bench{n}.log').read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ('value', 'ms_per_step', 'parity_full_size', 'max_rel_err_vs_fp64', 'exchange')})
    print('e2e', d.get('e2e', {}).get('value'), 'roofline', {k: d['roofline'].get(k) for k in ('frac', 'kernel_ms', 'step_ms_local_spmv')})
    for k in ('spgemm', 'bfs', 'sssp'):
        v = d.get(k, {})
        print(k, {kk: vv for kk, vv in v.items() if kk in ('value', 'ms', 'ms_total', 'ms_per_sweep', 'parity_full_size', 'error')})
except Exception as e:
    print('no line', e)
PY
  grep -v "^W\|warn\|^\*\*\*\|OMP_NUM" gpurun_out/i_bench$n.err | tail -8
done
