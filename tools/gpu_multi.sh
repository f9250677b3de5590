#!/bin/bash
# multi-GPU pass (run under gpurun --gpus 8): the library's exchange checked at 8 ranks, then the SpMV scaling line at N = 8, 4, 2
# (headline only: --no-extras keeps the box time down; the full line with the distributed extras is what the driver runs)
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --quick > /dev/null 2>&1     # graph cache
echo "== dist check x8"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 tools/dist_check.py 18 > gpurun_out/multi_dist8.log 2>&1; echo "rc=$?"; grep -E "over 8 GPUs|DIST_CHECK|Error|error" gpurun_out/multi_dist8.log | tail -6
for n in ${NS:-8 4 2}; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2964$n bench.py --gpus $n --steps 50 --warmup 5 ${EXTRA:---no-extras} > gpurun_out/multi_bench$n.log 2> gpurun_out/multi_bench$n.err; echo "N=$n rc=$?"
  python - $n <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f'gpurun_out/multi_bench{n}.log').read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ('value', 'ms_per_step', 'parity_full_size', 'max_rel_err_vs_fp64')}, d.get('exchange', {}).get('ms_per_step'), d['roofline'].get('step_ms_local_spmv'), 'e2e', d.get('e2e', {}).get('value'))
except Exception as e:
    print('no line', e)
PY
done
