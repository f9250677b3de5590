#!/bin/bash
# round-2 pass P (8 GPUs): the full bench line (with the distributed extras) at N = 8, the headline at N = 4
NS=8 EXTRA=" " bash tools/gpu_multi.sh
python - <<'PY'
import json
d = json.loads(open('gpurun_out/multi_bench8.log').read().strip().splitlines()[-1])
for k in ('spgemm', 'bfs', 'sssp', 'exchange'):
    v = d.get(k, {})
    print(k, {kk: vv for kk, vv in v.items() if kk in ('value', 'ms', 'ms_total', 'ms_per_sweep', 'ms_per_step', 'local_spmv_ms', 'parity_full_size', 'error')})
PY
NS=4 bash tools/gpu_multi.sh 2>&1 | grep -v "dist check\|^rc=\|over 8\|DIST_CHECK"
