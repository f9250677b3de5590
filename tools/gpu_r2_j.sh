#!/bin/bash
# round-2 pass J (1 GPU): masked mxm with persistent scratch (phase trace), new matrix ops tests, full bench line
mkdir -p gpurun_out
echo "== pytest matrix ops"; timeout 900 python -m pytest tests/test_matrix_ops_gpu.py tests/test_reference_gpu.py -q -m gpu --maxfail=10 -p no:cacheprovider > gpurun_out/j_pytest.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/j_pytest.log; tail -25 gpurun_out/reference_suite_on_gpu.txt
echo "== masked spgemm s20, phase trace"; B200GRB_SPGEMM_TRACE=1 timeout 600 python tools/prof_spgemm.py 20 5 masked > gpurun_out/j_spgemm.log 2>&1; echo "rc=$?"; grep -v "^$" gpurun_out/j_spgemm.log | tail -20
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/j_bench.log 2> gpurun_out/j_bench.err; echo "rc=$?"; python - <<'PY'
import json
d = json.loads(open('gpurun_out/j_bench.log').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'parity_full_size')}); print('e2e', d.get('e2e', {}).get('value'))
for k in ('spgemm', 'spgemm_unmasked', 'bfs', 'sssp'):
    v = d.get(k, {})
    print(k, {kk: vv for kk, vv in v.items() if kk in ('value', 'ms', 'ms_total', 'ms_per_sweep', 'parity_full_size', 'error')}, v.get('roofline', {}).get('frac'))
PY
tail -5 gpurun_out/j_bench.err
