#!/bin/bash
# round-2 pass L (1 GPU): masked stream kernel with the lean inner loop: parity, phase trace, bench line
mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests -q -m gpu --maxfail=10 -p no:cacheprovider > gpurun_out/l_pytest.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/l_pytest.log
echo "== masked spgemm s20, phase trace"; B200GRB_SPGEMM_TRACE=1 timeout 600 python tools/prof_spgemm.py 20 3 masked > gpurun_out/l_spgemm.log 2>&1; echo "rc=$?"; grep "phases" gpurun_out/l_spgemm.log | tail -4
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/l_bench.log 2> gpurun_out/l_bench.err; echo "rc=$?"; python - <<'PY'
import json
d = json.loads(open('gpurun_out/l_bench.log').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'parity_full_size')}); print('e2e', d.get('e2e', {}).get('value'))
for k in ('spgemm', 'spgemm_unmasked', 'bfs', 'sssp'):
    v = d.get(k, {})
    print(k, {kk: vv for kk, vv in v.items() if kk in ('value', 'ms', 'ms_total', 'ms_per_sweep', 'parity_full_size', 'error')}, v.get('roofline', {}).get('frac'))
PY
tail -5 gpurun_out/l_bench.err
