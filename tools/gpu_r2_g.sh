#!/bin/bash
# round-2 pass G (1 GPU): masked SpGEMM stream kernel v2 (hash / dense-map resolve): parity, timing, launch list; then the bench line
mkdir -p gpurun_out
echo "== pytest mxm"; timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_matrix_ops_gpu.py -q -m gpu --maxfail=10 -p no:cacheprovider -k "mxm or triangle or spgemm or goldens or api_forms" > gpurun_out/g_pytest.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/g_pytest.log
echo "== masked spgemm s20"; timeout 600 python tools/prof_spgemm.py 20 4 masked > gpurun_out/g_spgemm.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/g_spgemm.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/g_launches_spgemm.csv python tools/prof_spgemm.py 20 1 masked > gpurun_out/g_ncu.log 2>&1; echo "launch list rc=$?"
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/g_launches_spgemm.csv')))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hi]; kn = h.index('Kernel Name'); mv = h.index('Metric Value'); gs = h.index('Grid Size'); bs = h.index('Block Size')
for r in rows[hi + 1:]:
    if len(r) > mv and ('masked' in r[kn] or 'row_found' in r[kn] or 'flops' in r[kn]):
        print(r[kn][:60], r[gs], r[bs], float(r[mv].replace(',', '')) / 1e3, 'us')
PY
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/g_bench.log 2> gpurun_out/g_bench.err; echo "rc=$?"; python - <<'PY'
import json
d = json.loads(open('gpurun_out/g_bench.log').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'parity_full_size', 'max_rel_err_vs_fp64', 'max_rel_diff_vs_cpu_port')})
print('e2e', d.get('e2e'))
print('cpu', d.get('cpu_baseline'))
for k in ('spgemm', 'spgemm_unmasked', 'bfs', 'sssp'):
    v = d.get(k, {})
    print(k, {kk: vv for kk, vv in v.items() if kk in ('value', 'ms', 'ms_total', 'ms_per_sweep', 'parity_full_size', 'error')})
PY
tail -5 gpurun_out/g_bench.err
