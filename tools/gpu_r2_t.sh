#!/bin/bash
# round-2 pass T (1 GPU): the state with the TMA stage guard and the expand-sort-compress numeric kernels: whole GPU suite, bench line,
# unmasked SpGEMM with the ESC kernels against the hash kernels they replace
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu --maxfail=12 -p no:cacheprovider > gpurun_out/t_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/t_pytest_gpu.log | cut -c1-400
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/t_bench.log 2> gpurun_out/t_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open('gpurun_out/t_bench.log').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'parity_full_size', 'max_rel_err_vs_fp64', 'gpu_launches')})
print('roofline', {k: d['roofline'].get(k) for k in ('frac', 'step_frac', 'kernel_ms', 'traffic')})
print('e2e', {k: d['e2e'].get(k) for k in ('value', 'ms_per_step', 'pipelined_equals_serial', 'matches_device_result', 'mismatch')}, 'serial', d['e2e'].get('serial', {}).get('value'))
print('cpu', d.get('cpu_baseline', {}).get('value'), d.get('cpu_baseline', {}).get('cores'))
for k in ('spgemm', 'spgemm_unmasked', 'spgemm_unmasked_streamed', 'bfs', 'sssp'):
    v = d.get(k, {})
    print(k, {kk: vv for kk, vv in v.items() if kk in ('value', 'ms', 'ms_total', 'ms_per_sweep', 'products', 'nnz_out', 'parity_full_size', 'error')}, v.get('roofline', {}).get('frac'), 'cpu', v.get('cpu_baseline', {}).get('value'))
PY
tail -3 gpurun_out/t_bench.err
for sc in 17 14; do for esc in 1 0; do echo "== unmasked A.A scale $sc, B200GRB_SPGEMM_ESC=$esc"; B200GRB_SPGEMM_ESC=$esc timeout 200 python tools/prof_spgemm.py $sc 6 unmasked 2>&1 | tail -3; done; done
