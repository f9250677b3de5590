#!/bin/bash
# One GPU-box pass: smoke, the whole GPU suite, the bench line and the reference arm.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu --maxfail=${MAXFAIL:-12} -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps ${STEPS:-20} --warmup 5 ${BENCH_ARGS} > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'parity_full_size', 'max_rel_err_vs_fp64', 'gpu_launches')})
print('roofline', {k: d['roofline'].get(k) for k in ('frac', 'step_frac', 'kernel_ms', 'traffic')})
print('e2e', {k: d['e2e'].get(k) for k in ('value', 'ms_per_step', 'pipelined_equals_serial')}, 'serial', d['e2e'].get('serial', {}).get('value'))
print('cpu', d.get('cpu_baseline', {}).get('value'), d.get('cpu_baseline', {}).get('cores'))
for k in ('spgemm', 'spgemm_unmasked', 'bfs', 'sssp'):
    v = d.get(k, {})
    print(k, {kk: vv for kk, vv in v.items() if kk in ('value', 'ms', 'ms_total', 'ms_per_sweep', 'parity_full_size', 'error')}, v.get('roofline', {}).get('frac'), 'cpu', v.get('cpu_baseline', {}).get('value'))
PY
tail -5 gpurun_out/bench.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_ref.log').read().strip().splitlines()[-1]); print(d['value'], d['cpu_baseline']['cores'], d['cpu_baseline']['host'])"
echo "== spgemm item size"; for blk in 7 8; do B200GRB_STREAM_BLK=$blk B200GRB_SPGEMM_TRACE=1 timeout 300 python tools/prof_spgemm.py 20 3 masked_S 2>&1 | grep phases | tail -1; done
