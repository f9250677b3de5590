#!/bin/bash
# round-2 pass Q (1 GPU, re-entry after the container was replaced): the whole GPU suite, the bench line,
# the launch list of the bench command, and the local SpMV of one rank's block of an 8- / 4- / 2-way partition (kernel list under ncu)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu --maxfail=${MAXFAIL:-12} -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'parity_full_size', 'max_rel_err_vs_fp64', 'gpu_launches')})
print('roofline', {k: d['roofline'].get(k) for k in ('frac', 'step_frac', 'kernel_ms', 'traffic')})
print('e2e', {k: d['e2e'].get(k) for k in ('value', 'ms_per_step', 'pipelined_equals_serial', 'through')}, 'serial', d['e2e'].get('serial', {}).get('value'))
print('cpu', d.get('cpu_baseline', {}).get('value'), d.get('cpu_baseline', {}).get('cores'))
for k in ('spgemm', 'spgemm_unmasked', 'bfs', 'sssp'):
    v = d.get(k, {})
    print(k, {kk: vv for kk, vv in v.items() if kk in ('value', 'ms', 'ms_total', 'ms_per_sweep', 'parity_full_size', 'error')}, v.get('roofline', {}).get('frac'), 'cpu', v.get('cpu_baseline', {}).get('value'))
PY
tail -5 gpurun_out/bench.err
echo "== launch list of the bench command"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/q_launches_spmv_bench.csv python bench.py --steps 5 --warmup 3 --quick > gpurun_out/q_ncu_bench.log 2>&1; echo "rc=$?"
for w in 8 4; do
  echo "== block of a $w-way partition"; timeout 300 python tools/probe_block.py $w 2>&1 | tail -6
done
echo "== launch list: block of an 8-way partition"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:"spmv|comm" -c 60 --csv --log-file gpurun_out/q_launches_block8.csv python tools/probe_block.py 8 > gpurun_out/q_ncu_block8.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv, collections
for f in ('gpurun_out/q_launches_spmv_bench.csv', 'gpurun_out/q_launches_block8.csv'):
    try:
        rows = [r for r in csv.reader(open(f)) if len(r) > 10]
        hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
        agg = collections.OrderedDict()
        for r in rows[1:]:
            k = r[ki][:70]; agg.setdefault(k, []).append(float(r[vi].replace(',', '')))
        print(f)
        for k, v in agg.items():
            print(f"  {k:70s} n={len(v):3d} last={v[-1]/1e3:9.1f} us  min={min(v)/1e3:9.1f} us")
    except Exception as e:
        print(f, 'unreadable', e)
PY
