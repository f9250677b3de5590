#!/bin/bash
# round-2 pass R (1 GPU): why the pipelined e2e result differs from the serial one (bench e2e.mismatch), with and without T formed in
# w's own buffers; first run of the streamed unmasked SpGEMM at scale 20 (bench extra spgemm_unmasked_streamed)
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'parity_full_size')})
print('e2e', {k: d['e2e'].get(k) for k in ('value', 'ms_per_step', 'pipelined_equals_serial', 'matches_device_result', 'mismatch')}, 'serial', d['e2e'].get('serial', {}).get('value'))
for k in ('spgemm', 'spgemm_unmasked', 'spgemm_unmasked_streamed', 'bfs', 'sssp'):
    v = d.get(k)
    if v: print(k, {kk: vv for kk, vv in v.items() if kk in ('value', 'ms', 'ms_total', 'ms_per_sweep', 'panels', 'products', 'nnz_out', 'gproducts_per_s', 'sum_of_values', 'sum_of_values_closed_form', 'parity_full_size', 'error')}, v.get('roofline', {}).get('frac'), 'cpu', v.get('cpu_baseline', {}).get('value'))
PY
}
echo "== T not in place"; B200GRB_MXV_INPLACE=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --cpu-seconds 1 > gpurun_out/r_bench_noinplace.log 2> gpurun_out/r_bench_noinplace.err; echo "rc=$?"; show gpurun_out/r_bench_noinplace.log; tail -3 gpurun_out/r_bench_noinplace.err
echo "== default"; timeout 600 python bench.py --steps 20 --warmup 5 --cpu-seconds 2 > gpurun_out/r_bench.log 2> gpurun_out/r_bench.err; echo "rc=$?"; show gpurun_out/r_bench.log; tail -3 gpurun_out/r_bench.err
