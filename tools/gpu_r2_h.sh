#!/bin/bash
# round-2 pass H (1 GPU): where does a masked mxm call spend its device time (phase trace), and the 3-deep e2e pipeline
mkdir -p gpurun_out
echo "== masked spgemm s20, phase trace"; B200GRB_SPGEMM_TRACE=1 timeout 600 python tools/prof_spgemm.py 20 4 masked > gpurun_out/h_spgemm.log 2>&1; echo "rc=$?"; grep -v "^$" gpurun_out/h_spgemm.log | tail -20
echo "== bench (no extras)"; timeout 900 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/h_bench.log 2> gpurun_out/h_bench.err; echo "rc=$?"; python - <<'PY'
import json
d = json.loads(open('gpurun_out/h_bench.log').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'parity_full_size')}); print('e2e', d.get('e2e'))
PY
tail -5 gpurun_out/h_bench.err
