#!/bin/bash
# One GPU-box pass: smoke, GPU parity tests, short bench.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu --maxfail=${MAXFAIL:-12} -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -60 gpurun_out/pytest_gpu.log
echo "== pytest gpu, run kernel forced for every dense-u specialised case (incl. tiny matrices)"; B200GRB_SPMV_RUN=1 timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "random_mxv or goldens or api_forms or sssp or config1" --maxfail=8 -p no:cacheprovider > gpurun_out/pytest_gpu_run.log 2>&1; echo "pytest(run) rc=$?"; tail -5 gpurun_out/pytest_gpu_run.log
echo "== bench"; timeout 900 python bench.py --steps ${STEPS:-20} --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.log; tail -15 gpurun_out/bench.err
