#!/bin/bash
# round-2 pass A: parity of the SpMV paths after the hot-table/TMA rework, then the probe and a quick bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest parity"; timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -x -p no:cacheprovider > gpurun_out/a_pytest.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/a_pytest.log
echo "== probe"; timeout 600 python tools/probe_spmv.py 22 > gpurun_out/a_probe.log 2>&1; echo "rc=$?"; cat gpurun_out/a_probe.log | tail -20
echo "== quick bench"; timeout 600 python bench.py --quick --steps 50 --warmup 5 > gpurun_out/a_bench.log 2> gpurun_out/a_bench.err; echo "rc=$?"; tail -2 gpurun_out/a_bench.log; tail -5 gpurun_out/a_bench.err
