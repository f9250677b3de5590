#!/bin/bash
# round-2 pass D (1 GPU): pipelined / unpipelined sweep, then the full bench line (new bench.py) + reference arm
mkdir -p gpurun_out
echo "== sweep"; timeout 600 python tools/probe_spmv.py 22 sweep > gpurun_out/d_probe.log 2>&1; echo "rc=$?"; cat gpurun_out/d_probe.log | tail -30
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/d_bench.log 2> gpurun_out/d_bench.err; echo "rc=$?"; tail -c 6000 gpurun_out/d_bench.log; tail -15 gpurun_out/d_bench.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/d_ref.log 2> gpurun_out/d_ref.err; echo "rc=$?"; tail -c 1500 gpurun_out/d_ref.log; tail -5 gpurun_out/d_ref.err
