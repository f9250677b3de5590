#!/bin/bash
# round-2 pass C: the whole GPU suite (incl. the unmodified reference's own tests and matrix assign), then a table-size sweep
# of the pipelined hot kernel
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -q -m gpu --maxfail=15 -p no:cacheprovider > gpurun_out/c_pytest.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/c_pytest.log
echo "== sweep"; timeout 600 python tools/probe_spmv.py 22 sweep > gpurun_out/c_probe.log 2>&1; echo "rc=$?"; cat gpurun_out/c_probe.log | tail -30
