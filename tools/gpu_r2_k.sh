#!/bin/bash
# round-2 pass K (1 GPU): full ncu capture of the streaming masked SpGEMM kernel (class S launch) with source-level stalls
mkdir -p gpurun_out
python tools/prof_spgemm.py 20 1 masked > /dev/null 2>&1      # graph cache
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:masked_stream_kernel -c 1 -f -o gpurun_out/k_prof python tools/prof_spgemm.py 20 1 masked > gpurun_out/k_ncu.log 2>&1; echo "capture rc=$?"
ncu -i gpurun_out/k_prof.ncu-rep --page raw --csv > gpurun_out/k_prof_raw.csv 2>/dev/null
ncu -i gpurun_out/k_prof.ncu-rep --page source --csv --print-source sass 2>/dev/null | cut -c1-700 > gpurun_out/k_prof_sass.csv
ncu -i gpurun_out/k_prof.ncu-rep --page source --csv --print-source cuda 2>/dev/null | cut -c1-700 > gpurun_out/k_prof_cuda.csv
rm -f gpurun_out/k_prof.ncu-rep
ls -la gpurun_out/k_prof*; tail -3 gpurun_out/k_ncu.log
echo "== 1/8 block probe"; timeout 600 python tools/probe_block.py 8 > gpurun_out/k_block.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/k_block.log
echo "== pytest new (hyper, extract, assign, ref suite)"; timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_matrix_ops_gpu.py tests/test_reference_gpu.py -q -m gpu --maxfail=10 -p no:cacheprovider -k "hyper or extract or diag or assign or reference" > gpurun_out/k_pytest.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/k_pytest.log; grep -c passed gpurun_out/reference_suite_on_gpu.txt; tail -12 gpurun_out/reference_suite_on_gpu.txt
