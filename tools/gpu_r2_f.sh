#!/bin/bash
# round-2 pass F (1 GPU): streaming masked SpGEMM kernel: parity, then timing at scale 20
mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests/test_parity_gpu.py tests/test_matrix_ops_gpu.py tests/test_vector_ops_gpu.py -q -m gpu --maxfail=10 -p no:cacheprovider > gpurun_out/f_pytest.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/f_pytest.log
echo "== masked spgemm s20"; timeout 600 python tools/prof_spgemm.py 20 5 masked > gpurun_out/f_spgemm.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/f_spgemm.log
echo "== unmasked spgemm s17"; timeout 600 python tools/prof_spgemm.py 17 3 unmasked > gpurun_out/f_spgemm_u.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/f_spgemm_u.log
