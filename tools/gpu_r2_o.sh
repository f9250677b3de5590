#!/bin/bash
# round-2 pass O (2 GPUs): push kernel variant: exchange check + collective timing + bench at N = 2
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --quick > /dev/null 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29651 tools/dist_check.py 18 > gpurun_out/o_dist2.log 2>&1; echo "rc=$?"; grep -E "over 2 GPUs|DIST_CHECK|MISMATCH|Error" gpurun_out/o_dist2.log | tail -6
NS=2 bash tools/gpu_multi.sh 2>&1 | grep -v "dist check\|^rc=\|over 8" | tail -4
echo "== 1 GPU: masked spgemm trace + mxm tests"; B200GRB_SPGEMM_TRACE=1 timeout 300 python tools/prof_spgemm.py 20 3 masked_S 2>&1 | grep phases | tail -2
timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu --maxfail=5 -p no:cacheprovider -k "mxm or triangle" 2>&1 | tail -2
