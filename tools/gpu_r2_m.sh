#!/bin/bash
# round-2 pass M (1 GPU): masked stream kernel variants (block size of long B rows), phase trace
mkdir -p gpurun_out
python tools/prof_spgemm.py 20 1 masked_S > /dev/null 2>&1
for blk in 7; do
  echo "== blk=$blk"; B200GRB_STREAM_BLK=$blk B200GRB_SPGEMM_TRACE=1 timeout 300 python tools/prof_spgemm.py 20 3 masked_S 2>&1 | grep phases | tail -1
done
echo "== pytest mxm"; timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu --maxfail=5 -p no:cacheprovider -k "mxm or triangle" > gpurun_out/m_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/m_pytest.log
