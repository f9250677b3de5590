#!/bin/bash
# ncu passes (1 GPU): launch list of the bench step, full capture of the dominant SpMV kernel and of the masked SpGEMM kernel.
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --quick > gpurun_out/quick.log 2>&1     # builds the /tmp graph cache
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_spmv.csv \
    python bench.py --steps 3 --warmup 3 --quick > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_run -s 6 -c 2 -f -o gpurun_out/prof_spmv \
    python bench.py --steps 3 --warmup 3 --quick > gpurun_out/ncu_spmv.log 2>&1; echo "spmv capture rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_spgemm.csv \
    python tools/prof_spgemm.py ${SPGEMM_SCALE:-20} 1 masked > gpurun_out/ncu_launches2.log 2>&1; echo "launch list 2 rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:masked_hash_kernel -c 2 -f -o gpurun_out/prof_mhash \
    python tools/prof_spgemm.py ${SPGEMM_SCALE:-20} 1 masked > gpurun_out/ncu_mhash.log 2>&1; echo "mhash capture rc=$?"
# launch lists of the two other configured workloads: BFS (configs[2]) and SSSP sweeps (configs[4] shape at scale 22)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bfs.csv \
    python tools/bfs_bench.py 22 > gpurun_out/ncu_bfs.log 2>&1; echo "bfs launch list rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_sssp.csv \
    python tools/sssp_bench.py 22 --no-cpu > gpurun_out/ncu_sssp.log 2>&1; echo "sssp launch list rc=$?"
# keep the reports small: raw metric tables and per-instruction stall tables as CSV, drop the .ncu-rep
for n in prof_spmv prof_mhash; do
  ncu -i gpurun_out/$n.ncu-rep --page raw --csv > gpurun_out/${n}_raw.csv 2>/dev/null
  ncu -i gpurun_out/$n.ncu-rep --page source --csv --print-source sass 2>/dev/null | cut -c1-400 > gpurun_out/${n}_sass.csv
  rm -f gpurun_out/$n.ncu-rep
done
ls -la gpurun_out | head -30
