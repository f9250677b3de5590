#!/bin/bash
# ncu passes of round 2 (1 GPU): launch list of the bench step, full captures of the dominant SpMV kernel and of the streaming masked
# SpGEMM kernel, launch lists of the masked SpGEMM / BFS / SSSP workloads.  Everything lands in gpurun_out/r02_*; tools/make_profile_json.py
# turns the launch list + raw pages into profiles/spmv_traffic.json and profiles/spgemm_traffic.json (read by bench.py).
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --quick > gpurun_out/quick.log 2>&1     # builds the /tmp graph cache
# steady-state launches of the timed loop: skip the plan / warm-up launches (the first ~60), list 4 steps
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 90 -c 12 --csv --log-file gpurun_out/r02_launches_spmv_bench.csv \
    python bench.py --steps 10 --warmup 5 --quick > gpurun_out/r02_ncu_launches.log 2>&1; echo "spmv launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_run_hot2 -s 8 -c 1 -f -o gpurun_out/r02_prof_spmv \
    python bench.py --steps 3 --warmup 3 --quick > gpurun_out/r02_ncu_spmv.log 2>&1; echo "spmv capture rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_launches_masked_spgemm.csv \
    python tools/prof_spgemm.py 20 2 masked_S > gpurun_out/r02_ncu_launches2.log 2>&1; echo "spgemm launch list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:masked_stream_kernel -c 3 -f -o gpurun_out/r02_prof_mstream \
    python tools/prof_spgemm.py 20 1 masked_S > gpurun_out/r02_ncu_mstream.log 2>&1; echo "mstream capture rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bfs_s22.csv \
    python tools/bfs_bench.py 22 > gpurun_out/r02_ncu_bfs.log 2>&1; echo "bfs launch list rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_sssp_s22.csv \
    python tools/sssp_bench.py 22 --no-cpu > gpurun_out/r02_ncu_sssp.log 2>&1; echo "sssp launch list rc=$?"
# keep the reports small: raw metric tables and per-instruction stall tables as CSV, drop the .ncu-rep
for n in r02_prof_spmv r02_prof_mstream; do
  ncu -i gpurun_out/$n.ncu-rep --page raw --csv > gpurun_out/${n}_raw.csv 2>/dev/null
  ncu -i gpurun_out/$n.ncu-rep --page source --csv --print-source sass 2>/dev/null | cut -c1-500 > gpurun_out/${n}_sass.csv
  rm -f gpurun_out/$n.ncu-rep
done
ls -la gpurun_out | grep r02_
