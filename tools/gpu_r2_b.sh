#!/bin/bash
# round-2 pass B: why does the TMA-staged hot kernel alternate between ~225 and ~315 us with the table size?
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --quick > gpurun_out/quick.log 2>&1     # builds the /tmp graph cache
for kb in 128 160; do
  B200GRB_SPMV_HOT=$kb timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 12 --csv --log-file gpurun_out/b_launches_$kb.csv \
      python bench.py --steps 3 --warmup 3 --quick > gpurun_out/b_ncu_l_$kb.log 2>&1; echo "launch list $kb rc=$?"
  B200GRB_SPMV_HOT=$kb timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_run_hot2 -s 3 -c 1 -f -o gpurun_out/b_prof_$kb \
      python bench.py --steps 3 --warmup 3 --quick > gpurun_out/b_ncu_f_$kb.log 2>&1; echo "full capture $kb rc=$?"
  ncu -i gpurun_out/b_prof_$kb.ncu-rep --page raw --csv > gpurun_out/b_prof_${kb}_raw.csv 2>/dev/null
  ncu -i gpurun_out/b_prof_$kb.ncu-rep --page source --csv --print-source sass 2>/dev/null | cut -c1-600 > gpurun_out/b_prof_${kb}_sass.csv
  rm -f gpurun_out/b_prof_$kb.ncu-rep
done
grep -h "spmv\|Duration" gpurun_out/b_launches_128.csv | tail -8
grep -h "spmv\|Duration" gpurun_out/b_launches_160.csv | tail -8
