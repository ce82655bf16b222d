#!/bin/bash
export BATMAN_B200_SIGMA=int8
CMD="python bench.py --steps 1 --warmup 1 --n-base 20000 --no-cpu-baseline"
OUT=gpurun_out
$CMD > $OUT/plain_oz.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:k_var_ozaki -s 2 -c 1 -f -o /tmp/prof_oz $CMD > $OUT/ncu_oz.log 2>&1
ncu -i /tmp/prof_oz.ncu-rep --page raw --csv > $OUT/prof_k_var_ozaki_raw.csv 2>/dev/null
ncu -i /tmp/prof_oz.ncu-rep --page details --csv > $OUT/prof_k_var_ozaki_details.csv 2>/dev/null
