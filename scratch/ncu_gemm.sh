#!/bin/bash
CMD="python scratch/lml_profile.py 4096 10 32 1"
OUT=gpurun_out
$CMD > $OUT/gemm_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:k_gemm_nt -s 60 -c 3 -f -o /tmp/prof_gemm $CMD > $OUT/ncu_gemm.log 2>&1
ncu -i /tmp/prof_gemm.ncu-rep --page raw --csv > $OUT/prof_gemm_raw.csv 2>/dev/null
ncu -i /tmp/prof_gemm.ncu-rep --page details --csv > $OUT/prof_gemm_details.csv 2>/dev/null
ncu -i /tmp/prof_gemm.ncu-rep --page source --csv > $OUT/prof_gemm_source.csv 2>/dev/null
tail -3 $OUT/ncu_gemm.log
