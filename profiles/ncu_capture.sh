#!/bin/bash
# Run on the GPU box under gpurun: launch list + one full capture per hot kernel, exported as CSV
# (the .ncu-rep files with imported source exceed gpurun's 64 MiB return limit, so only CSV pages come back).
CMD="python bench.py --steps 2 --warmup 1 --n-base 50000 --no-cpu-baseline --no-configs"
OUT=gpurun_out
$CMD > $OUT/plain.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
for K in k_var_ozaki k_predict k_sobol_static; do
  SKIP=3; [ $K = k_sobol_static ] && SKIP=1
  ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -f -o /tmp/prof_$K $CMD > $OUT/ncu_$K.log 2>&1
  ncu -i /tmp/prof_$K.ncu-rep --page raw --csv > $OUT/prof_${K}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$K.ncu-rep --page details --csv > $OUT/prof_${K}_details.csv 2>/dev/null
  ncu -i /tmp/prof_$K.ncu-rep --page source --csv > $OUT/prof_${K}_source.csv 2>/dev/null
  ls -la /tmp/prof_$K.ncu-rep >> $OUT/ncu_sizes.txt
done
