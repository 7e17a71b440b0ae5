#!/bin/bash
# ncu_summary.sh NAME KERNEL_REGEX SKIP CMD...: one full ncu capture, exported as raw CSV (small), report deleted
name=$1; pat=$2; skip=$3; shift 3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c 1 -o gpurun_out/prof_$name -f "$@" > gpurun_out/ncu_$name.log 2>&1
echo "ncu $name rc=$?"
ncu -i gpurun_out/prof_$name.ncu-rep --page raw --csv > gpurun_out/prof_$name.raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$name.ncu-rep --page source --csv > gpurun_out/prof_$name.source.csv 2>/dev/null
gzip -f gpurun_out/prof_$name.source.csv
rm -f gpurun_out/prof_$name.ncu-rep
