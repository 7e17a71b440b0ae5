#!/bin/bash
# r02b round: bench line, its launch list, full ncu of S1 + the other kernel families, streamed PIMM timing
set -u
mkdir -p gpurun_out; rm -rf gpurun_out/summary
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pimm streamed"; timeout 600 python tools/pimm_run.py 100000 200000 6 2>&1 | tail -3
echo "== bench"; S=$(date +%s); timeout 1500 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$? wall $(( $(date +%s) - S )) s"; cut -c1-400 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
echo "== ncu full S1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:llgs_balanced -s 1 -c 1 -o gpurun_out/prof_s1 -f \
    python tools/s1_run.py 100000 20000 100 2 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
for spec in "s3:regex:llgs_:0.1" "s4:regex:llgs_:0.02" "s5s:regex:stack_kernel:0.25" "s5g:regex:group_:0.25"; do
  IFS=: read cfg kind pat scale <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k $kind:$pat -s 1 -c 1 -o gpurun_out/prof_$cfg -f \
      python tools/bench_configs.py $cfg $scale > gpurun_out/ncu_$cfg.log 2>&1; echo "ncu $cfg rc=$?"
done
SUMMARY_OUT=gpurun_out/summary python tools/summarise_profiles.py r02b > gpurun_out/summarise.log 2>&1; echo "summarise rc=$?"; tail -3 gpurun_out/summarise.log
for n in s1 s3 s4 s5s s5g; do
  ncu -i gpurun_out/prof_$n.ncu-rep --page raw --csv > gpurun_out/prof_$n.raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_$n.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/prof_$n.source.csv.gz
  rm -f gpurun_out/prof_$n.ncu-rep
done
du -sh gpurun_out; ls gpurun_out/summary
