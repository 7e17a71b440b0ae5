#!/bin/bash
# r02c round: full GPU suite, smoke, bench line, its launch list, full ncu of S1 / stack / group / background moments kernel
set -u
mkdir -p gpurun_out; rm -rf gpurun_out/summary
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu (all)"; S=$(date +%s); timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? wall $(( $(date +%s) - S )) s"; tail -6 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "== bench"; S=$(date +%s); timeout 1500 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$? wall $(( $(date +%s) - S )) s"; cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
echo "== ncu full S1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:llgs_balanced -s 1 -c 1 -o gpurun_out/prof_s1 -f \
    python tools/s1_run.py 100000 20000 100 2 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
for spec in "s3:regex:llgs_:0.1:1" "s4:regex:llgs_:0.02:1" "s5s:regex:stack_kernel:1:1" "s5g:regex:group_grid:1:0"; do
  IFS=: read cfg kind pat scale skip <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k $kind:$pat -s $skip -c 1 -o gpurun_out/prof_$cfg -f \
      python tools/bench_configs.py $cfg $scale > gpurun_out/ncu_$cfg.log 2>&1; echo "ncu $cfg rc=$?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cheb_moments_bg -s 2 -c 1 -o gpurun_out/prof_mombg -f \
    python tools/pimm_run.py 100000 200000 1 > gpurun_out/ncu_mombg.log 2>&1; echo "ncu mombg rc=$?"
SUMMARY_OUT=gpurun_out/summary python tools/summarise_profiles.py r02c > gpurun_out/summarise.log 2>&1; echo "summarise rc=$?"; tail -3 gpurun_out/summarise.log
for n in s1 s3 s4 s5s s5g mombg; do
  ncu -i gpurun_out/prof_$n.ncu-rep --page raw --csv > gpurun_out/prof_$n.raw.csv 2>/dev/null
  rm -f gpurun_out/prof_$n.ncu-rep
done
python tools/ncu_show.py gpurun_out/prof_mombg.raw.csv gpurun_out/prof_s5g.raw.csv gpurun_out/prof_s5s.raw.csv 2>/dev/null | head -40
du -sh gpurun_out; ls gpurun_out/summary
