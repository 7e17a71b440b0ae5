#!/bin/bash
# One GPU-box session: parity tests, the bench line, the ncu launch list of the same command and a
# full ncu capture of the dominant kernel.  Outputs land in gpurun_out/ (scratch); summaries are
# copied into profiles/ by tools/summarise_profiles.py afterwards.
set -u
rm -rf gpurun_out/summary; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log

echo "== bench"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== ncu launch list (same bench command, reduced steps)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
echo "== ncu full (S1 kernel, 100000 instances x 20000 steps, log every 100)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:llgs_balanced -s 1 -c 1 -o gpurun_out/prof_s1 -f \
    python tools/s1_run.py 100000 20000 100 2 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
if [ "${CONFIGS:-0}" = "1" ]; then
  echo "== per-configuration throughput (S2..S5)"; timeout 900 python tools/bench_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; echo "configs rc=$?"; cut -c1-200 gpurun_out/configs.jsonl
fi
if [ "${PROFILE_CONFIGS:-0}" = "1" ]; then
  # one full capture per other kernel family (reduced sizes: ncu replays each kernel ~40 times)
  for spec in "s3:regex:llgs_:0.1" "s4:regex:llgs_:0.02" "s5s:regex:stack_kernel:0.25" "s5g:regex:group_:0.25"; do
    IFS=: read cfg kind pat scale <<< "$spec"
    timeout 600 ncu --set full --clock-control none --import-source on -k $kind:$pat -s 1 -c 1 -o gpurun_out/prof_$cfg -f \
        python tools/bench_configs.py $cfg $scale > gpurun_out/ncu_$cfg.log 2>&1; echo "ncu $cfg rc=$?"
  done
fi
# summarise on the box (the reports together exceed what gpurun copies back) and keep only the S1 report
SUMMARY_OUT=gpurun_out/summary python tools/summarise_profiles.py ${TAG:-r01} > gpurun_out/summarise.log 2>&1; echo "summarise rc=$?"
for f in gpurun_out/prof_s3.ncu-rep gpurun_out/prof_s4.ncu-rep gpurun_out/prof_s5s.ncu-rep gpurun_out/prof_s5g.ncu-rep; do rm -f $f; done
du -sh gpurun_out; ls -la gpurun_out gpurun_out/summary
