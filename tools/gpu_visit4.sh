#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== s1 sched x pin at k=4, 200000 steps"
for sched in ticket claim; do for pin in 1 0; do
  if [ $pin = 0 ]; then export CMTJB_BALANCE_NOPIN=1; else unset CMTJB_BALANCE_NOPIN; fi
  CMTJB_BALANCE_SCHED=$sched timeout 300 python tools/s1_run.py 100000 200000 100 4 > gpurun_out/s1_${sched}_pin$pin.log 2>&1; echo "$sched pin=$pin: $(tail -2 gpurun_out/s1_${sched}_pin$pin.log | cut -c60-140 | tr '\n' '|')"
done; done
unset CMTJB_BALANCE_NOPIN
echo "== s3 carried vs stateless"; timeout 300 python tools/bench_configs.py s3 > gpurun_out/s3_carry.jsonl 2>&1; cut -c1-150 gpurun_out/s3_carry.jsonl | tail -2
CMTJB_TUNE=1 timeout 300 python tools/bench_configs.py s3 > gpurun_out/s3_stateless.jsonl 2>&1; cut -c1-150 gpurun_out/s3_stateless.jsonl | tail -2
echo "== s5g 768 / 512"; timeout 300 python tools/bench_configs.py s5g > gpurun_out/s5g_768.jsonl 2>&1; head -2 gpurun_out/s5g_768.jsonl | tail -1 | tr ',' '\n' | grep -E "us_per_step|overhead|pair_inter"
CMTJB_GROUP_THREADS=512 timeout 300 python tools/bench_configs.py s5g > gpurun_out/s5g_512.jsonl 2>&1; head -2 gpurun_out/s5g_512.jsonl | tail -1 | tr ',' '\n' | grep -E "us_per_step|overhead|pair_inter"
echo "== pimm"; CMTJB_TRACE_PROC=1 timeout 600 python tools/pimm_run.py 100000 200000 3 > gpurun_out/pimm_phases.log 2>&1; tail -12 gpurun_out/pimm_phases.log
echo "== procedure tests"; timeout 600 python -m pytest tests/test_gpu_procedures.py -m gpu -q -x 2>&1 | tail -3
echo "== ncu"
CMTJB_BALANCE_SCHED=ticket bash tools/ncu_summary.sh s1_ticket llgs_balanced 1 python tools/s1_run.py 100000 20000 100 2
CMTJB_BALANCE_SCHED=claim bash tools/ncu_summary.sh s1_claim llgs_balanced 1 python tools/s1_run.py 100000 20000 100 2
bash tools/ncu_summary.sh s5g group_grid 1 python tools/bench_configs.py s5g
bash tools/ncu_summary.sh s3 llgs_ 1 python tools/bench_configs.py s3 0.1
CMTJB_TUNE=1 bash tools/ncu_summary.sh s3_stateless llgs_ 1 python tools/bench_configs.py s3 0.1
du -sh gpurun_out; ls -la gpurun_out
