#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== s1 k=auto / 4 / 5, pinned"; for k in auto 4 5; do
  if [ $k = auto ]; then timeout 300 python tools/s1_run.py 100000 200000 100 3 > gpurun_out/s1_k$k.log 2>&1; else CMTJB_BALANCE_WARPS=$k timeout 300 python tools/s1_run.py 100000 200000 100 3 > gpurun_out/s1_k$k.log 2>&1; fi; tail -1 gpurun_out/s1_k$k.log; done
CMTJB_BALANCE_NOPIN=1 CMTJB_BALANCE_WARPS=5 timeout 300 python tools/s1_run.py 100000 200000 100 3 > gpurun_out/s1_k5_nopin.log 2>&1; tail -1 gpurun_out/s1_k5_nopin.log
echo "== s5g 768 / 512"; timeout 300 python tools/bench_configs.py s5g > gpurun_out/s5g_768.jsonl 2>&1; cut -c1-330 gpurun_out/s5g_768.jsonl | head -2 | tail -1
CMTJB_GROUP_THREADS=512 timeout 300 python tools/bench_configs.py s5g > gpurun_out/s5g_512.jsonl 2>&1; cut -c1-330 gpurun_out/s5g_512.jsonl | head -2 | tail -1
echo "== s3 s4 s5s"; timeout 600 python tools/bench_configs.py s3 s4 s5s > gpurun_out/configs.jsonl 2>&1; cut -c1-260 gpurun_out/configs.jsonl
echo "== ncu s1"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:llgs_balanced -s 1 -c 1 -o gpurun_out/prof_s1 -f python tools/s1_run.py 100000 20000 100 2 > gpurun_out/ncu_s1.log 2>&1; echo "ncu rc=$?"
echo "== ncu s5g"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:group_grid -s 1 -c 1 -o gpurun_out/prof_s5g -f python tools/bench_configs.py s5g > gpurun_out/ncu_s5g.log 2>&1; echo "ncu rc=$?"
echo "== ncu s3"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:llgs_ -s 1 -c 1 -o gpurun_out/prof_s3 -f python tools/bench_configs.py s3 0.1 > gpurun_out/ncu_s3.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out
