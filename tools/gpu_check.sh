#!/bin/bash
set -u
mkdir -p gpurun_out
for L in 32 16 8 auto; do
  echo "== lanes per warp: $L"
  if [ "$L" = "auto" ]; then unset CMTJB_LANES; else export CMTJB_LANES=$L; fi
  timeout 300 python tools/vsd_run.py 256 64 100000 2>&1 | tail -1
  timeout 300 python tools/bench_configs.py s2 1 2>&1 | grep config | cut -c1-170
done
unset CMTJB_LANES
echo "== parity quick (auto lanes)"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_procedures.py tests/test_gpu_balanced.py -m gpu -q -x 2>&1 | tail -3
