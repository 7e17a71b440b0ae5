#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== stochastic tests"; timeout 900 python -m pytest tests/test_stochastic_gpu.py -m gpu -q -x 2>&1 | tail -6
echo "== S3 / S5 rates with the new noise branch"; timeout 300 python tools/bench_configs.py s3 s5s 1 2>&1 | grep config | cut -c1-160
