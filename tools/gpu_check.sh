#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== coupled + sharded GPU tests"; timeout 900 python -m pytest tests/test_gpu_coupled.py tests/test_sharded_group.py -m gpu -q -x 2>&1 | tail -5
