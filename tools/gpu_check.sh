#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== wide-kernel tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "many_layer or custom" 2>&1 | tail -5
