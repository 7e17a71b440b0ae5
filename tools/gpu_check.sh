#!/bin/bash
# final-state check: whole GPU suite, smoke, the bench line
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -1
echo "== quick parity"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_coupled.py -m gpu -q -x 2>&1 | tail -3
