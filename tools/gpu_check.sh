#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest procedures"; timeout 900 python -m pytest tests/test_gpu_procedures.py -m gpu -q -x 2>&1 | tail -3
echo "== pimm"; timeout 600 python tools/pimm_run.py 100000 200000 6 2>&1 | tail -4
echo "== pimm, overlap experiment"; CMTJB_PIMM_OVERLAP=1 timeout 600 python tools/pimm_run.py 100000 200000 5 2>&1 | tail -3
