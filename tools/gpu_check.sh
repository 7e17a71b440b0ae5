#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest procedures"; timeout 900 python -m pytest tests/test_gpu_procedures.py -m gpu -q -x 2>&1 | tail -3
echo "== pimm"; timeout 600 python tools/pimm_run.py 100000 200000 6 2>&1 | tail -4
echo "== pimm (trace)"; CMTJB_TRACE_PROC=1 timeout 600 python tools/pimm_run.py 100000 200000 4 2>&1 | grep -v "chunk" | tail -7
