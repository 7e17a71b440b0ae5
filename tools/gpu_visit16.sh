#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest procedures"; timeout 900 python -m pytest tests/test_gpu_procedures.py -m gpu -q -x > gpurun_out/pytest_gpu16.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu16.log
echo "== pimm overlapped (trace)"; CMTJB_TRACE_PROC=1 timeout 600 python tools/pimm_run.py 100000 200000 4 > gpurun_out/pimm_phases.log 2>&1; tail -26 gpurun_out/pimm_phases.log
echo "== pimm overlapped"; timeout 600 python tools/pimm_run.py 100000 200000 7 2>&1 | tail -4
echo "== pimm overlapped, 5 GB chunks"; CMTJB_PIMM_CHUNK_MB=5120 timeout 600 python tools/pimm_run.py 100000 200000 6 2>&1 | tail -3
