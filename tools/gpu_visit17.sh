#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pimm overlapped"; CMTJB_TRACE_PROC=1 timeout 600 python tools/pimm_run.py 100000 200000 10 > gpurun_out/pimm_phases.log 2>&1; grep -v "chunk  *[2-9]:\|chunk  1[0-3]:" gpurun_out/pimm_phases.log | tail -90
