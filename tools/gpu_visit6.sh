#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== tests (procedures, fullsize)"; timeout 900 python -m pytest tests/test_gpu_procedures.py tests/test_gpu_fullsize.py tests/test_gpu_balanced.py -m gpu -q -x 2>&1 | tail -5
echo "== s1 / s4 regression"; timeout 300 python tools/s1_run.py 100000 200000 100 3 2>&1 | tail -1
timeout 300 python tools/bench_configs.py s4 s3 2>&1 | cut -c1-140 | tail -3
echo "== pimm pruned"; CMTJB_TRACE_PROC=1 timeout 600 python tools/pimm_run.py 100000 200000 3 > gpurun_out/pimm_phases.log 2>&1; tail -6 gpurun_out/pimm_phases.log
echo "== pimm cufft"; CMTJB_SPECTRUM=cufft CMTJB_TRACE_PROC=1 timeout 600 python tools/pimm_run.py 100000 50000 2 > gpurun_out/pimm_cufft.log 2>&1; tail -5 gpurun_out/pimm_cufft.log
CMTJB_TRACE_PROC=1 timeout 600 python tools/pimm_run.py 100000 50000 2 2>&1 | tail -5
bash tools/ncu_summary.sh pruned pruned_dft 0 python tools/pimm_run.py 20000 200000 1
bash tools/ncu_summary.sh s5g group_grid 3 python tools/bench_configs.py s5g
du -sh gpurun_out
