#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest"; timeout 1800 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_procedures.py -m gpu -q > gpurun_out/pytest_gpu8.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_gpu8.log
echo "== pimm"; CMTJB_TRACE_PROC=1 timeout 600 python tools/pimm_run.py 100000 200000 3 > gpurun_out/pimm_phases.log 2>&1; tail -10 gpurun_out/pimm_phases.log
echo "== bench"; S=$(date +%s); timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$? wall $(( $(date +%s) - S )) s"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
bash tools/ncu_summary.sh pruned4 pruned_dft4 0 python tools/pimm_run.py 20000 200000 1
