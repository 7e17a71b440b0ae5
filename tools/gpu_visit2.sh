#!/bin/bash
# round-2 GPU visit: parity (one pytest session -> one ledger), the group kernel before/after numbers, block pinning of the
# balanced kernel, the FP64 peak microbenchmark shapes, PIMM phases.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt
echo "== pytest"; CMTJB_PARITY_FRESH=1 timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== s5g"; timeout 300 python tools/bench_configs.py s5g > gpurun_out/s5g.jsonl 2>&1; cut -c1-400 gpurun_out/s5g.jsonl
echo "== s1 pinned vs not"; timeout 300 python tools/s1_run.py 100000 20000 100 3 > gpurun_out/s1_pin.log 2>&1; tail -2 gpurun_out/s1_pin.log
CMTJB_BALANCE_NOPIN=1 timeout 300 python tools/s1_run.py 100000 20000 100 3 > gpurun_out/s1_nopin.log 2>&1; tail -2 gpurun_out/s1_nopin.log
CMTJB_BALANCE_WARPS=5 timeout 300 python tools/s1_run.py 100000 20000 100 3 > gpurun_out/s1_pin5.log 2>&1; tail -1 gpurun_out/s1_pin5.log
echo "== fp64 peak"; python -c "
import ctypes as C
from cmtj_b200 import _abi
lib=_abi.load_library(); f,ms=C.c_double(),C.c_double()
for k in range(3):
    lib.cmtjb_fp64_peak(0, C.c_int64(100000), C.byref(f), C.byref(ms)); print('fp64 peak TFLOP/s', f.value/1e12, 'ms', ms.value)
" 2>&1 | tee gpurun_out/fp64_peak.log
echo "== pimm phases"; CMTJB_TRACE_PROC=1 timeout 600 python tools/pimm_run.py 100000 200000 2 > gpurun_out/pimm_phases.log 2>&1; tail -20 gpurun_out/pimm_phases.log
echo "== ncu group kernel"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:group_grid -s 1 -c 1 -o gpurun_out/prof_s5g -f python tools/bench_configs.py s5g 0.999 > gpurun_out/ncu_s5g.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out
