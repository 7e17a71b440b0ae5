#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu (all)"; S=$(date +%s); timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? wall $(( $(date +%s) - S )) s"; tail -6 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "== probe"; timeout 300 python tools/fp64_probe.py > gpurun_out/fp64_probe.log 2>&1; tail -20 gpurun_out/fp64_probe.log
echo "== bench"; S=$(date +%s); timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$? wall $(( $(date +%s) - S )) s"; cut -c1-600 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
cat > /tmp/probe_one.py <<'PY'
import ctypes as C, sys, os
sys.path.insert(0, os.getcwd())
from cmtj_b200 import _abi
lib = _abi.load_library()
f = lib.cmtjb_fp64_pipe_probe
f.restype = C.c_int
f.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double)]
a, b = C.c_double(), C.c_double()
print(f(0, int(sys.argv[1]), int(sys.argv[2]), 20000, C.byref(a), C.byref(b)), a.value)
PY
bash tools/ncu_summary.sh probe3 fp64_probe_kernel 1 python /tmp/probe_one.py 3 8
bash tools/ncu_summary.sh probe1 fp64_probe_kernel 1 python /tmp/probe_one.py 1 8
