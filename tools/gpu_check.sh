#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== compute-sanitizer memcheck: streamed spectrum + F_PROJ + background kernel"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/sanitizer_proc.log python -m pytest tests/test_gpu_procedures.py -m gpu -q -x -k "stream or streamed" 2>&1 | tail -3; echo "rc=$?"; tail -4 gpurun_out/sanitizer_proc.log
echo "== compute-sanitizer memcheck: sharded arrays (callback + peer memory)"
CMTJB_PEER_TIMEOUT_MS=60000 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/sanitizer_group.log python -m pytest tests/test_sharded_group.py -m gpu -q -x -k "peer and 400 or thermal" 2>&1 | tail -3; echo "rc=$?"; tail -4 gpurun_out/sanitizer_group.log
echo "== compute-sanitizer memcheck: float64 normals"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/sanitizer_noise.log python -m pytest tests/test_stochastic_gpu.py -m gpu -q -x -k "float64_normal or philox" 2>&1 | tail -3; echo "rc=$?"; tail -3 gpurun_out/sanitizer_noise.log
