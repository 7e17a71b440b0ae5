#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest procedures"; timeout 900 python -m pytest tests/test_gpu_procedures.py -m gpu -q -x 2>&1 | tail -3
echo "== spec timing, split"; timeout 600 python tools/spec_run.py 20000 200000 80e9 moments moments:200 moments:50 2>&1 | tail -3
echo "== spec timing, single-block form"; CMTJB_SPECTRUM_SPLIT=0 timeout 600 python tools/spec_run.py 20000 200000 80e9 moments moments:200 moments:50 2>&1 | tail -3
echo "== pimm"; timeout 600 python tools/pimm_run.py 100000 200000 6 2>&1 | tail -3
echo "== pimm, single-block form"; CMTJB_SPECTRUM_SPLIT=0 timeout 600 python tools/pimm_run.py 100000 200000 5 2>&1 | tail -2
