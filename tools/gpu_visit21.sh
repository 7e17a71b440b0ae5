#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== sharded group tests (one GPU)"; CMTJB_PEER_TIMEOUT_MS=3000 timeout 240 python -m pytest tests/test_sharded_group.py -m gpu -q -x -k peer 2>&1 | tail -15
