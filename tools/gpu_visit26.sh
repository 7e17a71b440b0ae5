#!/bin/bash
set -u
for v in base exp; do
  cp tools/_libs/lib_$v.so cmtj_b200/libcmtj_b200.so
  echo "== $v"; timeout 300 python tools/bench_configs.py s3 s5s 1 2>&1 | grep config | cut -c1-200
done
