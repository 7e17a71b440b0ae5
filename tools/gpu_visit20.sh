#!/bin/bash
set -u
for v in base cap96 cap80; do
  cp tools/_libs/lib_$v.so cmtj_b200/libcmtj_b200.so
  echo "== stack kernel, $v"; timeout 300 python tools/bench_configs.py s5s 1 2>&1 | cut -c1-260
done
