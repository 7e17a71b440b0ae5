#!/bin/bash
set -u
mkdir -p gpurun_out
N=${N:-2}
for spec in "4096 2000" "16384 400"; do
  set -- $spec
  echo "== sharded group, $1 elements, $N ranks"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/group_sharded_run.py $1 $2 > gpurun_out/group_sharded_$1_n$N.log 2>&1; echo "rc=$?"; grep "cmtjb\|what" gpurun_out/group_sharded_$1_n$N.log | tail -12
done
