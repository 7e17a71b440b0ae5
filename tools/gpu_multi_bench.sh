#!/bin/bash
set -u
mkdir -p gpurun_out
N=${N:-8}
echo "== bench $N GPUs"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"; tail -3 gpurun_out/bench_n$N.err; python - <<PY
import json
for line in open('gpurun_out/bench_n$N.json'):
    if line.startswith('{'):
        d=json.loads(line)
        print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],d['e2e']['ms_each'], d['e2e'].get('warmup_steps'))
        print('pipeline',d['e2e'].get('pipeline'))
        print('pimm',d['e2e_pimm']['value'],d['e2e_pimm']['ms_per_step'],d['e2e_pimm']['ms_each'],d['e2e_pimm'].get('warmup_steps'))
PY
echo "== sharded group, 16384 elements, $N ranks"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/group_sharded_run.py 16384 400 > gpurun_out/group_sharded_16384_n$N.log 2>&1; echo "rc=$?"; grep "cmtjb\|what" gpurun_out/group_sharded_16384_n$N.log | tail -20
