#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest procedures + sharded group"; timeout 900 python -m pytest tests/test_gpu_procedures.py tests/test_sharded_group.py -m gpu -q -x > gpurun_out/pytest_gpu13.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu13.log
echo "== pimm streamed"; CMTJB_TRACE_PROC=1 timeout 600 python tools/pimm_run.py 100000 200000 5 > gpurun_out/pimm_phases.log 2>&1; tail -32 gpurun_out/pimm_phases.log
echo "== pimm streamed, 2 GB chunks"; CMTJB_PIMM_CHUNK_MB=2048 timeout 600 python tools/pimm_run.py 100000 200000 4 2>&1 | tail -2
echo "== pimm streamed, 16 GB chunks"; CMTJB_PIMM_CHUNK_MB=16384 timeout 600 python tools/pimm_run.py 100000 200000 4 2>&1 | tail -2
echo "== pimm launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/pimm_launches.csv python tools/pimm_run.py 100000 200000 1 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/pimm_launches.csv')) if len(r)>5 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    k=r[4][:60]; agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=float(r[-1])/1e6
for k,(n,ms) in agg.items(): print(f"{n:4d} x {k:60s} {ms:9.3f} ms")
PY
echo "== full gpu suite"; S=$(date +%s); timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? wall $(( $(date +%s) - S )) s"; tail -8 gpurun_out/pytest_gpu.log
