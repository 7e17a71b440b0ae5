#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest procedures"; timeout 900 python -m pytest tests/test_gpu_procedures.py -m gpu -q -x > gpurun_out/pytest_gpu15.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu15.log
echo "== pimm overlapped (trace)"; CMTJB_TRACE_PROC=1 timeout 600 python tools/pimm_run.py 100000 200000 5 > gpurun_out/pimm_phases.log 2>&1; tail -64 gpurun_out/pimm_phases.log
echo "== pimm overlapped"; timeout 600 python tools/pimm_run.py 100000 200000 7 2>&1 | tail -4
echo "== pimm one stream"; CMTJB_PIMM_NO_OVERLAP=1 timeout 600 python tools/pimm_run.py 100000 200000 6 2>&1 | tail -3
echo "== pimm launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/pimm_launches.csv python tools/pimm_run.py 100000 200000 1 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/pimm_launches.csv')) if len(r)>5 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    k=r[4][:60]; agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=float(r[-1])/1e6
for k,(n,ms) in agg.items(): print(f"{n:4d} x {k:60s} {ms:9.3f} ms")
PY
bash tools/ncu_summary.sh s5g group_grid 0 python tools/bench_configs.py s5g 1
