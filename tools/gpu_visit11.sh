#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests/test_gpu_procedures.py -m gpu -q -x > gpurun_out/pytest_gpu11.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu11.log
echo "== spec timing"; timeout 600 python tools/spec_run.py 20000 200000 80e9 moments moments:50 moments:200 moments:125 grouped > gpurun_out/spec_run.log 2>&1; cat gpurun_out/spec_run.log
echo "== launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/spec_launches.csv python tools/spec_run.py 20000 200000 80e9 moments > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/spec_launches.csv')) if len(r)>5 and r[0].isdigit()]
for r in rows: print(r[4][:70], r[-1], r[-2])
PY
echo "== pimm"; CMTJB_TRACE_PROC=1 timeout 600 python tools/pimm_run.py 100000 200000 4 > gpurun_out/pimm_phases.log 2>&1; tail -14 gpurun_out/pimm_phases.log
bash tools/ncu_summary.sh chebmom cheb_moments 0 python tools/spec_run.py 20000 200000 80e9 moments
bash tools/ncu_summary.sh chebcomb cheb_combine 0 python tools/spec_run.py 20000 200000 80e9 moments
