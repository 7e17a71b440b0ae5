#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests/test_gpu_procedures.py -m gpu -q -x > gpurun_out/pytest_gpu9.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu9.log
echo "== pimm"; CMTJB_TRACE_PROC=1 timeout 600 python tools/pimm_run.py 100000 200000 4 > gpurun_out/pimm_phases.log 2>&1; tail -12 gpurun_out/pimm_phases.log
echo "== pimm launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/pimm_launches.csv python tools/pimm_run.py 100000 200000 1 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/pimm_launches.csv')) if len(r)>5 and r[0].isdigit()]
for r in rows: print(r[4][:70], r[-1], r[-2])
PY
bash tools/ncu_summary.sh pruned4 pruned_dft4 0 python tools/pimm_run.py 20000 200000 1
