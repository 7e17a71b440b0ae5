#!/bin/bash
# final-state check: whole GPU suite, smoke, the bench line
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu (all)"; S=$(date +%s); timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? wall $(( $(date +%s) - S )) s"; tail -4 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -1
echo "== bench"; S=$(date +%s); timeout 1500 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$? wall $(( $(date +%s) - S )) s"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench.json') if l.startswith('{')][-1])
print('value',d['value'],'e2e',d['e2e']['value'],[round(x) for x in d['e2e']['ms_each']],'pimm',d['e2e_pimm']['value'],[round(x) for x in d['e2e_pimm']['ms_each']])
for k,v in d['secondary'].items(): print(k[:70], '%.3e'%v.get('value',0))
PY
