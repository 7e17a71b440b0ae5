#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== bench"; S=$(date +%s); timeout 1500 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$? wall $(( $(date +%s) - S )) s"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench.json') if l.startswith('{')][-1])
print('value',d['value'],'e2e',d['e2e']['value'],'pimm',d['e2e_pimm']['value'])
for k,v in d['secondary'].items(): print(k[:90], '%.3e'%v.get('value',0), v.get('seconds'), v.get('error'))
PY
