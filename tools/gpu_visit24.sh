#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== bench (driver command)"; S=$(date +%s); timeout 1500 python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$? wall $(( $(date +%s) - S )) s"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench.json') if l.startswith('{')][-1])
print('value',d['value'],'e2e',d['e2e']['value'],d['e2e']['ms_each'],'pimm',d['e2e_pimm']['value'],d['e2e_pimm']['ms_each'])
print(d['roofline']['frac'], d['roofline']['pipe_probe_issue_slot_frac'], d['roofline']['frac_of_rhs_shape_ceiling'])
for k,v in d['secondary'].items(): print(k[:70], '%.3e'%v.get('value',0))
PY
echo "== reference arm"; timeout 900 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-400
