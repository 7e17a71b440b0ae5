#!/bin/bash
set -u
mkdir -p gpurun_out
for ms in 100 1000 100000; do
echo "== sampler period $ms ms"
BENCH_SAMPLER_MS=$ms timeout 600 python bench.py --no-secondary --no-cpu-baseline --steps 6 --warmup 3 > gpurun_out/bench_trace.json 2> gpurun_out/bench_trace.err; echo rc=$?
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_trace.json') if l.startswith('{')][-1])
print('e2e', [round(x) for x in d['e2e']['ms_each']], d['e2e']['clocks']['samples'])
print('pimm', [round(x) for x in d['e2e_pimm']['ms_each']], d['e2e_pimm']['clocks']['samples'])
PY
done
