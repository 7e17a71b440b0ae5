#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== bench"; timeout 1500 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench.json'))
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')})
    print('e2e', {k:d['e2e'][k] for k in ('value','ms_per_step','pipeline')} if d.get('e2e') else None)
    print('e2e_pimm', {k:d['e2e_pimm'][k] for k in ('value','ms_per_step','ms_each','d2h_bytes_per_step')} if d.get('e2e_pimm') else None)
    print('roofline', {k:d['roofline'][k] for k in ('achieved','peak','frac','peak_source')})
    for k,v in (d.get('secondary') or {}).items(): print(k, {a:b for a,b in v.items() if a not in ('clocks','note')})
    print('parity', d.get('parity_gate'))
    print('cpu', d.get('cpu_baseline'))
except Exception as e:
    print('bench parse failed', e)
PY
tail -5 gpurun_out/bench.err
echo "== s3 64 regs"; timeout 300 python tools/bench_configs.py s3 s5s > gpurun_out/s3_64.jsonl 2>&1; cut -c1-150 gpurun_out/s3_64.jsonl | tail -3
CMTJB_STACK_BLOCK=128 timeout 300 python tools/bench_configs.py s5s > gpurun_out/s5s_128.jsonl 2>&1; cut -c1-150 gpurun_out/s5s_128.jsonl | tail -1
echo "== pimm"; CMTJB_TRACE_PROC=1 timeout 600 python tools/pimm_run.py 100000 200000 2 > gpurun_out/pimm_phases.log 2>&1; tail -6 gpurun_out/pimm_phases.log
bash tools/ncu_summary.sh s5g group_grid 3 python tools/bench_configs.py s5g
du -sh gpurun_out
