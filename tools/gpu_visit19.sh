#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== S4 rolled Heun"; timeout 600 python tools/bench_configs.py s4 0.25 2>&1 | cut -c1-330
echo "== parity (Heun cases)"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_balanced.py -m gpu -q -x -k "Heun or heun or balanced or split" 2>&1 | tail -4
bash tools/ncu_summary.sh s4 llgs_ 1 python tools/bench_configs.py s4 0.02
python tools/ncu_show.py gpurun_out/prof_s4.raw.csv | head -30
