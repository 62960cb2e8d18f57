#!/bin/bash
# Round-2 GPU call 32 (1 GPU): start order of a hinted batch -- heavy chunks spread over 50 / 75 / 90 % of the order or all
# first (plain descending) -- on the headline batch; cost-hint parity tests.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_regular.py -m gpu -x -q -k "hint or order or cost or pipeline" 2>&1 | tail -2
for sp in 255 50 75 90 255 75; do
  timeout 300 python bench.py --no-configs --no-cpu --no-e2e --steps 30 --warmup 5 --spread $sp > gpurun_out/r02_c32_spread_$sp.json 2> gpurun_out/r02_c32_spread_$sp.err
  python -c "
import json
d=json.loads(open('gpurun_out/r02_c32_spread_$sp.json').read().strip().splitlines()[-1]); print('spread $sp', d['ms_per_step'], d['roofline']['frac'])
"
done
