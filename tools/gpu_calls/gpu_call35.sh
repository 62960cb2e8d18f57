#!/bin/bash
# Round-2 GPU call 35 (1 GPU): per-step tile width at edge 32 (full-row steps: 32 cells per tile) against the narrow build:
# parity (regular, seams, edit, weld, stress), planet set, per-rank shard, terrain batch, LOD-seam config.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_regular.py tests/test_lod_seams.py tests/test_edit.py tests/test_gpu_weld.py tests/test_gpu_stress.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -2
for v in default narrow; do
  if [ $v = default ]; then unset HVX_LIBRARY; else export HVX_LIBRARY=$PWD/build/variants/libhvx_$v.so; fi
  timeout 300 python tools/bench_aux.py 2>/dev/null | grep -E "batch_4096x32|single_page" | cut -c1-170
  timeout 300 python tools/probe_planet_shard.py 1 8 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    d=json.loads(line); print('$v', d['case'], 'regular', d['whole_chunks_only_regular_ms']['ms_median'], 'step', d['whole_chunks_only_step_ms']['ms_median'])
"
done
unset HVX_LIBRARY
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_c35_bench.json 2> gpurun_out/r02_c35_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_c35_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'])
for k,v in d['configs'].items(): print(k, json.dumps(v)[:330])
"
