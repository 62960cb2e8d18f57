#!/bin/bash
# Round-2 GPU call 17 (1 GPU): idle-wait sleep A/B (0 / 100 / 400 ns) on the headline, all-surface and planet workloads;
# transition path without memset; tests.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_regular.py tests/test_gpu_transition.py tests/test_lod_seams.py -m gpu -x -q 2>&1 | tail -3
for v in default idle0 idle400; do
  if [ $v = default ]; then unset HVX_LIBRARY; else export HVX_LIBRARY=$PWD/build/variants/libhvx_$v.so; fi
  timeout 300 python bench.py --no-configs --steps 20 --warmup 5 > gpurun_out/r02_c17_bench_$v.json 2> gpurun_out/r02_c17_bench_$v.err
  timeout 300 python bench.py --workload surface --no-configs --steps 20 --warmup 5 > gpurun_out/r02_c17_surface_$v.json 2> gpurun_out/r02_c17_surface_$v.err
  timeout 300 python tools/probe_planet_shard.py 1 8 > gpurun_out/r02_c17_shard_$v.jsonl 2> gpurun_out/r02_c17_shard_$v.err
  python - <<PY
import json
for f in ('bench','surface'):
    d=json.loads(open('gpurun_out/r02_c17_%s_$v.json'%f).read().strip().splitlines()[-1])
    print('$v', f, d['ms_per_step'], d['roofline']['frac'])
for line in open('gpurun_out/r02_c17_shard_$v.jsonl'):
    d=json.loads(line); print('$v', d['case'], 'regular', d['whole_chunks_only_regular_ms']['ms_median'], 'step', d.get('whole_chunks_only_step_ms',{}).get('ms_median'), 'transition', d.get('transition_ms',{}).get('ms_median'))
PY
done
unset HVX_LIBRARY
timeout 300 python tools/bench_aux.py --latency-only 2>/dev/null | cut -c1-330
