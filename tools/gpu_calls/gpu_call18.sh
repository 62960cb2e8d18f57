#!/bin/bash
# Round-2 GPU call 18 (1 GPU): transition path fix (aligned table copy) + idle-sleep A/B on the planet set.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_transition.py tests/test_lod_seams.py tests/test_gpu_regular.py -m gpu -x -q 2>&1 | tail -8
for v in default idle0; do
  if [ $v = default ]; then unset HVX_LIBRARY; else export HVX_LIBRARY=$PWD/build/variants/libhvx_$v.so; fi
  timeout 300 python tools/probe_planet_shard.py 1 8 > gpurun_out/r02_c18_shard_$v.jsonl 2> gpurun_out/r02_c18_shard_$v.err
  python - <<PY
import json
for line in open('gpurun_out/r02_c18_shard_$v.jsonl'):
    d=json.loads(line); print('$v', d['case'], 'regular', d['whole_chunks_only_regular_ms']['ms_median'], 'step', d.get('whole_chunks_only_step_ms',{}).get('ms_median'), 'transition', d.get('transition_ms',{}).get('ms_median'))
PY
  tail -2 gpurun_out/r02_c18_shard_$v.err | cut -c1-200
done
