#!/bin/bash
# Round-2 GPU call 23 (1 GPU): transition kernel, vertices in flight per thread (1 / 2 / 4): parity + timing.
mkdir -p gpurun_out
for v in default unroll2 unroll4; do
  if [ $v = default ]; then unset HVX_LIBRARY; else export HVX_LIBRARY=$PWD/build/variants/libhvxt_$v.so; fi
  timeout 600 python -m pytest tests/test_gpu_transition.py tests/test_lod_seams.py -m gpu -x -q 2>&1 | tail -1
  timeout 600 python tools/bench_aux.py 2>/dev/null | grep -E "lod_seam_1024x64\^3_mask0x3f_transition" | cut -c1-220
  timeout 600 python tools/probe_planet_shard.py 1 8 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    d=json.loads(line); print('$v', d['case'], 'transition', d['transition_ms']['ms_median'], 'step', d['whole_chunks_only_step_ms']['ms_median'])
"
done
