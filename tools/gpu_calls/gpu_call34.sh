#!/bin/bash
# Round-2 GPU call 34 (1 GPU): cells per tile at edge 32: 30 (default) / 31 / 32 -- parity and timing on the planet set,
# the per-rank shard and the 4096-page terrain batch.
mkdir -p gpurun_out
for v in default tc32 tc31; do
  if [ $v = default ]; then unset HVX_LIBRARY; else export HVX_LIBRARY=$PWD/build/variants/libhvx_$v.so; fi
  timeout 600 python -m pytest tests/test_gpu_regular.py tests/test_lod_seams.py -m gpu -x -q 2>&1 | tail -1
  timeout 300 python tools/bench_aux.py 2>/dev/null | grep -E "batch_4096x32" | cut -c1-160
  timeout 300 python tools/probe_planet_shard.py 1 8 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    d=json.loads(line); print('$v', d['case'], 'regular', d['whole_chunks_only_regular_ms']['ms_median'])
"
done
