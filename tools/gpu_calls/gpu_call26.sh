#!/bin/bash
# Round-2 GPU call 26 (1 GPU): the stress builds after the fix (a warp waits for the next slab only while it holds a tile);
# parity tests and the headline number with the fixed kernel.
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" 2>/dev/null
t() { local lim=$1; shift; local out; out=$(timeout $lim python tools/repro_race.py "$@" 2>&1 | tail -1); echo "$* :: ${out:-NO OUTPUT (time limit $lim s)}"; }
{
for lib in jitter1 jitter2 check; do
  t 90 --lib build/variants/libhvx_$lib.so --edge 32 --chunks 140 --iters 10 --full-every 5
  t 90 --lib build/variants/libhvx_$lib.so --edge 32 --chunks 140 --iters 10 --full-every 5 --no-split
  t 90 --lib build/variants/libhvx_$lib.so --edge 32 --chunks 900 --iters 10 --full-every 5
  t 90 --lib build/variants/libhvx_$lib.so --edge 64 --chunks 60 --iters 10 --full-every 5
  t 90 --lib build/variants/libhvx_$lib.so --edge 64 --chunks 300 --iters 10 --full-every 5
done
} 2>&1 | tee gpurun_out/r02_stress_final.txt
timeout 600 python -m pytest tests/test_gpu_regular.py tests/test_edit.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --no-configs --steps 20 --warmup 5 > gpurun_out/r02_c26_bench.json 2> gpurun_out/r02_c26_bench.err
timeout 300 python bench.py --workload surface --no-configs --no-cpu --no-e2e --steps 20 --warmup 5 > gpurun_out/r02_c26_surface.json 2> gpurun_out/r02_c26_surface.err
python -c "
import json
for f in ('r02_c26_bench','r02_c26_surface'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d['roofline']['frac'])
"
timeout 200 python tools/timeline.py 2>&1 | tee gpurun_out/r02_timeline.jsonl | cut -c1-700
