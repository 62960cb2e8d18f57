#!/bin/bash
# Round-2 GPU call 27 (1 GPU): more stress shapes (edit-like sparse dirty sets, mixed surface / empty batches), then the full
# parity suite, smoke and the bench with the final kernel.
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" 2>/dev/null
t() { local lim=$1; shift; local out; out=$(timeout $lim python tools/repro_race.py "$@" 2>&1 | tail -1); echo "$* :: ${out:-NO OUTPUT (time limit $lim s)}"; }
{
for lib in jitter1 jitter2; do
  t 90 --lib build/variants/libhvx_$lib.so --edge 64 --chunks 256 --iters 10 --full-every 5 --sparse-dirty
  t 90 --lib build/variants/libhvx_$lib.so --edge 32 --chunks 256 --iters 10 --full-every 5 --sparse-dirty
  t 90 --lib build/variants/libhvx_$lib.so --edge 32 --chunks 1500 --iters 10 --full-every 5 --mixed
  t 90 --lib build/variants/libhvx_$lib.so --edge 64 --chunks 600 --iters 10 --full-every 5 --mixed
  t 90 --lib build/variants/libhvx_$lib.so --edge 64 --chunks 5 --iters 20 --full-every 5
  t 90 --lib build/variants/libhvx_$lib.so --edge 32 --chunks 1 --iters 20 --full-every 5
done
} 2>&1 | tee gpurun_out/r02_stress_more.txt
( python -m pytest tests -m gpu -q ) > gpurun_out/r02_c27_tests.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/r02_c27_tests.log
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_c27_bench.json 2> gpurun_out/r02_c27_bench.err; echo "bench exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/r02_c27_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'], d.get('fill_kernel'), d['clocks'])
for k,v in d['configs'].items(): print(k, json.dumps(v)[:420])
"
