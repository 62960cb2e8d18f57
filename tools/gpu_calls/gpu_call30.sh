#!/bin/bash
# Round-2 GPU call 30 (1 GPU): soak -- the product build and the jittered builds, many iterations, every shape.
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" 2>/dev/null
t() { local lim=$1; shift; local out; out=$(timeout $lim python tools/repro_race.py "$@" 2>&1 | tail -1); echo "$* :: ${out:-NO OUTPUT (time limit $lim s)}"; }
{
t 200 --edge 32 --chunks 140 --iters 400 --full-every 20
t 200 --edge 32 --chunks 900 --iters 300 --full-every 20
t 200 --edge 64 --chunks 60 --iters 300 --full-every 20
t 200 --edge 64 --chunks 1184 --iters 60 --full-every 10
t 200 --edge 64 --chunks 256 --iters 400 --full-every 20 --sparse-dirty
t 200 --edge 32 --chunks 4000 --iters 60 --full-every 10 --mixed
for lib in jitter2 jitter1; do
  t 200 --lib build/variants/libhvx_$lib.so --edge 32 --chunks 140 --iters 150 --full-every 10
  t 200 --lib build/variants/libhvx_$lib.so --edge 32 --chunks 900 --iters 100 --full-every 10
  t 200 --lib build/variants/libhvx_$lib.so --edge 64 --chunks 60 --iters 150 --full-every 10
  t 200 --lib build/variants/libhvx_$lib.so --edge 64 --chunks 300 --iters 60 --full-every 10 --mixed
  t 200 --lib build/variants/libhvx_$lib.so --edge 64 --chunks 256 --iters 150 --full-every 10 --sparse-dirty
  t 200 --lib build/variants/libhvx_$lib.so --edge 64 --chunks 1184 --iters 15 --full-every 5
done
} 2>&1 | tee gpurun_out/r02_stress_soak.txt | cut -c1-330
