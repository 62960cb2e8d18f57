#!/bin/bash
# Round-2 GPU call 22 (1 GPU): stress builds of the final protocol (jittered emission / scheduler, invariant checks) on split
# and unsplit walks; initcheck on the small all-kernel workload.
mkdir -p gpurun_out
{
for lib in jitter2 jitter1 check; do
  for cfg in "32 140" "64 60" "64 300" "32 900"; do
    set -- $cfg
    timeout 600 python tools/repro_race.py --lib build/variants/libhvx_$lib.so --edge $1 --chunks $2 --iters 25 --full-every 5 2>&1 | tail -1
  done
done
} | tee gpurun_out/r02_stress_final.txt
( echo "== compute-sanitizer --tool initcheck tools/sanitize_small.py"; timeout 900 compute-sanitizer --tool initcheck python tools/sanitize_small.py 2>&1 | tail -25 ) > gpurun_out/r02_initcheck.txt 2>&1; tail -25 gpurun_out/r02_initcheck.txt | cut -c1-250
