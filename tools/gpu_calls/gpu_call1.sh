#!/bin/bash
# Round-2 GPU call 1: parity suite on the reworked kernel selection, stress reproducer on every build, sanitizers, quick bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_c1_gpu.txt 2>&1
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_c1_gpu_tests.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/r02_c1_gpu_tests.log
python __graft_entry__.py smoke > gpurun_out/r02_c1_smoke.log 2>&1; echo "smoke exit $?"
: > gpurun_out/r02_c1_repro.log
for lib in "" build/variants/libhvx_legacy.so build/variants/libhvx_check.so build/variants/libhvx_edge0.so \
           build/variants/libhvx_jitter1.so build/variants/libhvx_jitter2.so build/variants/libhvx_edge0_jitter2.so \
           build/variants/libhvx_legacy_jitter2.so; do
  for edge in 64 32; do
    n=1184; [ $edge = 32 ] && n=2368
    timeout 240 python tools/repro_race.py ${lib:+--lib $lib} --edge $edge --chunks $n --iters 60 2>&1 | tail -25 >> gpurun_out/r02_c1_repro.log
    echo "exit $? lib=$lib edge=$edge" >> gpurun_out/r02_c1_repro.log
  done
done
grep -E "RESULT|exit" gpurun_out/r02_c1_repro.log
( timeout 400 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitize_small.py ) > gpurun_out/r02_c1_racecheck.log 2>&1; echo "racecheck exit $?"
( timeout 300 compute-sanitizer --tool synccheck python tools/sanitize_small.py ) > gpurun_out/r02_c1_synccheck.log 2>&1; echo "synccheck exit $?"
( timeout 300 compute-sanitizer --tool memcheck python tools/sanitize_small.py ) > gpurun_out/r02_c1_memcheck.log 2>&1; echo "memcheck exit $?"
tail -3 gpurun_out/r02_c1_racecheck.log gpurun_out/r02_c1_synccheck.log gpurun_out/r02_c1_memcheck.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r02_c1_bench.json 2> gpurun_out/r02_c1_bench.err; tail -c 1500 gpurun_out/r02_c1_bench.json
python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload surface > gpurun_out/r02_c1_bench_surface.json 2>/dev/null; tail -c 600 gpurun_out/r02_c1_bench_surface.json
tail -5 gpurun_out/r02_c1_gpu_tests.log
