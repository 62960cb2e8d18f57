#!/bin/bash
# Round-2 GPU call 2: which of the two protocol changes removes the round-1 fault; full parity suite with the new paths.
mkdir -p gpurun_out
: > gpurun_out/r02_c2_repro.log
for lib in build/variants/libhvx_legacy.so build/variants/libhvx_legacy_lane.so build/variants/libhvx_legacy_first.so \
           build/variants/libhvx_legacy_check.so build/variants/libhvx_legacy_edge0.so build/variants/libhvx_legacy_jitter2.so \
           "" build/variants/libhvx_jitter2.so; do
  for edge in 64 32; do
    n=1184; [ $edge = 32 ] && n=2368
    echo "=== lib=$lib edge=$edge" >> gpurun_out/r02_c2_repro.log
    timeout 200 python tools/repro_race.py ${lib:+--lib $lib} --edge $edge --chunks $n --iters 40 2>&1 | grep -v "^  File\|^    \|Traceback\|During handling" | tail -20 >> gpurun_out/r02_c2_repro.log
    echo "exit $?" >> gpurun_out/r02_c2_repro.log
  done
done
grep -E "===|RESULT|invariant|^iteration" gpurun_out/r02_c2_repro.log | cut -c1-260
# the legacy build under memcheck on a smaller batch (does the tool see the access that faults?)
( timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python tools/repro_race.py --lib build/variants/libhvx_legacy.so --edge 64 --chunks 296 --iters 3 ) > gpurun_out/r02_c2_legacy_memcheck.log 2>&1
grep -E "Invalid|at 0x|by thread|RESULT|ERROR SUMMARY" gpurun_out/r02_c2_legacy_memcheck.log | head -20
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02_c2_gpu_tests.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/r02_c2_gpu_tests.log
tail -15 gpurun_out/r02_c2_gpu_tests.log
python __graft_entry__.py smoke > gpurun_out/r02_c2_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r02_c2_smoke.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r02_c2_bench.json 2> gpurun_out/r02_c2_bench.err; tail -c 900 gpurun_out/r02_c2_bench.json
