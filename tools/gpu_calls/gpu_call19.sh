#!/bin/bash
# Round-2 GPU call 19 (1 GPU): hvx_weld_meshes parity + timing; compute-sanitizer on the weld path.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_weld.py tests/test_gpu_cpp_mirror.py tests/test_gpu_transition.py -m gpu -x -q 2>&1 | tail -12
timeout 600 python tools/bench_aux.py > gpurun_out/r02_c19_aux.jsonl 2> gpurun_out/r02_c19_aux.err; grep -E "weld|batch_4096" gpurun_out/r02_c19_aux.jsonl | cut -c1-500; tail -2 gpurun_out/r02_c19_aux.err
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_weld.py -m gpu -x -q -k "errors or transition" 2>&1 | tail -6
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_weld.py -m gpu -x -q -k "errors" 2>&1 | tail -6
