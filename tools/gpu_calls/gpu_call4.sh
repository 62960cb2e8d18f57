#!/bin/bash
# Round-2 GPU call 4 (1 GPU): bench with the new e2e / configs, seam tests, ncu of every other kernel, launch list.
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_c4_bench_n1.json 2> gpurun_out/r02_c4_bench_n1.err; echo "bench n1 exit $?"; tail -c 7000 gpurun_out/r02_c4_bench_n1.json; tail -5 gpurun_out/r02_c4_bench_n1.err
( python -m pytest tests/test_lod_seams.py tests/test_gpu_transition.py tests/test_gpu_gather.py tests/test_gpu_publish.py tests/test_meshlets.py tests/test_brick_extract.py -m gpu -q ) > gpurun_out/r02_c4_tests.log 2>&1; echo "tests exit $?"; tail -6 gpurun_out/r02_c4_tests.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fill|heights|transition|gather|meshlet|publish|visibility|brick|edit|records|pack|commit" \
    -o gpurun_out/r02_aux_kernels -f python tools/profile_kernels.py > gpurun_out/r02_c4_ncu_aux.log 2>&1; echo "ncu aux exit $?"; tail -3 gpurun_out/r02_c4_ncu_aux.log
ls -la gpurun_out/r02_aux_kernels.ncu-rep
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_c4_bench_ref.json 2> gpurun_out/r02_c4_bench_ref.err; echo "ref exit $?"; tail -c 1200 gpurun_out/r02_c4_bench_ref.json
