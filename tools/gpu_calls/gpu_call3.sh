#!/bin/bash
# Round-2 GPU call 3 (2 GPUs): NCCL gather test + bandwidth, 2-GPU bench; then 1-GPU bench with the new e2e / configs.
mkdir -p gpurun_out
( python -m pytest tests/test_multi_gpu.py -m gpu -q -rs ) > gpurun_out/r02_c3_nccl_test.log 2>&1; echo "nccl test exit $?"; tail -4 gpurun_out/r02_c3_nccl_test.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/bench_gather.py > gpurun_out/r02_c3_gather_n2.json 2> gpurun_out/r02_c3_gather_n2.err; cat gpurun_out/r02_c3_gather_n2.json; tail -3 gpurun_out/r02_c3_gather_n2.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_c3_bench_n1.json 2> gpurun_out/r02_c3_bench_n1.err; echo "bench n1 exit $?"; tail -c 6000 gpurun_out/r02_c3_bench_n1.json; tail -5 gpurun_out/r02_c3_bench_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_c3_bench_n2.json 2> gpurun_out/r02_c3_bench_n2.err; echo "bench n2 exit $?"; tail -c 3500 gpurun_out/r02_c3_bench_n2.json; tail -5 gpurun_out/r02_c3_bench_n2.err
( python -m pytest tests/test_gpu_regular.py tests/test_edit.py -m gpu -q -x ) > gpurun_out/r02_c3_tests.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/r02_c3_tests.log
