#!/bin/bash
# Round-2 GPU call 10 (2 GPUs): NCCL gather test + bandwidth, bench at N = 2 (weak headline, strong planet set).
mkdir -p gpurun_out
( python -m pytest tests/test_multi_gpu.py -m gpu -q -rs ) > gpurun_out/r02_c10_nccl_test.log 2>&1; echo "nccl test exit $?"; tail -3 gpurun_out/r02_c10_nccl_test.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/bench_gather.py > gpurun_out/r02_gather_n2.json 2> gpurun_out/r02_c10_gather_n2.err; echo "gather exit $?"; cat gpurun_out/r02_gather_n2.json; tail -3 gpurun_out/r02_c10_gather_n2.err | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_c10_bench_n2.err; echo "bench n2 exit $?"; tail -c 2500 gpurun_out/r02_bench_n2.json; tail -4 gpurun_out/r02_c10_bench_n2.err | cut -c1-300
