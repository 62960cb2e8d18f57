#!/bin/bash
# Round-2 GPU call 14 (8 GPUs): the driver's scaling run at N = 8, once, to see it finish (weak headline, e2e against the
# host fabric, strong planet set), and the optional NCCL gather at eight ranks.
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_c14_bench_n8.err; echo "bench n8 exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_n8.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['n_gpus']); print(d['e2e']); print(d['configs']['planet'])
"; tail -3 gpurun_out/r02_c14_bench_n8.err | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 tools/bench_gather.py > gpurun_out/r02_gather_n8.json 2> gpurun_out/r02_c14_gather_n8.err; echo "gather exit $?"; cat gpurun_out/r02_gather_n8.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/r02_c14_ref_n8.json 2>/dev/null; echo "ref n8 exit $?"; cut -c1-200 gpurun_out/r02_c14_ref_n8.json
