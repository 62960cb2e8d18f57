#!/bin/bash
# Round-2 GPU call 29 (2 GPUs): the final build under torchrun (bench at N = 2, both arms) and the NCCL gather test.
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_c29_bench_n2.err; echo "bench n2 exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['n_gpus'], d['e2e']['ms_per_step'], d['e2e'].get('h2d_frac')); print(json.dumps(d['configs']['planet'])[:300])
"; tail -2 gpurun_out/r02_c29_bench_n2.err | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | cut -c1-200
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -2
