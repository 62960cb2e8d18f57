#!/bin/bash
# Round-2 GPU call 15 (1 GPU): parallel look-back + 16 parts + split last wave: tests, latency cases, per-rank planet shard A/B.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_regular.py tests/test_edit.py tests/test_gpu_cpp_mirror.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/bench_aux.py --latency-only > gpurun_out/r02_c15_aux.jsonl 2> gpurun_out/r02_c15_aux.err
grep -E "single_page|single_chunk" gpurun_out/r02_c15_aux.jsonl | cut -c1-600
timeout 600 python tools/probe_planet_shard.py > gpurun_out/r02_planet_shard.jsonl 2> gpurun_out/r02_c15_shard.err; cat gpurun_out/r02_planet_shard.jsonl; tail -3 gpurun_out/r02_c15_shard.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_c15_bench.json 2> gpurun_out/r02_c15_bench.err; echo "bench exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/r02_c15_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'])
for k,v in d['configs'].items(): print(k, json.dumps(v)[:700])
"
