#!/bin/bash
# Round-2 GPU call 6 (1 GPU): the split walk -- parity, then the latency configurations.
mkdir -p gpurun_out
( python -m pytest tests/test_gpu_regular.py tests/test_edit.py tests/test_gpu_fullsize.py tests/test_gpu_cpp_mirror.py -m gpu -q -x ) > gpurun_out/r02_c6_tests.log 2>&1; echo "tests exit $?"; tail -15 gpurun_out/r02_c6_tests.log
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python tools/bench_aux.py > gpurun_out/r02_c6_bench_aux.jsonl 2> gpurun_out/r02_c6_bench_aux.err; echo "aux exit $?"; cut -c1-330 gpurun_out/r02_c6_bench_aux.jsonl; tail -3 gpurun_out/r02_c6_bench_aux.err
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r02_c6_bench.json 2> gpurun_out/r02_c6_bench.err; echo "bench exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/r02_c6_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'])
print(json.dumps(d['configs']['edit_latency']))
print(d['configs']['planet']['ms_per_step'], d['configs']['lod_seam']['ms_both'], d['configs']['page_pass']['whole_pass_ms'])
"
timeout 300 python tools/repro_race.py --edge 64 --chunks 1184 --iters 30 2>&1 | tail -2
timeout 300 python tools/repro_race.py --edge 32 --chunks 140 --iters 100 --full-every 5 2>&1 | tail -2
timeout 300 python tools/repro_race.py --edge 64 --chunks 60 --iters 100 --full-every 5 2>&1 | tail -2
