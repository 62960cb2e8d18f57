#!/bin/bash
mkdir -p gpurun_out
( python -m pytest tests/test_gpu_gather.py tests/test_gpu_regular.py tests/test_edit.py -m gpu -q -x ) > gpurun_out/r02_c12_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/r02_c12_tests.log
timeout 600 python tools/bench_aux.py > gpurun_out/r02_c12_bench_aux.jsonl 2> gpurun_out/r02_c12_bench_aux.err; echo "aux exit $?"; grep -E "gather|single_" gpurun_out/r02_c12_bench_aux.jsonl | cut -c1-360; tail -2 gpurun_out/r02_c12_bench_aux.err
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r02_c12_bench.json 2> gpurun_out/r02_c12_bench.err; echo "bench exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/r02_c12_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'])
print(json.dumps(d['configs']['edit_latency']))
print(d['configs']['page_pass'])
"
timeout 200 python tools/repro_race.py --edge 32 --chunks 20 --iters 200 --full-every 5 2>&1 | tail -1
timeout 200 python tools/repro_race.py --edge 64 --chunks 9 --iters 200 --full-every 5 2>&1 | tail -1
