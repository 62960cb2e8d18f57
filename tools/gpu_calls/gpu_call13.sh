#!/bin/bash
mkdir -p gpurun_out
( python -m pytest tests/test_gpu_gather.py tests/test_gpu_regular.py -m gpu -q -x ) > gpurun_out/r02_c13_tests.log 2>&1; echo "tests exit $?"; tail -2 gpurun_out/r02_c13_tests.log
timeout 600 python tools/bench_aux.py > gpurun_out/r02_bench_aux.jsonl 2> gpurun_out/r02_c13_bench_aux.err; echo "aux exit $?"; grep -E "gather_1800|single_" gpurun_out/r02_bench_aux.jsonl | cut -c1-300
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r02_c13_bench.json 2> gpurun_out/r02_c13_bench.err; echo "bench exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/r02_c13_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'])
print(json.dumps(d['configs']['edit_latency']))
"
