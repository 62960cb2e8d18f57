#!/bin/bash
# Round-2 GPU call 33 (1 GPU): final pass of the round: full parity suite, smoke, bench (+ reference arm), launch list.
mkdir -p gpurun_out
( python -m pytest tests -m gpu -q ) > gpurun_out/r02_c33_tests.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/r02_c33_tests.log
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_c33_bench.json 2> gpurun_out/r02_c33_bench.err; echo "bench exit $?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_c33_bench_reference.json 2>/dev/null; echo "reference exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/r02_c33_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'], d['e2e'].get('h2d_frac'), d.get('fill_kernel'), d['gpu_launches'], d['clocks'])
for k,v in d['configs'].items(): print(k, json.dumps(v)[:330])
r=json.loads(open('gpurun_out/r02_c33_bench_reference.json').read().strip().splitlines()[-1]); print('reference', r['value'], r['ms_per_step'])
"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-configs > /dev/null 2>&1; echo "launch list exit $?"; wc -l gpurun_out/r02_launches_bench.csv
