#!/bin/bash
# Round-2 GPU call 9 (1 GPU): full parity suite, latency configurations, quick bench.
mkdir -p gpurun_out
( python -m pytest tests -m gpu -q ) > gpurun_out/r02_c9_tests.log 2>&1; echo "tests exit $?"; tail -6 gpurun_out/r02_c9_tests.log
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python tools/bench_aux.py > gpurun_out/r02_c9_bench_aux.jsonl 2> gpurun_out/r02_c9_bench_aux.err; echo "aux exit $?"; grep -E "single_page|dirty_edit|gather" gpurun_out/r02_c9_bench_aux.jsonl | cut -c1-300; tail -3 gpurun_out/r02_c9_bench_aux.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_c9_bench.json 2> gpurun_out/r02_c9_bench.err; echo "bench exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/r02_c9_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['fill_kernel'])
print(json.dumps(d['configs']['edit_latency']))
print(d['configs']['planet']['ms_per_step'], d['configs']['lod_seam']['ms_both'], d['configs']['page_pass'])
"
timeout 300 python tools/repro_race.py --edge 32 --chunks 140 --iters 100 --full-every 5 2>&1 | tail -1
timeout 300 python tools/repro_race.py --edge 64 --chunks 60 --iters 100 --full-every 5 2>&1 | tail -1
timeout 300 python tools/repro_race.py --edge 64 --chunks 1184 --iters 20 2>&1 | tail -1
