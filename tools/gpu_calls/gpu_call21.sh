#!/bin/bash
# Round-2 GPU call 21 (1 GPU): full parity suite, smoke, bench (+ reference arm), launch list, ncu captures of the
# regular kernel (headline / all-surface / planet) with pipe utilisation and SASS mix, stress loop, sanitizers.
mkdir -p gpurun_out
( python -m pytest tests -m gpu -q ) > gpurun_out/r02_c21_tests.log 2>&1; echo "tests exit $?"; tail -6 gpurun_out/r02_c21_tests.log
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_c21_bench.json 2> gpurun_out/r02_c21_bench.err; echo "bench exit $?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_c21_bench_reference.json 2> gpurun_out/r02_c21_ref.err; echo "reference exit $?"; cut -c1-400 gpurun_out/r02_c21_bench_reference.json
python -c "
import json
d=json.loads(open('gpurun_out/r02_c21_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline'], d['e2e']['ms_per_step'], d.get('fill_kernel'), d['cpu_baseline'], d['clocks'], d['gpu_launches'])
for k,v in d['configs'].items(): print(k, json.dumps(v)[:600])
print(json.dumps(d.get('e2e_variants'))[:900])
"
timeout 600 python tools/probe_fill.py 2>&1 | tee gpurun_out/r02_probe_fill.jsonl | cut -c1-160
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-configs > /dev/null 2>&1; echo "launch list exit $?"; wc -l gpurun_out/r02_launches_bench.csv
for wl in terrain surface; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:regular_extract_decoupled -s 4 -c 1 -o /tmp/r02_regular_$wl -f \
      python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --no-configs --workload $wl > gpurun_out/r02_c21_ncu_$wl.log 2>&1; echo "ncu $wl exit $?"
  python tools/summarize_ncu.py /tmp/r02_regular_$wl.ncu-rep gpurun_out/r02_regular_extract_${wl}_ncu_full.txt > /dev/null; echo "summary $wl exit $?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:regular_extract_decoupled -s 3 -c 1 -o /tmp/r02_regular_planet -f \
    python tools/bench_planet.py --steps 2 --warmup 1 > gpurun_out/r02_c21_ncu_planet.log 2>&1; echo "ncu planet exit $?"
python tools/summarize_ncu.py /tmp/r02_regular_planet.ncu-rep gpurun_out/r02_regular_extract_planet_ncu_full.txt > /dev/null; echo "summary planet exit $?"
grep -E "time_duration|dram__bytes_(read|write)|issue_active|pipe_(alu|fma|lsu).avg" gpurun_out/r02_regular_extract_*_ncu_full.txt
timeout 300 python tools/repro_race.py --edge 32 --chunks 140 --iters 60 --full-every 5 2>&1 | tail -1
timeout 300 python tools/repro_race.py --edge 64 --chunks 60 --iters 60 --full-every 5 2>&1 | tail -1
( for tool in memcheck synccheck; do echo "== compute-sanitizer --tool $tool tools/sanitize_small.py"; timeout 900 compute-sanitizer --tool $tool python tools/sanitize_small.py 2>&1 | tail -4; done ) > gpurun_out/r02_sanitizer.txt 2>&1; tail -12 gpurun_out/r02_sanitizer.txt
du -sh gpurun_out
