#!/bin/bash
# Round-2 GPU call 16 (1 GPU): queue-state fast skip A/B (headline + all-surface), kernel times of the latency cases and of
# the per-rank planet shard (ncu launch list: where the fixed costs sit).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_regular.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --no-configs --steps 20 --warmup 5 > gpurun_out/r02_c16_bench.json 2> gpurun_out/r02_c16_bench.err; echo "bench exit $?"
timeout 300 python bench.py --workload surface --no-configs --steps 20 --warmup 5 > gpurun_out/r02_c16_surface.json 2> gpurun_out/r02_c16_surface.err; echo "surface exit $?"
python -c "
import json
for f in ('r02_c16_bench','r02_c16_surface'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'])
"
timeout 300 python tools/probe_planet_shard.py 8 2>&1 | tail -2 | cut -c1-900
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_c16_launches_shard.csv python tools/probe_planet_shard.py 8 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r02_c16_launches_shard.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.defaultdict(list)
for r in rows:
    agg[r[4][:90]].append(float(r[-1].replace(',','')))
for k,v in agg.items():
    v=sorted(v); print(len(v), 'median %.1f us'%(v[len(v)//2]/1000 if v[len(v)//2]>1000 else v[len(v)//2]), k)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c16_launches_latency.csv python tools/bench_aux.py --latency-only > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r02_c16_launches_latency.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.defaultdict(list)
for r in rows:
    agg[r[4][:90]].append(float(r[-1].replace(',','')))
for k,v in agg.items():
    v=sorted(v); print(len(v), 'median', v[len(v)//2], 'min', v[0], k)
print(rows[0]); 
PY
head -3 gpurun_out/r02_c16_launches_latency.csv | cut -c1-400
