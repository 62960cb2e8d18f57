timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for v in c3n6 altc2 ship; do
  export HVX_LIBRARY=$PWD/variants/lib_$v.so
  python tools/probe_e32.py 16 8 24 16      # all air
  python tools/probe_e32.py 32 -8 8 32      # 16384 pages terrain
  python tools/probe_e32.py 64 -1 0 64      # all surface
  timeout 120 python tools/bench_planet.py 2>/dev/null | python -c "import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('e32 $v planet', round(d['ms_per_step'],4), 'ms', round(d['algorithmic_GBps']), 'GB/s')"
done
for v in base ship; do export HVX_LIBRARY=$PWD/variants/lib_$v.so
for w in terrain surface empty; do timeout 120 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --workload $w 2>/dev/null | tail -1 | python -c "import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('e64 $v $w', round(d['ms_per_step'],4), 'ms', round(d['roofline']['frac'],3))"; done; done
