# E=32 emission-warp count A/B (variants/lib_*.so) + the parity suite on the shipping library
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for v in base nw10 nw12; do
  export HVX_LIBRARY=$PWD/variants/lib_$v.so
  timeout 120 python tools/bench_planet.py 2>/dev/null | python -c "import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v planet', round(d['ms_per_step'],4), 'ms', round(d['algorithmic_GBps']), 'GB/s')"
  timeout 200 bash tools/run_e32.sh 2>&1 | grep "e32 variant None" | sed "s/^/$v /"
done
unset HVX_LIBRARY
for w in terrain surface; do timeout 120 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --workload $w 2>/dev/null | tail -1 | python -c "import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('base $w', round(d['ms_per_step'],4), 'ms', round(d['roofline']['frac'],3))"; done
