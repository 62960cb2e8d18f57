run() { HVX_DEBUG_FLAGS=$1 HVX_REGULAR_VARIANT=$2 timeout 120 python bench.py --steps 5 --warmup 2 --no-e2e --no-cpu --workload $3 2>/dev/null | tail -1 | python -c "import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('flags $1 variant $2 $3', round(d['ms_per_step'],4), round(d['roofline']['achieved']), round(d['roofline']['frac'],3))
except Exception as e: print('flags $1 variant $2 $3 FAILED')"; }
for f in 0 1; do for v in 2; do for w in terrain empty surface; do run $f $v $w; done; done; done
