# A/B of regular-kernel variants: usage  bash tools/run_ab.sh "0 10" "terrain empty surface"
run() { HVX_REGULAR_VARIANT=$1 timeout 180 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --workload $2 2>/dev/null | tail -1 | python -c "import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variant $1 $2', round(d['ms_per_step'],4), round(d['roofline']['achieved']), round(d['roofline']['frac'],3))
except Exception as e: print('variant $1 $2 FAILED')"; }
for v in ${1:-0 10}; do for w in ${2:-terrain empty surface}; do run $v $w; done; done
