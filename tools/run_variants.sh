timeout 300 python -m pytest tests/test_gpu_regular.py -x -q -m gpu 2>&1 | tail -3
run() { HVX_REGULAR_VARIANT=$1 HVX_DEBUG_STREAM_ONLY=$2 timeout 120 python bench.py --steps 5 --warmup 2 --no-e2e --no-cpu --workload $3 2>/dev/null | tail -1 | python -c "import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variant $1 mode $2 $3', round(d['ms_per_step'],4), round(d['roofline']['achieved']), round(d['roofline']['frac'],3))
except Exception as e: print('variant $1 mode $2 $3 FAILED')"; }
for v in 0 1 2 3 4; do run $v 0 terrain; done
run 0 0 empty; run 0 0 surface; run 1 0 surface; run 0 2 empty; run 0 1 empty
