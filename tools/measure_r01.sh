set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 2>gpurun_out/bench_n1.err | tail -1 > gpurun_out/bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:regular_extract -s 1 -c 1 -f -o gpurun_out/prof_final python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
python tools/bench_aux.py > gpurun_out/bench_aux.jsonl 2>gpurun_out/bench_aux.err
ls -la gpurun_out | tail -8
