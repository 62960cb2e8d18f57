set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_ref.json
python bench.py --steps 20 --warmup 3 2>gpurun_out/bench_n1.err | tail -1 > gpurun_out/bench_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > /dev/null 2>&1
python tools/bench_aux.py > gpurun_out/bench_aux.jsonl 2>gpurun_out/bench_aux.err
python tools/bench_planet.py > gpurun_out/planet_n1.json 2>gpurun_out/planet_n1.err
python tools/probe_e32.py 32 -8 8 32 > gpurun_out/probe_e32.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()"
ls -la gpurun_out | tail -12
