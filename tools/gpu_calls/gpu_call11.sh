#!/bin/bash
# Round-2 GPU call 11 (1 GPU): gather rewrite parity + timing; ncu captures of the regular kernel (headline, all-surface,
# planet set), summarised on the box (reports stay on the box: too large to travel); launch list of the bench command.
mkdir -p gpurun_out
( python -m pytest tests/test_gpu_gather.py tests/test_gpu_regular.py -m gpu -q -x ) > gpurun_out/r02_c11_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/r02_c11_tests.log
timeout 600 python tools/bench_aux.py 2>/dev/null | grep -E "gather|single_page" | cut -c1-330
# launch list of the bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-configs > /dev/null 2>&1; echo "launch list exit $?"; wc -l gpurun_out/r02_launches_bench.csv
# ncu --set full: one launch of the extraction kernel per workload
for wl in terrain surface; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:regular_extract_decoupled -s 4 -c 1 -o /tmp/r02_regular_$wl -f \
      python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --no-configs --workload $wl > gpurun_out/r02_c11_ncu_$wl.log 2>&1; echo "ncu $wl exit $?"
  python tools/summarize_ncu.py /tmp/r02_regular_$wl.ncu-rep gpurun_out/r02_regular_extract_${wl}_ncu_full.txt > /dev/null; echo "summary $wl exit $?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:regular_extract_decoupled -s 3 -c 1 -o /tmp/r02_regular_planet -f \
    python tools/bench_planet.py --steps 2 --warmup 1 > gpurun_out/r02_c11_ncu_planet.log 2>&1; echo "ncu planet exit $?"
python tools/summarize_ncu.py /tmp/r02_regular_planet.ncu-rep gpurun_out/r02_regular_extract_planet_ncu_full.txt > /dev/null; echo "summary planet exit $?"
head -24 gpurun_out/r02_regular_extract_surface_ncu_full.txt
grep -E "time_duration|dram__bytes_(read|write)|issue_active" gpurun_out/r02_regular_extract_terrain_ncu_full.txt gpurun_out/r02_regular_extract_planet_ncu_full.txt
du -sh gpurun_out
