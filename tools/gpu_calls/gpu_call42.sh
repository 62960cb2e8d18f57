#!/bin/bash
# Round-2 GPU call 42 (1 GPU): ncu --set full of the regular kernel on the planet set with the shipped build.
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:regular_extract_decoupled -s 3 -c 1 -o /tmp/r02_regular_planet -f \
    python tools/bench_planet.py --steps 2 --warmup 1 > gpurun_out/r02_c42_ncu_planet.log 2>&1; echo "ncu planet exit $?"
python tools/summarize_ncu.py /tmp/r02_regular_planet.ncu-rep gpurun_out/r02_regular_extract_planet_ncu_full.txt > /dev/null; echo "summary exit $?"
grep -E "time_duration|dram__bytes_(read|write)|issue_active|inst_executed.sum|pipe_(alu|fma|lsu).avg" gpurun_out/r02_regular_extract_planet_ncu_full.txt
