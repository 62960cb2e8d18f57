#!/bin/bash
# Round-2 GPU call 20 (1 GPU): K1 fill vs a plain memset (write-only reference) and ncu of the fill kernels at both edges.
mkdir -p gpurun_out
timeout 300 python tools/probe_fill.py 2>&1 | tee gpurun_out/r02_probe_fill.jsonl | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fill_terrain|fill_fields|terrain_heights" -s 30 -c 24 -o /tmp/r02_fill -f python tools/probe_fill.py > gpurun_out/r02_c20_ncu.log 2>&1; echo "ncu exit $?"
python tools/summarize_ncu.py --multi /tmp/r02_fill.ncu-rep gpurun_out/r02_fill_kernels_ncu.txt | head -90
timeout 300 python -m pytest tests/test_gpu_weld.py -m gpu -x -q 2>&1 | tail -3
