#!/bin/bash
# Round-2 GPU call 28 (1 GPU): ncu --set full of every kernel other than the regular extractor with the final build (one
# block per kernel in profiles/r02_aux_kernels_ncu.txt), the stress test inside the GPU suite, the edit-frame percentiles again.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stress.py -m gpu -x -q 2>&1 | tail -3
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"fill|heights|transition|gather|meshlet|publish|visibility|brick|edit|records|pack|commit|weld" \
    -o /tmp/r02_aux_kernels -f python tools/profile_kernels.py > gpurun_out/r02_c28_ncu_aux.log 2>&1; echo "ncu aux exit $?"; tail -2 gpurun_out/r02_c28_ncu_aux.log
python tools/summarize_ncu.py --multi /tmp/r02_aux_kernels.ncu-rep gpurun_out/r02_aux_kernels_ncu.txt > /dev/null; grep -E "^kernel|time_duration|per_second" gpurun_out/r02_aux_kernels_ncu.txt | paste - - - | cut -c1-250
python - <<'PY'
import json, sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tools'))
import torch, helio_b200 as H, bench_cases
torch.cuda.set_device(0)
for k in range(3):
    r = bench_cases.edit_latency_case(H, torch, torch.device('cuda', 0), 0)
    print(json.dumps({key: r[key] for key in ('device_ms_p50', 'device_ms_p95', 'device_ms_p99', 'wall_ms_p50', 'wall_ms_p95', 'device_ms_p50_extract_only', 'device_ms_p50_whole_chunk_walks')}))
PY
