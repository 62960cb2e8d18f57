#!/bin/bash
# Before every gpurun call: everything that can fail without a GPU fails here.
set -e
cd "$(dirname "$0")/.."
python -m compileall -q bench.py __graft_entry__.py tools tests helio_b200 oracle > /dev/null
python -c "import sys; sys.path.insert(0,'tools'); import bench_cases, bench" 
(cd helio_b200/csrc && make -j8 2>&1 | grep -E "rror" && exit 1 || true)
# the stress variant the GPU suite loads links the same objects: keep it in step (a stale one lacks new exports)
(cd helio_b200/csrc && make ../../build/variants/libhvx_jitter2.so 2>&1 | grep -E "rror" && exit 1 || true)
python -m pytest tests -x -q -m "not gpu" 2>&1 | tail -2
