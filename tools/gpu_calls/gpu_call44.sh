#!/bin/bash
# Round-2 GPU call 44 (1 GPU): the stress tests of the GPU suite with the relinked variant.
timeout 400 python -m pytest tests/test_gpu_stress.py tests/test_gpu_regular.py -m gpu -x -q -k "stress or jitter or hint or split" 2>&1 | tail -3
