#!/usr/bin/env python3
"""One GPU, the planet set's rank-0 shard of an 8- (and 4-, 2-) rank strong-scaling run: per-rank step time with the last
wave's pages split over the idle CTAs and with whole pages only (tools/bench_cases.py: planet_shard_probe)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch  # noqa: E402

import bench_cases  # noqa: E402
import helio_b200 as H  # noqa: E402

device = torch.device("cuda:0")
torch.cuda.set_device(device)
for world in ([int(a) for a in sys.argv[1:]] or [8, 4, 2]):
    print(json.dumps(bench_cases.planet_shard_probe(H, torch, device, 0, world=world)), flush=True)
