#!/usr/bin/env python3
"""K1 fill against a write-only reference: the density fills at both edges (device time of hvx_fill_density, host
page list already uploaded on the first call) beside a plain device memset of the same number of bytes."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import helio_b200 as H  # noqa: E402

torch.cuda.set_device(0)
stream = torch.cuda.Stream()


def timed(fn, warm=3, iters=15):
    for _ in range(warm):
        fn()
    stream.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def grid(n_axis, ys):
    xs = np.arange(-n_axis // 2, n_axis // 2)
    z, y, x = np.meshgrid(xs, np.array(ys), xs, indexing="ij")
    return np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1).astype(np.int64)


pages = grid(16, range(-8, 8))
for edge in (64, 32):
    nbytes = len(pages) * (edge + 2) ** 3 * 4
    with torch.cuda.stream(stream):
        buf = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        ms = timed(lambda: buf.zero_())
    print(json.dumps({"case": f"memset_{edge}", "bytes": nbytes, "ms": ms, "GBps": nbytes / ms / 1e6}), flush=True)
    del buf
    b = H.ChunkBatchExtractor(0, edge=edge, max_chunks=len(pages), max_vertices=8, max_indices=8)
    b.ctx.set_stream(stream.cuda_stream)
    for kind, name in ((16, "terrain_fbm"), (0, "plane"), (1, "sphere")):
        ms = timed(lambda: b.fill_density(kind, pages))
        print(json.dumps({"case": f"fill_{edge}^3_{name}", "bytes": nbytes, "ms": ms, "GBps": nbytes / ms / 1e6}), flush=True)
    b.close()
