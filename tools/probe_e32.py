#!/usr/bin/env python3
"""E=32 probe: extraction time of a (gx x gy x gz) grid of 32^3 fBm pages; env HVX_DEBUG_STREAM_ONLY=1/2 for the
bare stream / stream + sign bits.  usage: probe_e32.py gx gy0 gy1 gz [kind]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, '.')
import helio_b200 as H
gx, gy0, gy1, gz = (int(v) for v in sys.argv[1:5])
kind = int(sys.argv[5]) if len(sys.argv) > 5 else 16
xs, ys, zs = np.arange(-gx // 2, gx - gx // 2), np.arange(gy0, gy1), np.arange(-gz // 2, gz - gz // 2)
z, y, x = np.meshgrid(zs, ys, xs, indexing='ij')
pages = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1).astype(np.int64)
n = len(pages)
b = H.ChunkBatchExtractor(0, edge=32, max_chunks=n, max_vertices=12288, max_indices=18432)
s = torch.cuda.Stream(); b.ctx.set_stream(s.cuda_stream)
b.fill_density(kind, pages); d = H.make_descs(n)
for _ in range(3): b.ctx.extract_regular(None, d, n)
s.synchronize(); ts = []
for _ in range(10):
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s); b.ctx.extract_regular(None, d, n); e.record(s); e.synchronize(); ts.append(a.elapsed_time(e))
c = b.counters(n); v = int(c['emitted_vertices'].astype(np.int64).sum()); i = int(c['emitted_indices'].astype(np.int64).sum())
t = float(np.median(ts)); nb = n * 34**3 * 4 + 32 * v + 4 * i
print(f"e32 {os.environ.get('HVX_LIBRARY','ship').split('/')[-1]} mode={os.environ.get('HVX_DEBUG_STREAM_ONLY','0')} pages={n} surface_pages={int((c['emitted_vertices']>0).sum())} "
      f"{t:.4f} ms  {nb / t / 1e6:.0f} GB/s  input-only {n*34**3*4/t/1e6:.0f} GB/s  vertices={v}")
