#!/usr/bin/env python3
"""Secondary measurements for the other BASELINE.json configs (not the headline bench line):

  single_page      config 0: one 32^3 sphere page, per-dispatch latency through the C ABI
  batch32          the reference's page size in batch: 4096 x 32^3 fBm pages
  lod_transition   config 2: coarse pages with all six transition faces, regular + transition kernels
  dirty_edit       config 4: 256 resident 64^3 chunks re-extracted per frame with dirty-microbrick masks
  fill             K1 density fill GB/s per field kind

Prints one JSON object per line; CUDA events on the ctx stream, inputs resident in HBM.
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import helio_b200 as H  # noqa: E402


def timed(stream, fn, warmup=3, iters=20):
    for _ in range(warmup):
        fn()
    stream.synchronize()
    times = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        b.synchronize()
        times.append(a.elapsed_time(b))
    t = np.array(times)
    return {"ms_median": float(np.median(t)), "ms_p95": float(np.percentile(t, 95)), "ms_min": float(t.min())}


def grid(n_axis, y_layers, edge_unused=None):
    xs = np.arange(-n_axis // 2, n_axis // 2)
    z, y, x = np.meshgrid(xs, np.array(y_layers), xs, indexing="ij")
    return np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1).astype(np.int64)


def main():
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    out = []

    # ---- config 0: single 32^3 sphere page, the reference's dispatch shape ----------------------
    ex = H.TransvoxelGpuExtractor(0, debug_records=False)
    ex.context.set_stream(stream.cuda_stream)
    ex.context.fill_density(1, [[0, 0, 0]])
    descs = H.make_descs(1, 7)
    r = timed(stream, lambda: ex.context.extract_regular(None, descs, 1), 10, 50)
    c = ex.counters_buffer()
    out.append({"case": "single_page_32_sphere", **r, "vertices": int(c["emitted_vertices"]), "triangles": int(c["emitted_indices"]) // 3,
                "note": "reference wgpu/RTX 3060: 0.048 ms median (docs/planetary_voxel_extraction_benchmark.md:61)"})
    ex.close()

    # ---- batch of the reference's page size ---------------------------------------------------------
    pages = grid(16, range(-8, 8))
    b = H.ChunkBatchExtractor(0, edge=32, max_chunks=len(pages), max_vertices=12288, max_indices=18432)
    b.ctx.set_stream(stream.cuda_stream)
    b.fill_density(16, pages)
    descs = H.make_descs(len(pages))
    r = timed(stream, lambda: b.ctx.extract_regular(None, descs, len(pages)))
    cnt = b.counters(len(pages))
    v, i = int(cnt["emitted_vertices"].astype(np.int64).sum()), int(cnt["emitted_indices"].astype(np.int64).sum())
    nbytes = len(pages) * 34 ** 3 * 4 + 32 * v + 4 * i
    out.append({"case": "batch_4096x32^3_fbm", **r, "cells_per_s": len(pages) * 32 ** 3 / (r["ms_median"] * 1e-3),
                "pages_per_s": len(pages) / (r["ms_median"] * 1e-3), "algorithmic_GBps": nbytes / (r["ms_median"] * 1e-3) / 1e9,
                "vertices": v, "overflowed": int((cnt["vertex_overflow"] | cnt["index_overflow"]).sum())})
    for kind, name in [(0, "plane"), (1, "sphere"), (16, "terrain_fbm"), (17, "dense_random")]:
        r = timed(stream, lambda: b.fill_density(kind, pages), 2, 10)
        out.append({"case": f"fill_32^3_{name}", **r, "GBps": len(pages) * 34 ** 3 * 4 / (r["ms_median"] * 1e-3) / 1e9})
    b.close()

    # ---- config 2: LOD seam path, coarse pages with six transition faces --------------------------
    pages = grid(32, [-1])                                  # 1024 coarse (lod 1) surface pages
    n = len(pages)
    b = H.ChunkBatchExtractor(0, edge=64, max_chunks=n, max_vertices=49152, max_indices=73728,
                              max_transition_vertices=16384, max_transition_indices=49152)
    b.ctx.set_stream(stream.cuda_stream)
    lods = np.ones(n, dtype=np.uint8)
    b.fill_density(16, pages, lods)
    b.fill_slabs(16, pages, lods)
    dr = H.make_descs(n, transition_mask=0x3F)
    r_reg = timed(stream, lambda: b.ctx.extract_regular(None, dr, n), 2, 10)
    r_tr = timed(stream, lambda: b.ctx.extract_transition(None, dr, n), 2, 10)
    tc = b.transition_counters(n)
    tv, ti = int(tc["emitted_vertices"].astype(np.int64).sum()), int(tc["emitted_indices"].astype(np.int64).sum())
    tbytes = n * 6 * 12 * 131 * 131 + 32 * tv + 4 * ti
    out.append({"case": "lod_seam_1024x64^3_mask0x3f_regular", **r_reg})
    out.append({"case": "lod_seam_1024x64^3_mask0x3f_transition", **r_tr, "transition_cells_per_s": n * 6 * 64 * 64 / (r_tr["ms_median"] * 1e-3),
                "algorithmic_GBps": tbytes / (r_tr["ms_median"] * 1e-3) / 1e9, "vertices": tv,
                "overflowed": int((tc["vertex_overflow"] | tc["index_overflow"]).sum())})
    for kind, name in [(0, "plane"), (16, "terrain_fbm")]:
        r = timed(stream, lambda: b.fill_density(kind, pages, lods), 2, 10)
        out.append({"case": f"fill_64^3_{name}", **r, "GBps": n * 66 ** 3 * 4 / (r["ms_median"] * 1e-3) / 1e9})
    b.close()

    # ---- config 4: incremental edits, 256 dirty chunks per frame ------------------------------------
    pages = grid(16, [-1])
    n = len(pages)
    b = H.ChunkBatchExtractor(0, edge=64, max_chunks=n, max_vertices=49152, max_indices=73728)
    b.ctx.set_stream(stream.cuda_stream)
    b.fill_density(16, pages)
    rng = np.random.default_rng(1)
    frames = []
    for _ in range(220):
        masks = []
        for _ in range(n):                                   # a 1.5 m sphere edit touches a 2x2x2 block of 16^3 microbricks
            mx, my, mz = rng.integers(0, 3, 3)
            m = 0
            for dz in (0, 1):
                for dy in (0, 1):
                    for dx in (0, 1):
                        m |= 1 << ((mx + dx) + 4 * (my + dy) + 16 * (mz + dz))
            masks.append(m)
        frames.append(H.make_descs(n, dirty_microbricks=masks))
    full = H.make_descs(n)
    times = []
    for f, d in enumerate(frames):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        b.ctx.extract_regular(None, d, n)
        e.record(stream)
        e.synchronize()
        if f >= 20:
            times.append(a.elapsed_time(e))
    t = np.array(times)
    r_full = timed(stream, lambda: b.ctx.extract_regular(None, full, n), 3, 30)
    out.append({"case": "dirty_edit_256x64^3_per_frame", "ms_p50": float(np.percentile(t, 50)), "ms_p95": float(np.percentile(t, 95)),
                "ms_p99": float(np.percentile(t, 99)), "frames": len(t), "full_reextract": r_full})
    b.close()
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
