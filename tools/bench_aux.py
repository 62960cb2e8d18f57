#!/usr/bin/env python3
"""Secondary measurements for the other BASELINE.json configs (not the headline bench line):

  single_page      config 0: one 32^3 sphere page, per-dispatch latency through the C ABI
  batch32          the reference's page size in batch: 4096 x 32^3 fBm pages
  lod_transition   config 2: coarse pages with all six transition faces, regular + transition kernels
  dirty_edit       config 4: 256 resident 64^3 chunks re-extracted per frame with dirty-microbrick masks
  fill             K1 density fill GB/s per field kind
  gather           SURVEY 8f-1: halo blocks gathered from a resident page atlas, alone and followed by extraction

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
    ex.context.debug_set_mode(0x100)
    r_one = timed(stream, lambda: ex.context.extract_regular(None, descs, 1), 10, 50)
    ex.context.debug_set_mode(0)
    c = ex.counters_buffer()
    out.append({"case": "single_page_32_sphere", **r, "ms_median_one_cta": r_one["ms_median"], "vertices": int(c["emitted_vertices"]), "triangles": int(c["emitted_indices"]) // 3,
                "note": "reference wgpu/RTX 3060: 0.048 ms median (docs/planetary_voxel_extraction_benchmark.md:61)"})
    ex.close()

    # ---- one dense-adversarial 64^3 chunk (~6 vertices per cell): the worst case for "one CTA per chunk" ----------
    cells = 64 ** 3
    ex = H.TransvoxelGpuExtractor(0, H.TransvoxelGpuExtractorConfig(cells * 12, cells * 15), edge=64, debug_records=False)
    ex.context.set_stream(stream.cuda_stream)
    ex.context.fill_density(17, [[1, 2, 3]])
    descs = H.make_descs(1, 7, transition_mask=0x3F)
    r = timed(stream, lambda: ex.context.extract_regular(None, descs, 1), 3, 10)
    ex.context.debug_set_mode(0x100)
    r_whole = timed(stream, lambda: ex.context.extract_regular(None, descs, 1), 3, 10)
    ex.context.debug_set_mode(0)
    c = ex.counters_buffer()
    out.append({"case": "single_chunk_64_dense_random", **r, "ms_median_one_cta": r_whole["ms_median"], "vertices": int(c["emitted_vertices"]),
                "triangles": int(c["emitted_indices"]) // 3, "note": "split over 16 CTAs by z-range vs walked by one CTA"})
    ex.close()
    if "--latency-only" in sys.argv:
        for rec in out:
            print(json.dumps(rec), flush=True)
        return

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
    # optional vertex-reuse output on the same batch (the weld rewrites the meshes, so it is timed behind an extraction)
    n_pages = len(pages)
    r_both = timed(stream, lambda: (b.ctx.extract_regular(None, descs, n_pages), b.ctx.weld_meshes(n_pages)))
    kept = int(b.counters(n_pages)["emitted_vertices"].astype(np.int64).sum())
    out.append({"case": "weld_4096x32^3_fbm", "extract_ms": r["ms_median"], "extract_plus_weld_ms": r_both["ms_median"],
                "weld_ms": r_both["ms_median"] - r["ms_median"], "vertices": v, "kept_vertices": kept, "kept_fraction": kept / max(v, 1),
                "vertex_bytes_saved": 32 * (v - kept)})
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
    # ---- surface gather from a resident page atlas, then extraction of the gathered blocks ----
    side, ys = 32, [-2, -1, 0, 1]
    xs = np.arange(-side // 2, side // 2)
    pages = grid(side, ys)                                          # page index = (iz * len(ys) + iy) * side + ix
    n_pages = len(pages)
    blocks = torch.empty(n_pages * 34 ** 3, dtype=torch.int32, device="cuda")
    fillb = H.ChunkBatchExtractor(0, edge=32, max_chunks=n_pages, max_vertices=8, max_indices=8)
    fillb.ctx.set_stream(stream.cuda_stream)
    fillb.fill_density(16, pages, out_ptr=blocks.data_ptr())        # 34^3 fBm blocks; their interiors are the resident pages
    fillb.ctx.synchronize()
    fillb.close()
    tiles = (side, side, len(ys))                                   # tile (tx, ty, tz) = (ix, iz, iy)
    interior = blocks.view(side, len(ys), side, 34, 34, 34)[:, :, :, 1:33, 1:33, 1:33]   # [iz][iy][ix][cz][cy][cx]
    atlas = torch.empty((tiles[2] * 32, tiles[1] * 32, tiles[0] * 32), dtype=torch.int32, device="cuda")
    atlas.view(tiles[2], 32, tiles[1], 32, tiles[0], 32).copy_(interior.permute(1, 3, 0, 4, 2, 5))
    del blocks, interior
    planet = (0x5A5A5A5A,) * 4
    table = H.PageTable(8192, 64)
    slot_of = {}
    for iy, y in enumerate(ys):
        for iz, z in enumerate(xs):
            for ix, x in enumerate(xs):
                slot = ix + tiles[0] * (iz + tiles[1] * iy)
                slot_of[(int(x), int(y), int(z))] = slot
                table.insert(H.GpuLookupKey(planet, (int(x) * 32, int(y) * 32, int(z) * 32), 0), slot, 5)
    inner = [k for k in slot_of if xs[0] < k[0] < xs[-1] and xs[0] < k[2] < xs[-1] and k[1] in (-1, 0)]
    jobs = np.concatenate([H.gather_job(planet, (x * 32, y * 32, z * 32), 0, 5, 0, slot_of[(x, y, z)], 9) for (x, y, z) in inner])
    nj = len(jobs)
    res = H.residency_uniform(table, tiles, n_pages, 9)
    gctx = H.Context(0, edge=32, max_chunks=nj, max_vertices=12288, max_indices=18432)
    gctx.set_stream(stream.cuda_stream)
    sampler = H.GpuSurfaceSampler(gctx)
    r_g = timed(stream, lambda: sampler.dispatch(res, table, atlas, jobs), 2, 10)
    c = sampler.counters_buffer()
    r_ge = timed(stream, lambda: (sampler.dispatch(res, table, atlas, jobs), sampler.extract()), 2, 10)
    ec = gctx.read(H._ffi.BUF_REGULAR_COUNTERS, 0, nj)
    out.append({"case": f"gather_{nj}x34^3_from_{n_pages}_resident_pages", **r_g, "jobs_per_s": nj / (r_g["ms_median"] * 1e-3),
                "GBps_read_plus_write": 2 * nj * 34 ** 3 * 4 / (r_g["ms_median"] * 1e-3) / 1e9,
                "completed": int(c["completed"].sum()), "page_misses": int(c["page_misses"].sum()),
                "note": "every call also stages the 384 KB page table and the job array from the host"})
    out.append({"case": f"gather_then_extract_{nj}x32^3", **r_ge, "pages_per_s": nj / (r_ge["ms_median"] * 1e-3),
                "vertices": int(ec["emitted_vertices"].astype(np.int64).sum())})
    # the reference's whole per-page pass, batched and device-resident: gather -> extract -> meshlets -> publish
    pub = H.SurfacePublisher(gctx, nj)
    meta = np.zeros(nj, dtype=H.PAGE_META_DTYPE)
    meta["relative_lod0_cell_min"], meta["slot"], meta["generation_low"] = jobs["relative_lod0_cell_min"], np.arange(nj), 5
    sjobs = np.concatenate([pub.surface_job(i, 5) for i in range(nj)])
    chunks = np.arange(nj, dtype=np.uint32)

    def whole_pass():
        sampler.dispatch(res, table, atlas, jobs)
        sampler.extract()
        gctx.build_meshlets(nj, 0)
        pub.publish(sjobs, chunks, meta)

    r_all = timed(stream, whole_pass, 2, 10)
    fb = pub.feedback()
    out.append({"case": f"gather_extract_meshlets_publish_{nj}x32^3", **r_all, "pages_per_s": nj / (r_all["ms_median"] * 1e-3),
                "published_per_call": int(fb["published_jobs"]) // 12, "note": "the reference runs this pass for one page per frame"})
    pub.close()
    gctx.close()

    # ---- legacy 8^3-brick marching cubes, 65,536 bricks of a rolling height field ------------
    gx, gy, gz = 64, 16, 64
    X, Z = np.meshgrid(np.arange(gx * 8 + 1), np.arange(gz * 8 + 1), indexing="xy")       # [z][x]
    height = (64 + 30 * np.sin(X * 0.021) * np.cos(Z * 0.017) + 9 * np.sin(X * 0.13 + Z * 0.11)).astype(np.int32)
    vol = (np.arange(gy * 8 + 1)[None, :, None] < height[:, None, :]).astype(np.uint8) * 3   # [z][y][x]
    sz, sy, sx = vol.strides
    view = np.lib.stride_tricks.as_strided(vol, (gz, gy, gx, 9, 9, 9), (8 * sz, 8 * sy, 8 * sx, sz, sy, sx))
    nb = gx * gy * gz
    padded = np.zeros((nb, 732), dtype=np.uint8)
    padded[:, :729] = view.reshape(nb, 729)
    words = torch.from_numpy(padded.view("<u4").reshape(-1).astype(np.int32)).cuda()
    ex = H.VoxelMeshExtractor(0, max_bricks=nb, max_dirty=nb)
    ex.ctx.set_stream(stream.cuda_stream)
    meta = np.zeros(nb, dtype=H.BRICK_META_DTYPE)
    meta["data_offset"] = np.arange(nb, dtype=np.uint32) * 183
    ex.write_brick_meta(meta)
    ex.write_voxel_data(words)
    dirty = np.zeros(nb, dtype=H.DIRTY_BRICK_DTYPE)
    dirty["brick_slot"] = np.arange(nb)
    bz, by, bx = np.unravel_index(np.arange(nb), (gz, gy, gx))
    dirty["origin_size"] = np.stack([bx * 8, by * 8, bz * 8, np.ones(nb)], axis=1).astype(np.float32)
    r_host = timed(stream, lambda: ex.extract(dirty), 2, 10)      # lists staged from pageable host memory every call
    d_dirty = torch.from_numpy(dirty.view(np.uint8).reshape(-1).copy()).cuda()
    d_meta = torch.from_numpy(meta.view(np.uint8).reshape(-1).copy()).cuda()
    r_b = timed(stream, lambda: ex.extract(d_dirty, d_meta), 2, 10)
    assert ex.rejected() == 0
    desc = ex.descriptors()
    entries = int(desc["vertex_count"].astype(np.int64).sum())
    bbytes = nb * (732 + 32 + 20) + 36 * entries
    out.append({"case": f"legacy_brick_marching_cubes_{nb}x8^3", **r_b, "bricks_per_s": nb / (r_b["ms_median"] * 1e-3),
                "cells_per_s": nb * 512 / (r_b["ms_median"] * 1e-3), "algorithmic_GBps": bbytes / (r_b["ms_median"] * 1e-3) / 1e9,
                "vertices": entries, "surface_bricks": int((desc["vertex_count"] > 0).sum()),
                "clamped_bricks": int((desc["vertex_count"] == 2048).sum()),
                "ms_median_host_lists": r_host["ms_median"],
                "note": "brick metadata and dirty list device-resident; ms_median_host_lists stages the 2.5 MB from pageable host memory every call"})
    ex.close()
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
