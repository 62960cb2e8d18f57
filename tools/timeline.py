#!/usr/bin/env python3
"""Where a short launch of the regular kernel spends its time: per-CTA time stamps from the HVX_TIMELINE variant build.

  make -C helio_b200/csrc ../../build/variants/libhvx_timeline.so
  python tools/timeline.py            # single page, planet shard of 8 ranks, edge-64 batch of 256 chunks

Events per CTA (regular_extract.cu): CTA up, ticket drawn, first slab of a walk landed, chunk records written, producer
exit.  Times are globaltimer nanoseconds; everything is reported relative to the first CTA's "up" event."""
import ctypes as C
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
os.environ.setdefault("HVX_LIBRARY", str(ROOT / "build" / "variants" / "libhvx_timeline.so"))
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import torch  # noqa: E402

import bench_cases  # noqa: E402
import helio_b200 as H  # noqa: E402

lib = C.CDLL(os.environ["HVX_LIBRARY"])
lib.hvx_debug_timeline_read.argtypes = [C.c_void_p, C.c_int]
KINDS = {0: "up", 1: "ticket", 2: "first_slab", 3: "chunk_end", 4: "exit"}


def read(reset):
    raw = np.zeros((1024, 64), dtype=np.uint64)
    assert lib.hvx_debug_timeline_read(raw.ctypes.data, int(reset)) == 0
    return raw


def pct(values, q):
    return float(np.percentile(np.asarray(values, dtype=np.float64), q)) if len(values) else None


def summarize(raw, label):
    ctas = []
    for row in raw:
        n = int(min(row[0], 63))
        if n == 0:
            continue
        ev = [(int(w & ((1 << 44) - 1)), int((w >> 44) & 15), int(w >> 48)) for w in row[1:1 + n]]
        ctas.append(sorted(ev))
    t0 = min(e[0] for c in ctas for e in c if e[1] == 0)
    up = [c_ev[0] - t0 for c in ctas for c_ev in c if c_ev[1] == 0]
    end = max(e[0] for c in ctas for e in c) - t0
    first_ticket, first_slab, walks, tails, first_walk, later_walks = [], [], [], [], [], []
    for c in ctas:
        u = next(e[0] for e in c if e[1] == 0)
        tickets = [e[0] for e in c if e[1] == 1]
        slabs = [e[0] for e in c if e[1] == 2]
        ends = [e[0] for e in c if e[1] == 3]
        if tickets:
            first_ticket.append(tickets[0] - u)
        if slabs:
            first_slab.append(slabs[0] - u)
        walks.append(len(ends))
        # a walk = first slab landed -> its chunk records written (walks of a CTA do not overlap in their ends)
        for k, (a, b) in enumerate(zip(slabs, ends)):
            (first_walk if k == 0 else later_walks).append(b - a)
        if ends:
            tails.append(end + t0 - ends[-1])
    out = {
        "case": label, "ctas": len(ctas), "span_us": end / 1e3,
        "cta_up_us_p50_max": [pct(up, 50) / 1e3, max(up) / 1e3],
        "up_to_first_ticket_us_p50": pct(first_ticket, 50) / 1e3 if first_ticket else None,
        "up_to_first_slab_us_p50_max": [pct(first_slab, 50) / 1e3, max(first_slab) / 1e3] if first_slab else None,
        "walks_per_cta_min_max": [int(min(walks)), int(max(walks))],
        "first_walk_us_p50": pct(first_walk, 50) / 1e3 if first_walk else None,
        "later_walk_us_p50_p95": [pct(later_walks, 50) / 1e3, pct(later_walks, 95) / 1e3] if later_walks else None,
        "idle_at_the_end_us_p50_max": [pct(tails, 50) / 1e3, max(tails) / 1e3] if tails else None,
        "idle_at_the_end_share_of_cta_time": float(np.mean(tails) / end) if tails else None,
    }
    print(json.dumps(out), flush=True)
    return out


def run(label, dispatch, sync, warm=3):
    for _ in range(warm):
        dispatch()
    sync()
    read(True)
    dispatch()
    sync()
    return summarize(read(True), label)


def main():
    torch.cuda.set_device(0)
    device = torch.device("cuda", 0)
    # one 32^3 sphere page (split over 16 CTAs)
    ex = H.TransvoxelGpuExtractor(0, debug_records=False)
    ex.context.fill_density(1, [[0, 0, 0]])
    descs = H.make_descs(1, 7)
    run("single_page_32_sphere", lambda: ex.context.extract_regular(None, descs, 1), ex.context.synchronize)
    ex.context.debug_set_mode(0x100)
    run("single_page_32_sphere_one_cta", lambda: ex.context.extract_regular(None, descs, 1), ex.context.synchronize)
    ex.close()
    # the planet set: rank 0's shard of an 8-rank run, and the whole set
    pages_all, lods_all, masks_all = bench_cases.planet_page_set(H, 256)
    plane = int(H.ExtractionFixtureKind.Plane)
    cost = np.array([H.chunk_cost(32, int(m)) for m in masks_all], dtype=np.uint64)
    for world in (8, 4):
        mine = np.flatnonzero(H.partition_chunks(cost, world) == 0)
        b = H.ChunkBatchExtractor(0, edge=32, max_chunks=len(mine), max_vertices=4608, max_indices=6912)
        b.ctx.fill_density(plane, np.ascontiguousarray(pages_all[mine]), np.ascontiguousarray(lods_all[mine]))
        d = H.make_descs(len(mine), transition_mask=[int(m) for m in masks_all[mine]])
        run(f"planet_shard_of_{world}_{len(mine)}_pages", lambda: b.ctx.extract_regular(None, d, len(mine)), b.ctx.synchronize)
        b.close()
    # 256 terrain chunks of 64^3 re-extracted in full (two waves of 148 CTAs)
    xs = np.arange(-8, 8)
    pages = np.array([[x, -1, z] for z in xs for x in xs], dtype=np.int64)
    b = H.ChunkBatchExtractor(0, edge=64, max_chunks=256, max_vertices=65_536, max_indices=98_304)
    b.fill_density(16, pages)
    d = H.make_descs(256)
    run("256_terrain_chunks_64_full", lambda: b.ctx.extract_regular(None, d, 256), b.ctx.synchronize)
    b.close()


if __name__ == "__main__":
    main()
