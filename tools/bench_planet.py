#!/usr/bin/env python3
"""bench_planet.py -- BASELINE configs[3]: the planet-scale page set, sharded across 1/2/4/8 GPUs.

Workload: `HorizonLodFixturePlan::build_with_minimum_lod(focus, 11, minimum_lod, 192)`
for F camera foci drawn exactly like the reference's randomized horizon test (xorshift64, seed
0x4D595DF4D0F33173, focus = [+-rand % 127,420,000, -1, +-rand % 127,420,000], minimum_lod = rand % 6;
PV/src/lod_topology.rs:507-520,606-619) at the reference's page edge 32 with the Plane field (what
planet_voxel_demo extracts).  Every page of every plan is one chunk: regular extraction with its
coarse-owned `transition_mask` (secondary positions), plus a transition extraction for every page
whose mask is not zero.

The GLOBAL page list is fixed (strong scaling): `hvx_partition_chunks` (LPT by `hvx_chunk_cost`)
assigns pages to ranks, every rank fills + extracts its own shard, there is no data-path collective.
Totals (vertices / indices, regular and transition) are all-reduced afterwards only to show that the
sharded result is the one-GPU result.

  python tools/bench_planet.py [--foci 256]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_planet.py
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

EDGE = 32
MASK64 = (1 << 64) - 1


def next_random(state):
    """xorshift64 of PV/src/lod_topology.rs:613-618."""
    state ^= (state << 13) & MASK64
    state ^= state >> 7
    state ^= (state << 17) & MASK64
    return state


def planet_page_set(n_foci):
    """(page_xyz [n,3] int64, lod [n] uint8, transition_mask [n] uint32, plans) for n_foci horizon plans."""
    import helio_b200 as H
    state = 0x4D595DF4D0F33173
    pages, lods, masks = [], [], []

    def coordinate():
        nonlocal state
        state = next_random(state)
        magnitude = state % 127_420_000
        state = next_random(state)
        return magnitude if state & 1 == 0 else -magnitude

    for _ in range(n_foci):
        focus = [coordinate(), -1, coordinate()]
        state = next_random(state)
        minimum_lod = state % 6
        plan = H.HorizonLodFixturePlan.build_with_minimum_lod(focus, 11, minimum_lod, 192, EDGE)
        for key, mask in plan.topology().transition_masks().items():
            pages.append(key.page_xyz)
            lods.append(key.lod)
            masks.append(mask)
    return (np.array(pages, dtype=np.int64), np.array(lods, dtype=np.uint8), np.array(masks, dtype=np.uint32))


class quiet_stdout:
    """Send file descriptor 1 to stderr for a while: torch prints "NCCL version ..." on stdout when the
    communicator is created, and stdout is reserved for the one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self._saved, 1)
        os.close(self._saved)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--foci", type=int, default=256)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    import helio_b200 as H

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        with quiet_stdout():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            torch.cuda.set_device(local_rank)
            dist.barrier()
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)

    pages_all, lods_all, masks_all = planet_page_set(args.foci)
    costs = np.array([H.chunk_cost(EDGE, int(m)) for m in masks_all], dtype=np.uint64)
    owner = H.partition_chunks(costs, world)
    mine = np.flatnonzero(owner == rank)
    pages, lods, masks = np.ascontiguousarray(pages_all[mine]), np.ascontiguousarray(lods_all[mine]), masks_all[mine]
    n = len(mine)
    seam = np.flatnonzero(masks != 0)                 # pages that own at least one transition face (lod >= 1)
    nt = len(seam)

    # a plane page has exactly 32 x 32 surface cells: 4,096 vertices / 6,144 indices
    batch = H.ChunkBatchExtractor(local_rank, edge=EDGE, max_chunks=n, max_vertices=4608, max_indices=6912,
                                  max_transition_vertices=2048, max_transition_indices=6144)
    ctx = batch.ctx
    stream = torch.cuda.Stream(device=device)
    ctx.set_stream(stream.cuda_stream)
    plane = int(H.ExtractionFixtureKind.Plane)
    ctx.fill_density(plane, pages, lods)
    if nt:
        ctx.fill_slabs(plane, np.ascontiguousarray(pages[seam]), np.ascontiguousarray(lods[seam]))
    d_reg = H.make_descs(n, transition_mask=[int(m) for m in masks])
    d_tr = H.make_descs(max(nt, 1), transition_mask=[int(m) for m in masks[seam]] if nt else 0)

    def step():
        ctx.extract_regular(None, d_reg, n)
        if nt:
            ctx.extract_transition(None, d_tr, nt)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    for _ in range(args.warmup):
        step()
    barrier()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(args.steps):
        step()
    e.record(stream)
    barrier()
    ms = a.elapsed_time(e) / args.steps
    # where the step goes: the two launches timed apart (same stream, back to back)
    parts = []
    for fn in (lambda: ctx.extract_regular(None, d_reg, n), (lambda: ctx.extract_transition(None, d_tr, nt)) if nt else None):
        if fn is None:
            parts.append(0.0)
            continue
        a2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a2.record(stream)
        for _ in range(args.steps):
            fn()
        e2.record(stream)
        e2.synchronize()
        parts.append(a2.elapsed_time(e2) / args.steps)

    rc = batch.counters(n)
    totals = [int(rc["emitted_vertices"].astype(np.int64).sum()), int(rc["emitted_indices"].astype(np.int64).sum()), 0, 0,
              int((rc["vertex_overflow"] | rc["index_overflow"]).sum()), n, nt,
              int(np.array([bin(int(m)).count("1") for m in masks]).sum())]
    if nt:
        tc = batch.transition_counters(nt)
        totals[2] = int(tc["emitted_vertices"].astype(np.int64).sum())
        totals[3] = int(tc["emitted_indices"].astype(np.int64).sum())
        totals[4] += int((tc["vertex_overflow"] | tc["index_overflow"]).sum())
    alg_bytes = (n * (EDGE + 2) ** 3 * 4 + 32 * totals[0] + 4 * totals[1]
                 + totals[7] * 12 * (2 * EDGE + 3) ** 2 + 32 * totals[2] + 4 * totals[3])
    t = torch.tensor(totals + [alg_bytes], dtype=torch.int64, device=device)
    tmax = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    t = [int(x) for x in t.tolist()]
    ms = float(tmax.item())
    if rank == 0:
        n_all = len(pages_all)
        print(json.dumps({
            "case": "planet_scale_horizon_plans_32^3_plane", "n_gpus": world, "foci": args.foci, "pages": n_all,
            "pages_with_transition_faces": t[6], "transition_faces": t[7], "lod_histogram": np.bincount(lods_all).tolist(),
            "ms_per_step": ms, "ms_regular_rank0": parts[0], "ms_transition_rank0": parts[1], "pages_per_s": n_all / (ms * 1e-3), "cells_per_s": n_all * EDGE ** 3 / (ms * 1e-3),
            "algorithmic_GBps": t[8] / (ms * 1e-3) / 1e9,
            "regular_vertices": t[0], "regular_indices": t[1], "transition_vertices": t[2], "transition_indices": t[3],
            "overflowed": t[4], "scaling": "strong (fixed global page list, LPT partition, no data-path collective)",
            "pages_per_rank_min_max": [int(np.bincount(owner, minlength=world).min()), int(np.bincount(owner, minlength=world).max())],
        }))
    batch.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
