#!/usr/bin/env python3
"""Optional final gather of the planet-scale page set over NCCL: every rank extracts its LPT shard,
then helio_b200.distributed.gather_meshes brings every mesh to rank 0 in global page order.  Reports the time and
the bytes that crossed NVLink.  Not part of bench.py's `value`: the data path has no collective.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_gather.py
"""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))


def main():
    import torch
    import torch.distributed as dist
    import helio_b200 as H
    import bench_cases as BC
    from helio_b200.distributed import Shard, gather_meshes

    world, rank, local_rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    pages_all, lods_all, masks_all = BC.planet_page_set(H, int(os.environ.get("FOCI", "256")))
    costs = np.array([H.chunk_cost(32, int(m)) for m in masks_all], dtype=np.uint64)
    owner = H.partition_chunks(costs, world)
    mine = np.flatnonzero(owner == rank)
    n = len(mine)
    batch = H.ChunkBatchExtractor(local_rank, edge=32, max_chunks=n, max_vertices=4608, max_indices=6912)
    batch.fill_density(int(H.ExtractionFixtureKind.Plane), np.ascontiguousarray(pages_all[mine]), np.ascontiguousarray(lods_all[mine]))
    batch.extract_regular(None, n, transition_mask=[int(m) for m in masks_all[mine]])
    v, i, ranges = batch.ctx.read_meshes(0, 0, n)
    batch.close()
    v_t = torch.from_numpy(v.view(np.int32).copy()).to(device)
    i_t = torch.from_numpy(i.view(np.int32).copy()).to(device)
    shard = Shard(rank, world, mine, owner)
    times = []
    for it in range(6):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        out = gather_meshes(v_t, i_t, ranges, shard, dst=0)
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        times.append(time.perf_counter() - t0)
    if rank == 0:
        gv, gi, gr = out
        moved = (gv.numel() - v_t.numel() + gi.numel() - i_t.numel()) * 4
        t = float(np.median(times[1:]))
        print(json.dumps({"case": "planet_set_final_gather_to_rank0", "n_gpus": world, "pages": int(len(pages_all)),
                          "vertices": int(gv.numel() // 8), "indices": int(gi.numel()), "bytes_received_over_nvlink": int(moved),
                          "ms_median": t * 1e3, "GBps_received": moved / t / 1e9 if moved else None,
                          "note": "wall time of gather_meshes: size exchange, grouped isend/irecv, host placement, one interleave launch per arena"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
