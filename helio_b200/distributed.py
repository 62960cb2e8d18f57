"""Multi-GPU chunk scheduling: one process per GPU, static LPT partition, optional mesh gather.

Chunks are independent (every chunk's samples carry their own halo and transition slabs,
PV/src/fixture.rs:83-86), so the data path has NO collective: each rank fills and extracts its own
shard into its own arenas.  ``gather_meshes`` is the optional final step for callers that want
every mesh on one rank; it is an all-gather-v built from ``torch.distributed`` point-to-point /
all_gather calls (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .lod import chunk_cost, partition_chunks


@dataclass
class Shard:
    """The chunks one rank owns, in global chunk order."""
    rank: int
    world_size: int
    global_index: np.ndarray   # indices into the global chunk list, ascending
    owner: np.ndarray          # owner rank of every global chunk


def shard_chunks(n_chunks, world_size, rank, *, costs=None, transition_masks=None, edge=64) -> Shard:
    """LOD-aware static partition: cost = bytes the chunk moves (samples + slabs of its masked
    faces), heaviest first to the lightest rank.  Deterministic, identical on every rank."""
    if costs is None:
        masks = np.zeros(n_chunks, dtype=np.uint32) if transition_masks is None else np.asarray(transition_masks)
        base = {m: chunk_cost(edge, int(m)) for m in np.unique(masks)}
        costs = np.array([base[int(m)] for m in masks], dtype=np.uint64)
    owner = partition_chunks(costs, world_size)
    return Shard(rank, world_size, np.flatnonzero(owner == rank), owner)


def gather_meshes(vertices, indices, ranges, shard: Shard, dst=0, group=None):
    """Optional final gather.  Every rank passes its packed local meshes (``vertices`` as a uint8/
    structured tensor viewable as int32, ``indices`` int32/uint32, ``ranges`` = per-local-chunk
    (first_vertex, vertex_count, first_index, index_count)).  Rank ``dst`` receives
    (vertices, indices, ranges) for ALL chunks in global chunk order; other ranks get None.

    Index values stay chunk-local, so concatenation is exact: the result is byte-identical to what
    one GPU produces for the whole list (tests/test_multi_gpu.py).
    """
    import torch
    import torch.distributed as dist

    world, rank = shard.world_size, shard.rank
    v = torch.as_tensor(vertices).contiguous().view(torch.int32).reshape(-1)
    i = torch.as_tensor(indices).contiguous().view(torch.int32).reshape(-1)
    r = torch.as_tensor(np.ascontiguousarray(ranges).view(np.uint32).astype(np.int64)).reshape(-1, 4)
    device = v.device
    if world == 1:
        return v, i, r.numpy()
    # 1. everyone learns everyone's sizes (tiny all_gather)
    sizes = torch.tensor([v.numel(), i.numel(), r.shape[0]], dtype=torch.int64, device=device)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = torch.stack(all_sizes).cpu().numpy()
    # 2. variable-size gather to dst: grouped point-to-point (NCCL batches these over NVLink)
    recv_v, recv_i, recv_r = {}, {}, {}
    ops = []
    r_dev = r.to(device)
    if rank == dst:
        for src in range(world):
            if src == dst:
                recv_v[src], recv_i[src], recv_r[src] = v, i, r_dev
                continue
            nv, ni, nr = (int(x) for x in all_sizes[src])
            recv_v[src] = torch.empty(nv, dtype=torch.int32, device=device)
            recv_i[src] = torch.empty(ni, dtype=torch.int32, device=device)
            recv_r[src] = torch.empty((nr, 4), dtype=torch.int64, device=device)
            for buf in (recv_v[src], recv_i[src], recv_r[src]):
                if buf.numel():
                    ops.append(dist.P2POp(dist.irecv, buf, src, group=group))
    else:
        for buf in (v, i, r_dev):
            if buf.numel():
                ops.append(dist.P2POp(dist.isend, buf, dst, group=group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    if rank != dst:
        return None
    # 3. interleave back into global chunk order: every rank's ranges come to the host in ONE copy, the placement is
    #    computed there, and each arena is rearranged by ONE device launch (hvx_copy_segments) -- no per-chunk sync
    n_global = shard.owner.size
    all_r = torch.cat([recv_r[src] for src in range(world)]).cpu().numpy()        # [sum nr, 4]
    row0 = np.concatenate([[0], np.cumsum(all_sizes[:, 2])])[:world]              # first range row of each rank
    v0 = np.concatenate([[0], np.cumsum(all_sizes[:, 0])])[:world]                # first word of each rank in the concatenation
    i0 = np.concatenate([[0], np.cumsum(all_sizes[:, 1])])[:world]
    owner = shard.owner.astype(np.int64)
    local = np.zeros(n_global, dtype=np.int64)                                   # position of chunk g in its owner's list
    for src in range(world):
        mask = owner == src
        local[mask] = np.arange(int(mask.sum()))
    rows = all_r[row0[owner] + local]
    nv, ni = rows[:, 1], rows[:, 3]
    out_ranges = np.zeros((n_global, 4), dtype=np.int64)
    out_ranges[:, 1], out_ranges[:, 3] = nv, ni
    out_ranges[:, 0] = np.cumsum(nv) - nv
    out_ranges[:, 2] = np.cumsum(ni) - ni
    cat_v = torch.cat([recv_v[src] for src in range(world)])
    cat_i = torch.cat([recv_i[src] for src in range(world)])
    vertices_out, indices_out = torch.empty_like(cat_v), torch.empty_like(cat_i)
    seg_v = np.stack([v0[owner] + rows[:, 0] * 8, out_ranges[:, 0] * 8, nv * 8], axis=1).astype(np.uint64)
    seg_i = np.stack([i0[owner] + rows[:, 2], out_ranges[:, 2], ni], axis=1).astype(np.uint64)
    if device.type == "cuda":
        import ctypes as C
        from . import _ffi
        lib = _ffi.load()
        stream = torch.cuda.current_stream(device).cuda_stream
        for seg, src_t, dst_t in ((seg_v, cat_v, vertices_out), (seg_i, cat_i, indices_out)):
            seg = np.ascontiguousarray(seg[seg[:, 2] != 0])
            if len(seg):
                status = lib.hvx_copy_segments(device.index or 0, C.c_void_p(stream), C.c_void_p(src_t.data_ptr()), C.c_void_p(dst_t.data_ptr()),
                                               seg.ctypes.data_as(C.POINTER(C.c_uint64)), len(seg))
                if status != _ffi.HVX_OK:
                    raise RuntimeError(f"hvx_copy_segments failed: {lib.hvx_last_error(None).decode()}")
    else:   # gloo / CPU tensors: the host-logic tests; plain slice copies
        sv, dv, si, di = cat_v.numpy(), vertices_out.numpy(), cat_i.numpy(), indices_out.numpy()
        for (a0, b0, n0), (a1, b1, n1) in zip(seg_v.astype(np.int64), seg_i.astype(np.int64)):
            dv[b0:b0 + n0] = sv[a0:a0 + n0]
            di[b1:b1 + n1] = si[a1:a1 + n1]
    return vertices_out, indices_out, out_ranges
