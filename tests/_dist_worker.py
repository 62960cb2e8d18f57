"""Worker for tests/test_multi_gpu.py: one process per rank (gloo on CPU, nccl on GPUs)."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
for entry in (str(ROOT), str(ROOT / "tests")):
    if entry not in sys.path:
        sys.path.insert(0, entry)


def chunk_list():
    """A small planet patch: a mixed-LOD horizon plan (pages + transition masks) at edge 32."""
    import helio_b200 as H
    plan = H.HorizonLodFixturePlan.build_with_minimum_lod([40, -1, -17], 4, 0, 64)
    masks = plan.topology().transition_masks()
    return [(k.lod, list(k.page_xyz), m) for k, m in masks.items()]


def main(rank, world, port, backend, out_dir):
    import torch
    import torch.distributed as dist
    import helio_b200 as H
    from helio_b200.distributed import gather_meshes, shard_chunks
    from oracle import oracle as O

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group(backend, rank=rank, world_size=world)
    chunks = chunk_list()
    shard = shard_chunks(len(chunks), world, rank, transition_masks=[m for _, _, m in chunks], edge=32)
    mine = [chunks[g] for g in shard.global_index]
    use_gpu = backend == "nccl"
    if use_gpu:
        torch.cuda.set_device(rank)
        batch = H.ChunkBatchExtractor(rank, edge=32, max_chunks=max(len(mine), 1), max_vertices=16384, max_indices=24576)
        batch.fill_density(int(H.ExtractionFixtureKind.Plane), [p for _, p, _ in mine], [l for l, _, _ in mine])
        batch.extract_regular(None, len(mine), transition_mask=[m for _, _, m in mine])
        verts, idx, ranges = batch.ctx.read_meshes(0, 0, len(mine))
        device = torch.device("cuda", rank)
    else:
        parts_v, parts_i, ranges = [], [], np.zeros(len(mine), dtype=H.RANGE_DTYPE)
        tv = ti = 0
        for n, (lod, page, mask) in enumerate(mine):  # CPU ranks stand in with the oracle: this test is about
            mesh = O.extract_regular(O.fixture_fill(O.FIELD_PLANE, page, lod=lod), transition_mask=mask, debug=False)
            parts_v.append(mesh.vertices)             # the scheduler + gather, not about the kernels
            parts_i.append(mesh.indices)
            ranges[n] = (tv, len(mesh.vertices), ti, len(mesh.indices))
            tv += len(mesh.vertices)
            ti += len(mesh.indices)
        verts = np.concatenate(parts_v) if parts_v else np.zeros(0, dtype=H.VERTEX_DTYPE)
        idx = np.concatenate(parts_i) if parts_i else np.zeros(0, dtype=np.uint32)
        device = torch.device("cpu")
    v_t = torch.from_numpy(verts.view(np.int32).copy()).to(device)
    i_t = torch.from_numpy(idx.view(np.int32).copy()).to(device)
    result = gather_meshes(v_t, i_t, ranges, shard, dst=0)
    if rank == 0:
        gv, gi, gr = result
        np.savez(Path(out_dir) / "gathered.npz", vertices=gv.cpu().numpy(), indices=gi.cpu().numpy(), ranges=gr,
                 owner=shard.owner)
    else:
        assert result is None
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5])
