#!/usr/bin/env python3
"""One launch (or a few) of every kernel OTHER than the regular extractor, at its benchmark size, for
`ncu --set full` (tools/gpu_calls/gpu_call4.sh); profiles/r02_*_ncu.txt are cut from that capture."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import helio_b200 as H  # noqa: E402
import bench_cases as BC  # noqa: E402

torch.cuda.set_device(0)
device = torch.device("cuda", 0)


def grid(n_axis, ys):
    xs = np.arange(-n_axis // 2, n_axis // 2)
    z, y, x = np.meshgrid(xs, np.array(ys), xs, indexing="ij")
    return np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1).astype(np.int64)


# K1 fill: the headline batch (terrain, 4096 x 66^3) and the integer fields at both edges
pages = grid(16, range(-8, 8))
b = H.ChunkBatchExtractor(0, edge=64, max_chunks=len(pages), max_vertices=8, max_indices=8)
for kind in (16, 0, 17):
    b.fill_density(kind, pages)
b.ctx.synchronize()
b.close()
b = H.ChunkBatchExtractor(0, edge=32, max_chunks=len(pages), max_vertices=8, max_indices=8)
for kind in (16, 0):
    b.fill_density(kind, pages)
b.ctx.synchronize()
b.close()

# transition extraction + slab fill: 1024 coarse 64^3 pages, all six faces
pages = grid(32, [-1])
n = len(pages)
b = H.ChunkBatchExtractor(0, edge=64, max_chunks=n, max_vertices=49152, max_indices=73728, max_transition_vertices=16384,
                          max_transition_indices=49152, debug_records=False)
lods = np.ones(n, dtype=np.uint8)
b.fill_density(16, pages, lods)
b.fill_slabs(16, pages, lods)
d = H.make_descs(n, transition_mask=0x3F)
b.ctx.extract_regular(None, d, n)
b.ctx.extract_transition(None, d, n)
b.ctx.build_meshlets(n, 0)
b.ctx.build_meshlets(n, 1)
# sphere edit + dirty re-extraction, and the record kernels
dirty, touched = b.ctx.apply_edit(2, (3.0, -3.0, 3.0), 1.5, pages, lods)
b.ctx.synchronize()
b.close()
r = H.ChunkBatchExtractor(0, edge=32, max_chunks=512, max_vertices=12288, max_indices=18432, debug_records=True)
r.fill_density(16, grid(16, [-2, -1]))
r.ctx.extract_regular(None, H.make_descs(512), 512)
r.ctx.synchronize()
# optional vertex-reuse pass over 4096 terrain pages
w = H.ChunkBatchExtractor(0, edge=32, max_chunks=4096, max_vertices=12288, max_indices=18432)
w.fill_density(16, grid(16, range(-8, 8)))
w.ctx.extract_regular(None, H.make_descs(4096), 4096)
w.ctx.weld_meshes(4096)
w.ctx.synchronize()
w.close()
r.close()

# gather -> extract -> meshlets -> publish from a resident atlas (one pass)
stream = torch.cuda.Stream(device=device)
orig_timed = BC.timed
BC.timed = lambda torch_, stream_, fn, warmup=0, iters=1: (fn(), stream_.synchronize(), {"ms_median": 1.0, "ms_p95": 1.0, "ms_min": 1.0})[2]
BC.page_pass_case(H, torch, device, 0, stream=stream)
BC.timed = orig_timed

# bounded extraction publisher commit
b = H.ChunkBatchExtractor(0, edge=32, max_chunks=256, max_vertices=12288, max_indices=18432)
b.fill_density(16, grid(16, [-1]))
b.extract_regular(None, 256)
ec = b.counters(256)
pub = H.BoundedExtractionPublisher(H.ExtractionLimits.new(256, 256, 2_000_000, 3_000_000, 60_000))
pub.attach(b.ctx)
res = []
for k in range(256):
    ni = int(ec["emitted_indices"][k])
    counts = H.SurfaceCounts(int(ec["emitted_vertices"][k]), ni, H.max_meshlets_for_indices(ni) if ni else 0)
    res.append(pub.reserve(H.PlanetPageKey.new(bytes(16), H.PageKey(0, (k, 0, 0))), 1, counts).reservation)
pub.commit(range(256), range(256), res)
b.ctx.synchronize()
pub.close()
b.close()

# legacy 8^3-brick marching cubes: 65,536 bricks of a rolling height field, device-resident lists
gx, gy, gz = 64, 16, 64
X, Z = np.meshgrid(np.arange(gx * 8 + 1), np.arange(gz * 8 + 1), indexing="xy")
height = (64 + 30 * np.sin(X * 0.021) * np.cos(Z * 0.017) + 9 * np.sin(X * 0.13 + Z * 0.11)).astype(np.int32)
vol = (np.arange(gy * 8 + 1)[None, :, None] < height[:, None, :]).astype(np.uint8) * 3
sz, sy, sx = vol.strides
view = np.lib.stride_tricks.as_strided(vol, (gz, gy, gx, 9, 9, 9), (8 * sz, 8 * sy, 8 * sx, sz, sy, sx))
nb = gx * gy * gz
padded = np.zeros((nb, 732), dtype=np.uint8)
padded[:, :729] = view.reshape(nb, 729)
words = torch.from_numpy(padded.view("<u4").reshape(-1).astype(np.int32)).cuda()
ex = H.VoxelMeshExtractor(0, max_bricks=nb, max_dirty=nb)
meta = np.zeros(nb, dtype=H.BRICK_META_DTYPE)
meta["data_offset"] = np.arange(nb, dtype=np.uint32) * 183
ex.write_brick_meta(meta)
ex.write_voxel_data(words)
dirtyb = np.zeros(nb, dtype=H.DIRTY_BRICK_DTYPE)
dirtyb["brick_slot"] = np.arange(nb)
bz, by, bx = np.unravel_index(np.arange(nb), (gz, gy, gx))
dirtyb["origin_size"] = np.stack([bx * 8, by * 8, bz * 8, np.ones(nb)], axis=1).astype(np.float32)
d_dirty = torch.from_numpy(dirtyb.view(np.uint8).reshape(-1).copy()).cuda()
d_meta = torch.from_numpy(meta.view(np.uint8).reshape(-1).copy()).cuda()
ex.extract(d_dirty, d_meta)
ex.ctx.synchronize()
ex.close()
print("profile workload done")
