"""Small workload for compute-sanitizer (memcheck / racecheck): every kernel, both edges, a few chunks."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import helio_b200 as H  # noqa: E402

for edge in (32, 64):
    pages = [[0, -1, 0], [0, 0, 0], [1, -1, 1], [-1, -1, 0], [0, 1, 0], [2, -1, -2]]
    n = len(pages)
    b = H.ChunkBatchExtractor(0, edge=edge, max_chunks=n, max_vertices=40000, max_indices=60000,
                              max_transition_vertices=8192, max_transition_indices=24576, debug_records=(edge == 32))
    for kind in (16, 1, 17):
        b.fill_density(kind, pages)
        b.extract_regular(None, n, transition_mask=[0, 0x3F, 1, 2, 0, 0x15], dirty_microbricks=[(1 << 64) - 1] * 5 + [0xFF00FF])
        b.ctx.build_meshlets(n, 0)
    b.ctx.weld_meshes(n)          # optional vertex-reuse pass, then meshlets over the shared-vertex mesh
    b.ctx.build_meshlets(n, 0)
    b.fill_slabs(16, pages, [1] * n)
    b.extract_transition(None, n, [0x3F, 0x15, 0, 1, 0x2A, 0x3F])
    b.ctx.build_meshlets(n, 1)
    b.ctx.weld_meshes(n, kind=1)
    c = b.counters(n)
    t = b.transition_counters(n)
    print(edge, int(c["emitted_vertices"].sum()), int(t["emitted_vertices"].sum()))
    b.close()

# no debug records -> the decoupled (default) kernel at edge 32 as well
b = H.ChunkBatchExtractor(0, edge=32, max_chunks=6, max_vertices=40000, max_indices=60000)
b.fill_density(16, [[0, -1, 0], [0, 0, 0], [1, -1, 1], [-1, -1, 0], [0, 1, 0], [2, -1, -2]])
b.extract_regular(None, 6, transition_mask=[0, 0x3F, 1, 2, 0, 0x15])
ec = b.counters(6)

# bounded extraction publisher: commit into the packed arenas
pub = H.BoundedExtractionPublisher(H.ExtractionLimits.new(8, 6, 60000, 90000, 2000))
pub.attach(b.ctx)
res = []
for k in range(6):
    ni = int(ec["emitted_indices"][k])
    counts = H.SurfaceCounts(int(ec["emitted_vertices"][k]), ni, H.max_meshlets_for_indices(ni) if ni else 0)
    res.append(pub.reserve(H.PlanetPageKey.new(bytes(16), H.PageKey(0, (k, 0, 0))), 1, counts).reservation)
pub.commit(range(6), range(6), res)
print("commit", pub.read(3, H.EXTRACTION_COUNTERS_DTYPE)[0])
pub.close()
b.close()

# legacy 8^3-brick marching cubes: noise, checker (overflow), empty, full
rng = np.random.default_rng(5)
z, y, x = np.indices((9, 9, 9))
bricks = [(rng.random((9, 9, 9)) < 0.4) * 7, ((x + y + z) & 1) * 4, np.zeros((9, 9, 9)), np.full((9, 9, 9), 9)]
words = np.concatenate([H.pack_brick(v.astype(np.uint8)) for v in bricks])
ex = H.VoxelMeshExtractor(0, max_bricks=4)
meta = np.zeros(4, dtype=H.BRICK_META_DTYPE)
meta["data_offset"] = np.arange(4) * 183
ex.write_brick_meta(meta)
ex.write_voxel_data(words)
for k in range(4):
    ex.mark_dirty(k, 0, (0.0, 0.0, 0.0), 0.5, True)
ex.execute()
print("bricks", ex.descriptors()["vertex_count"])
ex.close()
print("sanitize workload done")
