"""CPU restatement of hvx_apply_edit (TEST INFRASTRUCTURE: only tests/ and bench.py's CPU legs import this).

Sphere edits on resident samples -- VoxelOp::AddSphere / SubtractSphere (crates/helio-voxel-core/src/edit.rs:5-10,
GpuVoxelEdit op_type 1 / 2, gpu_types.rs:47-54) -- and the dirty-microbrick set by the legacy octree's
sphere-vs-box rule (crates/helio-voxel-core/src/octree.rs:139-173: a box is untouched when
|centre_a - box_centre_a| > half_a + r on any axis).

parity unpinned: the reference queues edits (crates/helio/src/scene/voxel.rs:81-115) but ships no code that applies
them to planetary pages, so the density rule is this repo's (include/hvx.h, hvx_apply_edit); the dirty rule is the
reference's, applied per microbrick with r = radius + 2 cells.  All arithmetic in float32, operation by operation
as the CUDA kernel / host code does it.
"""
import numpy as np

F = np.float32


def _overlaps(center, lo, hi, r):
    c = (lo + hi) * F(0.5)
    h = (hi - lo) * F(0.5)
    return bool(np.all(np.abs(center - c) <= h + r))


def apply_edit(samples, page_xyz, lod, edge, op, center, radius, material=1):
    """Returns (edited copy of samples [n*(edge+2)^3] uint32, dirty [n] uint64, touched chunk indices)."""
    S = edge + 2
    words = S ** 3
    pages = np.asarray(page_xyz, dtype=np.int64).reshape(-1, 3)
    n = len(pages)
    lods = np.zeros(n, dtype=np.int64) if lod is None else np.broadcast_to(np.asarray(lod, dtype=np.int64), (n,))
    out = np.array(samples, dtype=np.uint32, copy=True)
    dirty = np.zeros(n, dtype=np.uint64)
    touched = []
    center = np.asarray(center, dtype=F)
    radius = F(radius)
    Q = edge // 4
    for i in range(n):
        cell_m = F(0.1) * F(1 << int(lods[i]))
        lo = (pages[i] * edge - 1).astype(F) * cell_m
        hi = (pages[i] * edge + edge).astype(F) * cell_m
        if not _overlaps(center, lo, hi, radius):
            continue
        touched.append(i)
        reach = radius + F(2.0) * cell_m
        bits = 0
        for mz in range(4):
            for my in range(4):
                for mx in range(4):
                    m = np.array([mx, my, mz], dtype=np.int64)
                    blo = (pages[i] * edge + m * Q).astype(F) * cell_m
                    bhi = (pages[i] * edge + (m + 1) * Q).astype(F) * cell_m
                    if _overlaps(center, blo, bhi, reach):
                        bits |= 1 << (mx + 4 * my + 16 * mz)
        dirty[i] = bits
        scale = 1 << int(lods[i])
        idx = np.arange(S, dtype=np.int64)
        ax = [((pages[i][a] * edge - 1 + idx) * scale).astype(F) * F(0.1) - center[a] for a in range(3)]
        dz, dy, dx = np.meshgrid(ax[2], ax[1], ax[0], indexing="ij")
        dist = np.sqrt((dx * dx + dy * dy) + dz * dz, dtype=F)
        inside = (dist <= radius) & (np.abs(dz) <= radius)
        q = np.rint(((radius - dist) / cell_m) * F(256.0))
        carve = np.where(q > F(32767.0), 32767, np.where(inside, q, 0).astype(np.int64)).astype(np.int64)
        block = out[i * words:(i + 1) * words].reshape(S, S, S)
        old = (block & 0xFFFF).astype(np.uint16).view(np.int16).astype(np.int64)
        mat = ((block >> 16) & 0xFF).astype(np.int64)
        if op == 2:
            d = np.maximum(old, carve)
            mat2 = np.where(d > 0, 0, mat)
        elif op == 1:
            d = np.minimum(old, -carve)
            mat2 = np.where((d <= 0) & (old > 0), material & 0xFF, mat)
        else:
            raise ValueError("op must be 1 (AddSphere) or 2 (SubtractSphere)")
        new = (block & np.uint32(0xFF000000)) | (mat2.astype(np.uint32) << np.uint32(16)) | (d.astype(np.int16).view(np.uint16).astype(np.uint32))
        block[...] = np.where(inside, new, block)
    return out, dirty, np.array(touched, dtype=np.uint32)
