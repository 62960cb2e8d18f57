"""CPU statement of the optional vertex-reuse output (hvx_weld_meshes) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product path never does.

PARITY UNPINNED: the reference has no vertex reuse (PV/src/transvoxel_emit.wgsl:322-358 ignores the reuse byte of the
vertex code, PV/src/transvoxel.rs:99-101 reuse() has no caller), so there is no golden vector to pin this
to.  The definition is therefore stated on the reference's OWN output, which is pinned: take the unshared mesh
(oracle.extract_regular / extract_transition), merge bit-identical 32-byte vertex records keeping the first occurrence
in the original order, redirect the indices.  A vertex is a pure function of its cell edge (corner samples, their
gradients, the chunk's transition mask), so this is the mesh edge-ownership reuse produces.
"""
from __future__ import annotations

import numpy as np


def weld_mesh(vertices, indices):
    """vertices: (V,) structured array or (V, 8) uint32 view of 32-byte records; indices: (I,) uint32, chunk-local.

    Returns (kept_vertices, new_indices, kept_from): kept_from[k] = original index of kept vertex k."""
    raw = np.ascontiguousarray(vertices).view(np.uint32).reshape(-1, 8)
    indices = np.asarray(indices, dtype=np.uint32)
    if len(raw) == 0:
        return vertices[:0], indices.copy(), np.zeros(0, dtype=np.int64)
    keys = np.ascontiguousarray(raw).view(np.dtype((np.void, 32))).ravel()
    _, first, inverse = np.unique(keys, return_index=True, return_inverse=True)
    inverse = inverse.ravel()
    order = np.argsort(first, kind="stable")            # distinct records by first occurrence
    rank = np.empty(len(order), dtype=np.int64)
    rank[order] = np.arange(len(order))
    kept_from = first[order]
    new_index_of_vertex = rank[inverse]                  # original vertex -> kept slot
    return np.ascontiguousarray(vertices)[kept_from], new_index_of_vertex[indices].astype(np.uint32), kept_from
