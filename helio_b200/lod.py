"""Mixed-LOD page sets -> per-page transition masks, and the bounded horizon page plan.

Mirrors PV/src/lod_topology.rs (``TerrainLodTopology``, ``HorizonLodFixturePlan``); the work is
done by the library's host-side C++ (helio_b200/csrc/lod_topology.cpp), no GPU involved.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

from . import _ffi
from .errors import raise_for_status
from .types import PAGE_EDGE, PageKey


@dataclass(frozen=True)
class TerrainLodTopologyStats:
    """PV/src/lod_topology.rs:7-13."""
    pages: int
    minimum_lod: int
    maximum_lod: int
    transition_faces: int


def _to_pages(keys):
    arr = (_ffi.Page * max(len(keys), 1))()
    for i, key in enumerate(keys):
        arr[i].page_xyz[:] = key.page_xyz
        arr[i].lod = key.lod
    return arr


class TerrainLodTopology:
    """PV/src/lod_topology.rs:20-146: validates a visible page set, derives coarse-owned masks."""

    def __init__(self, pages, edge=PAGE_EDGE):
        keys = [p if isinstance(p, PageKey) else PageKey(*p) for p in pages]
        arr = _to_pages(keys)
        stats = _ffi.LodStats()
        status = _ffi.load().hvx_lod_topology(arr, len(keys), edge, C.byref(stats))
        if status != _ffi.HVX_OK:
            raise_for_status(status, "")
        self._masks = {PageKey(arr[i].lod, tuple(arr[i].page_xyz)): int(arr[i].transition_mask)
                       for i in range(len(keys))}
        self._stats = TerrainLodTopologyStats(stats.pages, stats.minimum_lod, stats.maximum_lod, stats.transition_faces)
        self.edge = edge

    @classmethod
    def _from_plan(cls, masks, stats, edge):
        self = cls.__new__(cls)
        self._masks, self._stats, self.edge = masks, stats, edge
        return self

    def pages(self):
        return sorted(self._masks)

    def transition_mask(self, page):
        return self._masks.get(page)

    def transition_masks(self):
        return dict(sorted(self._masks.items()))

    def stats(self):
        return self._stats

    def __eq__(self, other):
        return isinstance(other, TerrainLodTopology) and self._masks == other._masks


class HorizonLodFixturePlan:
    """PV/src/lod_topology.rs:148-230: bounded mixed-LOD plan around a focus cell."""

    def __init__(self, root, focus, topology):
        self._root, self._focus, self._topology = root, tuple(focus), topology

    @classmethod
    def build(cls, focus_lod0_cell, root_lod, max_pages, edge=PAGE_EDGE):
        return cls.build_with_minimum_lod(focus_lod0_cell, root_lod, 0, max_pages, edge)

    @classmethod
    def build_with_minimum_lod(cls, focus_lod0_cell, root_lod, minimum_lod, max_pages, edge=PAGE_EDGE):
        focus = (C.c_int64 * 3)(*[int(v) for v in focus_lod0_cell])
        out = (_ffi.Page * max(max_pages, 1))()
        n = C.c_uint32()
        root = _ffi.Page()
        stats = _ffi.LodStats()
        status = _ffi.load().hvx_horizon_plan(focus, root_lod, minimum_lod, max_pages, edge, out, C.byref(n),
                                              C.byref(root), C.byref(stats))
        if status != _ffi.HVX_OK:
            raise_for_status(status, "")
        masks = {PageKey(out[i].lod, tuple(out[i].page_xyz)): int(out[i].transition_mask) for i in range(n.value)}
        topo = TerrainLodTopology._from_plan(
            masks, TerrainLodTopologyStats(stats.pages, stats.minimum_lod, stats.maximum_lod, stats.transition_faces), edge)
        return cls(PageKey(root.lod, tuple(root.page_xyz)), focus_lod0_cell, topo)

    def root(self):
        return self._root

    def focus_lod0_cell(self):
        return self._focus

    def topology(self):
        return self._topology

    def __eq__(self, other):
        return (isinstance(other, HorizonLodFixturePlan) and self._root == other._root
                and self._focus == other._focus and self._topology == other._topology)


def chunk_cost(edge, transition_mask=0):
    """Bytes a chunk moves through the extractor (samples + slabs of its masked faces)."""
    return int(_ffi.load().hvx_chunk_cost(edge, transition_mask))


def partition_chunks(costs, ranks):
    """LPT-greedy static partition: returns owner rank per chunk, deterministic."""
    import numpy as np
    costs = np.ascontiguousarray(costs, dtype=np.uint64)
    owner = np.zeros(costs.size, dtype=np.uint32)
    status = _ffi.load().hvx_partition_chunks(costs.ctypes.data_as(C.POINTER(C.c_uint64)), costs.size, ranks,
                                              owner.ctypes.data_as(C.POINTER(C.c_uint32)))
    if status != _ffi.HVX_OK:
        raise_for_status(status, "invalid partition request")
    return owner


def start_order(cost_hints, spread_pct=75):
    """The order in which hvx_extract_regular starts the chunks of a hinted batch (hvx_start_order): descending hint,
    a few heavy chunks among many light ones spread over the first ``spread_pct`` per cent of the order."""
    import numpy as np
    hints = np.ascontiguousarray(cost_hints, dtype=np.uint32)
    order = np.zeros(hints.size, dtype=np.uint32)
    status = _ffi.load().hvx_start_order(hints.ctypes.data, hints.size, int(spread_pct), order.ctypes.data)
    if status != _ffi.HVX_OK:
        raise_for_status(status, "invalid start-order request")
    return order
