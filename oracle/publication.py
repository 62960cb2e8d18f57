"""CPU restatement of the surface publication step.  TEST INFRASTRUCTURE.

Thin ctypes driver over hvxo_publish_surface / hvxo_refresh_visibility (oracle/hvx_oracle.c, which follow
PV/src/surface_publish.wgsl:104-225), holding the same render state the reference's pass owns."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import oracle as O

JOB_DTYPE = np.dtype([(n, "<u4") for n in ("slot", "transition_mask", "generation_low", "generation_high", "regular_max_vertices",
                                           "regular_max_indices", "transition_max_vertices", "transition_max_indices",
                                           "regular_max_meshlets", "transition_max_meshlets")] + [("_pad", "<u4", 2)])
PAGE_META_DTYPE = np.dtype([("relative_lod0_cell_min", "<i4", 3)] + [(n, "<u4") for n in ("lod", "slot", "generation_low",
                                                                                            "generation_high", "transition_mask")])
STATE_DTYPE = np.dtype([(n, "<u4") for n in ("generation_low", "generation_high", "active_bank", "valid", "regular_vertex_count",
                                             "regular_index_count", "transition_vertex_count", "transition_index_count",
                                             "regular_meshlet_count", "transition_meshlet_count")] + [("_pad", "<u4", 2)])
FEEDBACK_DTYPE = np.dtype([(n, "<u4") for n in ("submitted_jobs", "published_jobs", "stale_rejections", "overflow_rejections",
                                                "incomplete_rejections")] + [("_pad", "<u4", 3)])
DRAW_DTYPE = np.dtype([("index_count", "<u4"), ("instance_count", "<u4"), ("first_index", "<u4"), ("base_vertex", "<i4"),
                       ("first_instance", "<u4")])
assert JOB_DTYPE.itemsize == 48 and PAGE_META_DTYPE.itemsize == 32 and STATE_DTYPE.itemsize == 48
assert FEEDBACK_DTYPE.itemsize == 32 and DRAW_DTYPE.itemsize == 20


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Publisher:
    """Render state of `slots` residency slots: banked arenas, states, draws, feedback."""

    def __init__(self, slots, max_vertices, max_indices, max_tvertices, max_tindices):
        self.slots = slots
        self.caps = (max_vertices, max_indices, max_tvertices, max_tindices)
        self.vertices = np.zeros(2 * slots * max_vertices, dtype=O.VERTEX_DTYPE)
        self.indices = np.zeros(2 * slots * max_indices, dtype=np.uint32)
        self.tvertices = np.zeros(2 * slots * max_tvertices, dtype=O.VERTEX_DTYPE)
        self.tindices = np.zeros(2 * slots * max_tindices, dtype=np.uint32)
        self.states = np.zeros(slots, dtype=STATE_DTYPE)
        self.regular_draws = np.zeros(slots, dtype=DRAW_DTYPE)
        self.transition_draws = np.zeros(slots, dtype=DRAW_DTYPE)
        self.feedback = np.zeros(1, dtype=FEEDBACK_DTYPE)

    def job(self, slot, generation, transition_mask=0):
        j = np.zeros(1, dtype=JOB_DTYPE)
        mv, mi, tv, ti = self.caps
        j[0] = (slot, transition_mask, generation & 0xFFFFFFFF, generation >> 32, mv, mi, tv, ti, (mi + 62) // 63, (ti + 62) // 63, (0, 0))
        return j

    def publish(self, job, page_metadata, regular_counters, transition_counters, src_v=None, src_i=None, src_tv=None, src_ti=None):
        L = O.lib()
        L.hvxo_publish_surface.restype = None
        L.hvxo_publish_surface.argtypes = [C.c_void_p] * 16
        rc = np.ascontiguousarray(regular_counters).view(np.uint32).reshape(-1)[:8].copy()
        tc = np.ascontiguousarray(transition_counters).view(np.uint32).reshape(-1)[:12].copy()
        keep = [np.ascontiguousarray(a) if a is not None else None for a in (src_v, src_i, src_tv, src_ti)]
        meta = np.ascontiguousarray(page_metadata, dtype=PAGE_META_DTYPE)
        job = np.ascontiguousarray(job, dtype=JOB_DTYPE)
        L.hvxo_publish_surface(_p(job), _p(meta), _p(rc), _p(tc), _p(keep[0]), _p(keep[1]), _p(keep[2]), _p(keep[3]),
                               _p(self.states), _p(self.vertices) if keep[0] is not None else None,
                               _p(self.indices) if keep[1] is not None else None,
                               _p(self.tvertices) if keep[2] is not None else None,
                               _p(self.tindices) if keep[3] is not None else None,
                               _p(self.regular_draws), _p(self.transition_draws), _p(self.feedback))

    def refresh_visibility(self, visible):
        L = O.lib()
        L.hvxo_refresh_visibility.restype = None
        L.hvxo_refresh_visibility.argtypes = [C.c_uint32] + [C.c_void_p] * 4
        vis = np.ascontiguousarray(visible, dtype=np.uint32)
        L.hvxo_refresh_visibility(self.slots, _p(self.states), _p(vis), _p(self.regular_draws), _p(self.transition_draws))
