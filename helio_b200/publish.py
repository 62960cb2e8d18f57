"""Host-side mirror of the reference's surface publication state.

Reference: ``GpuSurfaceJob`` / ``GpuSurfaceState`` / ``GpuSurfaceFeedback`` / ``GpuDrawPage`` /
``DrawIndexedIndirectArgs`` (PV/src/render.rs:466-554), ``GpuPageMeta``
(crates/helio-planet-voxel-core/src/gpu.rs:62-99) and the four entry points of
PV/src/surface_publish.wgsl.  The renderer that consumes the draws is out of scope; this module owns
the double-banked arenas and the per-slot state so that extraction results can be published
generation-safely without leaving the device.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi
from .context import VERTEX_DTYPE, Context

SURFACE_JOB_DTYPE = np.dtype([(n, "<u4") for n in (
    "slot", "transition_mask", "generation_low", "generation_high", "regular_max_vertices", "regular_max_indices",
    "transition_max_vertices", "transition_max_indices", "regular_max_meshlets", "transition_max_meshlets")] + [("_pad", "<u4", 2)])
PAGE_META_DTYPE = np.dtype([("relative_lod0_cell_min", "<i4", 3)] + [(n, "<u4") for n in (
    "lod", "slot", "generation_low", "generation_high", "transition_mask")])
SURFACE_STATE_DTYPE = np.dtype([(n, "<u4") for n in (
    "generation_low", "generation_high", "active_bank", "valid", "regular_vertex_count", "regular_index_count",
    "transition_vertex_count", "transition_index_count", "regular_meshlet_count", "transition_meshlet_count")] + [("_pad", "<u4", 2)])
DRAW_PAGE_DTYPE = np.dtype([("relative_lod0_cell_min", "<i4", 3), ("lod", "<u4"), ("camera_relative_m", "<f4", 3),
                            ("lod0_cell_size_m", "<f4"), ("generation_low", "<u4"), ("generation_high", "<u4"),
                            ("transition_mask", "<u4"), ("visible", "<u4")])
SURFACE_FEEDBACK_DTYPE = np.dtype([(n, "<u4") for n in (
    "submitted_jobs", "published_jobs", "stale_rejections", "overflow_rejections", "incomplete_rejections")] + [("_pad", "<u4", 3)])
DRAW_INDEXED_INDIRECT_DTYPE = np.dtype([("index_count", "<u4"), ("instance_count", "<u4"), ("first_index", "<u4"),
                                        ("base_vertex", "<i4"), ("first_instance", "<u4")])
assert SURFACE_JOB_DTYPE.itemsize == 48 and PAGE_META_DTYPE.itemsize == 32 and SURFACE_STATE_DTYPE.itemsize == 48
assert DRAW_PAGE_DTYPE.itemsize == 48 and SURFACE_FEEDBACK_DTYPE.itemsize == 32 and DRAW_INDEXED_INDIRECT_DTYPE.itemsize == 20

_PUB_DTYPES = {
    _ffi.PUB_REGULAR_VERTICES: VERTEX_DTYPE, _ffi.PUB_REGULAR_INDICES: np.dtype("<u4"),
    _ffi.PUB_TRANSITION_VERTICES: VERTEX_DTYPE, _ffi.PUB_TRANSITION_INDICES: np.dtype("<u4"),
    _ffi.PUB_STATES: SURFACE_STATE_DTYPE, _ffi.PUB_REGULAR_DRAWS: DRAW_INDEXED_INDIRECT_DTYPE,
    _ffi.PUB_TRANSITION_DRAWS: DRAW_INDEXED_INDIRECT_DTYPE, _ffi.PUB_FEEDBACK: SURFACE_FEEDBACK_DTYPE,
}


def max_meshlets_for_indices(max_indices: int) -> int:
    return (max_indices + 62) // 63


class SurfacePublisher:
    """Double-banked per-slot arenas + surface states + indirect draws + feedback for ``slots`` residency slots."""

    def __init__(self, context: Context, slots: int):
        self._ctx = context
        self._lib = context._lib
        self._handle = C.c_void_p()
        context._check(self._lib.hvx_publisher_create(context._handle, slots, C.byref(self._handle)))
        self.slots = slots

    def close(self):
        if self._handle:
            self._lib.hvx_publisher_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def surface_job(self, slot: int, generation: int, transition_mask: int = 0) -> np.ndarray:
        """``GpuSurfaceJob::new`` (PV/src/render.rs:482-503) for this publisher's bank capacities."""
        c = self._ctx
        job = np.zeros(1, dtype=SURFACE_JOB_DTYPE)
        job[0] = (slot, transition_mask, generation & 0xFFFFFFFF, generation >> 32, c.max_vertices, c.max_indices,
                  c.max_transition_vertices, c.max_transition_indices, max_meshlets_for_indices(c.max_indices),
                  max_meshlets_for_indices(c.max_transition_indices), (0, 0))
        return job

    def publish(self, jobs, job_chunk, page_metadata) -> None:
        jobs = np.ascontiguousarray(jobs, dtype=SURFACE_JOB_DTYPE)
        chunks = np.ascontiguousarray(job_chunk, dtype=np.uint32)
        meta = np.ascontiguousarray(page_metadata, dtype=PAGE_META_DTYPE)
        if len(meta) != self.slots or len(chunks) != len(jobs):
            raise ValueError("page_metadata needs one entry per slot and job_chunk one per job")
        self._ctx._check(self._lib.hvx_publish_surfaces(self._handle, C.c_void_p(jobs.ctypes.data), C.c_void_p(chunks.ctypes.data),
                                                        C.c_void_p(meta.ctypes.data), len(jobs)))

    def refresh_visibility(self, draw_pages) -> None:
        pages = np.ascontiguousarray(draw_pages, dtype=DRAW_PAGE_DTYPE)
        if len(pages) != self.slots:
            raise ValueError("draw_pages needs one entry per slot")
        self._ctx._check(self._lib.hvx_refresh_visibility(self._handle, C.c_void_p(pages.ctypes.data)))

    def read(self, buffer_id, first=0, count=None) -> np.ndarray:
        dtype = _PUB_DTYPES[buffer_id]
        total = self._lib.hvx_publisher_buffer_bytes(self._handle, buffer_id) // dtype.itemsize
        count = total - first if count is None else count
        out = np.empty(count, dtype=dtype)
        self._ctx._check(self._lib.hvx_publisher_read(self._handle, buffer_id, first * dtype.itemsize, count * dtype.itemsize,
                                                      C.c_void_p(out.ctypes.data)))
        return out

    def write(self, buffer_id, array, first=0) -> None:
        dtype = _PUB_DTYPES[buffer_id]
        arr = np.ascontiguousarray(array, dtype=dtype)
        self._ctx._check(self._lib.hvx_publisher_write(self._handle, buffer_id, first * dtype.itemsize, arr.nbytes,
                                                       C.c_void_p(arr.ctypes.data)))

    def device_pointer(self, buffer_id) -> int:
        return int(self._lib.hvx_publisher_buffer(self._handle, buffer_id) or 0)

    # accessors named after the reference's render-pass buffers
    def surface_states(self):
        return self.read(_ffi.PUB_STATES)

    def regular_draws(self):
        return self.read(_ffi.PUB_REGULAR_DRAWS)

    def transition_draws(self):
        return self.read(_ffi.PUB_TRANSITION_DRAWS)

    def feedback(self):
        return self.read(_ffi.PUB_FEEDBACK)[0]
