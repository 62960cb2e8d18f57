"""Thin object wrapper over one ``hvx_ctx`` (device + stream + output arenas)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi
from .errors import raise_for_status

VERTEX_DTYPE = np.dtype([("position", "<f4", 3), ("material", "<u4"), ("normal", "<f4", 3), ("flags", "<u4")])
EMISSION_COUNTERS_DTYPE = np.dtype([(n, "<u4") for n in (
    "required_vertices", "required_indices", "emitted_vertices", "emitted_indices", "vertex_overflow",
    "index_overflow", "completed", "_pad")])
CLASSIFY_COUNTERS_DTYPE = np.dtype([(n, "<u4") for n in ("visited_cells", "active_cells", "vertices", "triangles")])
TRANSITION_COUNTERS_DTYPE = np.dtype([(n, "<u4") for n in (
    "active_cells", "active_faces", "required_vertices", "required_indices", "emitted_vertices", "emitted_indices",
    "vertex_overflow", "index_overflow", "completed", "_pad0", "_pad1", "_pad2")])
CELL_RECORD_DTYPE = np.dtype([("packed_case_class_counts", "<u4"), ("generation_low", "<u4"),
                              ("generation_high", "<u4"), ("_pad", "<u4")])
CELL_OFFSET_DTYPE = np.dtype([("first_vertex", "<u4"), ("first_index", "<u4"), ("generation_low", "<u4"),
                              ("generation_high", "<u4")])
SCAN_BLOCK_DTYPE = np.dtype([("vertex_count", "<u4"), ("index_count", "<u4"), ("first_vertex", "<u4"),
                             ("first_index", "<u4")])
RANGE_DTYPE = np.dtype([("first_vertex", "<u4"), ("vertex_count", "<u4"), ("first_index", "<u4"),
                        ("index_count", "<u4")])
MESHLET_DTYPE = np.dtype([(n, "<u4") for n in ("first_index", "index_count", "first_vertex", "vertex_count",
                                                 "bounds_offset", "generation_low", "generation_high", "_pad")])
MESHLET_BOUNDS_DTYPE = np.dtype([("center", "<f4", 3), ("radius", "<f4"), ("cone_apex", "<f4", 3),
                                 ("cone_cutoff", "<f4"), ("cone_axis", "<f4", 3), ("_pad", "<f4")])
GATHER_COUNTERS_DTYPE = np.dtype([(n, "<u4") for n in ("regular_samples", "transition_samples", "table_probes", "page_misses",
                                                        "stale_targets", "completed")] + [("_pad", "<u4", 2)])
TERRAIN_MESHLET_BUILD_INDICES = 63  # PV/src/terrain_meshlet.rs:7-8
assert VERTEX_DTYPE.itemsize == 32 and EMISSION_COUNTERS_DTYPE.itemsize == 32 and MESHLET_BOUNDS_DTYPE.itemsize == 48
assert TRANSITION_COUNTERS_DTYPE.itemsize == 48 and CLASSIFY_COUNTERS_DTYPE.itemsize == 16

_BUF_DTYPES = {
    _ffi.BUF_SAMPLES: np.dtype("<u4"), _ffi.BUF_SLABS: np.dtype("<u4"),
    _ffi.BUF_REGULAR_VERTICES: VERTEX_DTYPE, _ffi.BUF_REGULAR_INDICES: np.dtype("<u4"),
    _ffi.BUF_REGULAR_COUNTERS: EMISSION_COUNTERS_DTYPE, _ffi.BUF_REGULAR_CLASSIFY: CLASSIFY_COUNTERS_DTYPE,
    _ffi.BUF_REGULAR_RANGES: RANGE_DTYPE, _ffi.BUF_REGULAR_CELLS: CELL_RECORD_DTYPE,
    _ffi.BUF_REGULAR_OFFSETS: CELL_OFFSET_DTYPE, _ffi.BUF_REGULAR_BLOCKS: SCAN_BLOCK_DTYPE,
    _ffi.BUF_TRANSITION_VERTICES: VERTEX_DTYPE, _ffi.BUF_TRANSITION_INDICES: np.dtype("<u4"),
    _ffi.BUF_TRANSITION_COUNTERS: TRANSITION_COUNTERS_DTYPE, _ffi.BUF_TRANSITION_RANGES: RANGE_DTYPE,
    _ffi.BUF_TRANSITION_CELLS: CELL_RECORD_DTYPE, _ffi.BUF_TRANSITION_OFFSETS: CELL_OFFSET_DTYPE,
    _ffi.BUF_TRANSITION_BLOCKS: SCAN_BLOCK_DTYPE,
    _ffi.BUF_REGULAR_MESHLETS: MESHLET_DTYPE, _ffi.BUF_REGULAR_MESHLET_BOUNDS: MESHLET_BOUNDS_DTYPE,
    _ffi.BUF_REGULAR_MESHLET_COUNTS: np.dtype("<u4"), _ffi.BUF_TRANSITION_MESHLETS: MESHLET_DTYPE,
    _ffi.BUF_TRANSITION_MESHLET_BOUNDS: MESHLET_BOUNDS_DTYPE, _ffi.BUF_TRANSITION_MESHLET_COUNTS: np.dtype("<u4"),
    _ffi.BUF_GATHER_COUNTERS: GATHER_COUNTERS_DTYPE, _ffi.BUF_GATHER_INDIRECT: np.dtype("<u4"),
}


def _input_pointer(array, expected_dtype=np.uint32):
    """(pointer, element_count, keepalive) for a numpy array, a torch tensor or None."""
    if array is None:
        return None, None, None
    if hasattr(array, "data_ptr"):  # torch tensor (host or device) -- no torch import needed
        if not array.is_contiguous():
            raise ValueError("tensor must be contiguous")
        if array.element_size() != 4:
            raise ValueError("tensor must hold 32-bit CellWords")
        return C.c_void_p(array.data_ptr()), array.numel(), array
    arr = np.ascontiguousarray(array, dtype=expected_dtype)
    return C.c_void_p(arr.ctypes.data), arr.size, arr


_DESC_DTYPE = np.dtype([("generation", "<u8"), ("dirty_microbricks", "<u8"), ("transition_mask", "<u4"), ("cost_hint", "<u4"),
                        ("flags", "<u4"), ("_reserved", "<u4")])


def make_descs(n, generation=1, dirty_microbricks=(1 << 64) - 1, transition_mask=0, cost_hint=0, flags=0):
    """ctypes array of ``hvx_chunk_desc``; scalar arguments broadcast, sequences are per chunk.  ``cost_hint``
    (e.g. each chunk's vertex count last time) makes the batch start its heaviest chunks first; ``flags``:
    ``HVX_CHUNK_UNIFORM`` marks chunks the producer knows to hold no surface (not uploaded, not read)."""
    arr = np.zeros(max(n, 1), dtype=_DESC_DTYPE)

    def column(value, mask):
        if isinstance(value, np.ndarray) and value.dtype.kind in "ui":   # fast path: no per-element Python
            return value[:n].astype(np.uint64) & np.uint64(mask)
        if hasattr(value, "__len__"):
            return np.array([int(v) & mask for v in value[:n]], dtype=np.uint64)
        return np.uint64(int(value) & mask)

    arr["generation"][:n] = column(generation, (1 << 64) - 1)
    arr["dirty_microbricks"][:n] = column(dirty_microbricks, (1 << 64) - 1)
    arr["transition_mask"][:n] = column(transition_mask, 0xFFFFFFFF)
    arr["cost_hint"][:n] = column(cost_hint, 0xFFFFFFFF)
    arr["flags"][:n] = column(flags, 0xFFFFFFFF)
    descs = (_ffi.ChunkDesc * max(n, 1)).from_buffer(arr)
    descs._keepalive = arr
    return descs


class Context:
    """One device, one stream, one set of fixed-stride output arenas (``hvx_ctx``)."""

    def __init__(self, device=0, *, edge=32, max_chunks=1, max_vertices=393_216, max_indices=491_520,
                 max_transition_vertices=0, max_transition_indices=0, debug_records=False, first_generation=False,
                 error_kind="regular"):
        self._lib = _ffi.load()
        self._handle = C.c_void_p()
        self._error_kind = error_kind
        cfg = _ffi.Config(edge, max_chunks, max_vertices, max_indices, max_transition_vertices,
                          max_transition_indices, (_ffi.HVX_CFG_DEBUG_RECORDS if debug_records else 0) |
                          (_ffi.HVX_CFG_FIRST_GENERATION if first_generation else 0), 0)
        status = self._lib.hvx_create(C.byref(self._handle), int(device), C.byref(cfg))
        if status != _ffi.HVX_OK:
            message = self._lib.hvx_last_error(None).decode()
            self._handle = C.c_void_p()
            raise_for_status(status, message, error_kind, max_vertices=max_vertices, max_indices=max_indices)
        self.device = int(device)
        self.edge = edge
        self.max_chunks = max_chunks
        self.max_vertices = max_vertices
        self.max_indices = max_indices
        self.max_transition_vertices = max_transition_vertices
        self.max_transition_indices = max_transition_indices
        self.debug_records = debug_records
        self.first_generation = first_generation
        self.sample_words = (edge + 2) ** 3
        self.slab_words = 18 * (2 * edge + 3) ** 2
        self.cells = edge ** 3
        self.transition_cells = 6 * edge * edge

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_handle", None) and self._handle.value:
            self._lib.hvx_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, status, kind=None, **extra):
        if status != _ffi.HVX_OK:
            raise_for_status(status, self._lib.hvx_last_error(self._handle).decode(), kind or self._error_kind, **extra)

    # -- plumbing ---------------------------------------------------------------------------
    def set_stream(self, cuda_stream):
        """Run on a caller-owned cudaStream_t (integer handle, e.g. torch.cuda.current_stream().cuda_stream)."""
        self._check(self._lib.hvx_set_stream(self._handle, C.c_void_p(cuda_stream or 0)))

    def synchronize(self):
        self._check(self._lib.hvx_synchronize(self._handle))

    def regular_kernel_name(self, partial=False):
        return self._lib.hvx_regular_kernel_name(self._handle, int(bool(partial))).decode()

    def debug_set_mode(self, mode):
        """Roofline probes of the regular kernel: 0 normal, 1 stream only, 2 stream + sign bits (no meshes)."""
        self._check(self._lib.hvx_debug_set_mode(self._handle, int(mode)))

    @property
    def launch_count(self):
        return int(self._lib.hvx_launch_count(self._handle))

    @property
    def allocated_bytes(self):
        return int(self._lib.hvx_allocated_bytes(self._handle))

    def buffer_ptr(self, buffer_id):
        ptr = self._lib.hvx_buffer(self._handle, buffer_id)
        if not ptr:
            raise ValueError(f"buffer {buffer_id} is not available in this configuration")
        return int(ptr)

    def buffer_bytes(self, buffer_id):
        return int(self._lib.hvx_buffer_bytes(self._handle, buffer_id))

    def read(self, buffer_id, first=0, count=None):
        """Synchronising readback of ``count`` elements starting at element ``first``."""
        dtype = _BUF_DTYPES[buffer_id]
        total = self.buffer_bytes(buffer_id) // dtype.itemsize
        if count is None:
            count = total - first
        out = np.empty(count, dtype=dtype)
        self._check(self._lib.hvx_read(self._handle, buffer_id, first * dtype.itemsize, count * dtype.itemsize,
                                       C.c_void_p(out.ctypes.data)))
        return out

    def write(self, buffer_id, array, first=0):
        dtype = _BUF_DTYPES[buffer_id]
        arr = np.ascontiguousarray(array, dtype=dtype)
        self._check(self._lib.hvx_write(self._handle, buffer_id, first * dtype.itemsize, arr.nbytes,
                                        C.c_void_p(arr.ctypes.data)))

    # -- K1 -----------------------------------------------------------------------------------
    def _pages(self, page_xyz, lod):
        pages = np.ascontiguousarray(page_xyz, dtype=np.int64).reshape(-1, 3)
        n = pages.shape[0]
        lods = None
        if lod is not None:
            lods = np.ascontiguousarray(np.broadcast_to(np.asarray(lod, dtype=np.uint8), (n,)))
        return pages, lods, n

    def fill_density(self, kind, page_xyz, lod=None, out_ptr=None):
        pages, lods, n = self._pages(page_xyz, lod)
        self._check(self._lib.hvx_fill_density(
            self._handle, int(kind), pages.ctypes.data_as(C.POINTER(C.c_int64)),
            None if lods is None else lods.ctypes.data_as(C.POINTER(C.c_uint8)), n, C.c_void_p(out_ptr or 0)))
        return n

    def fill_slabs(self, kind, page_xyz, lod, out_ptr=None):
        pages, lods, n = self._pages(page_xyz, lod)
        self._check(self._lib.hvx_fill_slabs(
            self._handle, int(kind), pages.ctypes.data_as(C.POINTER(C.c_int64)),
            None if lods is None else lods.ctypes.data_as(C.POINTER(C.c_uint8)), n, C.c_void_p(out_ptr or 0)),
            kind="transition")
        return n

    def apply_edit(self, op, center, radius, page_xyz, lod=None, material=1, samples_ptr=None):
        """hvx_apply_edit: one AddSphere (op 1) / SubtractSphere (op 2) edit on the resident samples of the chunks at
        ``page_xyz``.  Returns (dirty_microbricks [n] uint64, number of chunks whose samples changed)."""
        pages, lods, n = self._pages(page_xyz, lod)
        edit = _ffi.VoxelEdit(0, int(op), int(material), (C.c_float * 3)(*[float(c) for c in center]), float(radius), 0)
        dirty = np.zeros(n, dtype=np.uint64)
        touched = C.c_uint32()
        self._check(self._lib.hvx_apply_edit(
            self._handle, C.byref(edit), pages.ctypes.data_as(C.POINTER(C.c_int64)),
            None if lods is None else lods.ctypes.data_as(C.POINTER(C.c_uint8)), n, C.c_void_p(samples_ptr or 0),
            dirty.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(touched)))
        return dirty, touched.value

    # -- K2-K4 --------------------------------------------------------------------------------
    def extract_regular(self, samples, descs, n, sample_words=None, classify_only=False):
        ptr, count, keep = _input_pointer(samples)
        if sample_words is None:
            sample_words = n * self.sample_words if count is None else count
        fn = self._lib.hvx_classify_regular if classify_only else self._lib.hvx_extract_regular
        self._check(fn(self._handle, ptr, int(sample_words), descs, n), kind="regular")
        del keep

    def extract_regular_to_host(self, samples, descs, n, vertices_out=None, indices_out=None, vertex_cap=None, index_cap=None):
        """hvx_extract_regular_to_host: pipelined upload -> extraction -> packed read-back in one call.

        ``vertices_out`` / ``indices_out``: numpy arrays or pinned torch tensors (32-byte vertices / u32 indices);
        if omitted they are allocated with ``vertex_cap`` / ``index_cap`` elements (default: the slot capacities).
        Returns (vertices, indices, ranges, counters) -- with caller-owned tensors the first two are the totals."""
        ptr, count, keep = _input_pointer(samples)
        sample_words = n * self.sample_words if count is None else count
        own = vertices_out is None or indices_out is None
        if own:
            vertices_out = np.empty(vertex_cap if vertex_cap is not None else n * self.max_vertices, dtype=VERTEX_DTYPE)
            indices_out = np.empty(index_cap if index_cap is not None else n * self.max_indices, dtype=np.uint32)
        torchy = hasattr(vertices_out, "data_ptr")
        vptr = vertices_out.data_ptr() if torchy else vertices_out.ctypes.data
        iptr = indices_out.data_ptr() if torchy else indices_out.ctypes.data
        vcap = vertices_out.numel() * vertices_out.element_size() // 32 if torchy else vertices_out.size
        icap = indices_out.numel() if torchy else indices_out.size
        ranges = np.zeros(n, dtype=RANGE_DTYPE)
        counters = np.zeros(n, dtype=EMISSION_COUNTERS_DTYPE)
        tv, ti = C.c_uint64(), C.c_uint64()
        self._check(self._lib.hvx_extract_regular_to_host(
            self._handle, ptr, int(sample_words), descs, n, C.c_void_p(vptr), vcap, C.c_void_p(iptr), icap,
            ranges.ctypes.data_as(C.POINTER(_ffi.Range)), C.c_void_p(counters.ctypes.data), C.byref(tv), C.byref(ti)),
            kind="regular")
        del keep
        if torchy:
            return tv.value, ti.value, ranges, counters
        return vertices_out[:tv.value], indices_out[:ti.value], ranges, counters

    def extract_transition(self, slabs, descs, n, slab_words=None):
        ptr, count, keep = _input_pointer(slabs)
        if slab_words is None:
            slab_words = n * self.slab_words if count is None else count
        self._check(self._lib.hvx_extract_transition(self._handle, ptr, int(slab_words), descs, n), kind="transition")
        del keep

    def build_meshlets(self, n, kind=0):
        """63-index meshlets + bounds for chunks [0, n) of the last extraction (PV/src/terrain_meshlet.rs)."""
        self._check(self._lib.hvx_build_meshlets(self._handle, kind, n), kind="transition" if kind else "regular")

    def weld_meshes(self, n, kind=0):
        """Optional vertex-reuse output: merge the bit-identical vertex records (the copies of one cell edge) of chunks
        [0, n) of the last extraction in place; indices follow, ranges / emitted_vertices shrink (hvx_weld_meshes)."""
        self._check(self._lib.hvx_weld_meshes(self._handle, kind, n), kind="transition" if kind else "regular")

    def gather_surface(self, residency, table, atlas, jobs):
        """hvx_gather_surface: halo blocks + transition slabs of ``jobs`` from the page atlas into the ctx arenas.

        residency / table / jobs: numpy structured arrays (helio_b200.gather dtypes); atlas: numpy uint32
        array or a torch tensor (host or device) of 32^3-word tiles in linear order."""
        residency = np.ascontiguousarray(residency)
        table = None if table is None else np.ascontiguousarray(table)     # None: the table bound with bind_page_table
        jobs = np.ascontiguousarray(jobs)
        aptr, awords, keep = _input_pointer(atlas)
        self._check(self._lib.hvx_gather_surface(self._handle, C.c_void_p(residency.ctypes.data),
                                                 C.c_void_p(table.ctypes.data if table is not None else 0),
                                                 aptr, awords, C.c_void_p(jobs.ctypes.data), len(jobs)))
        del keep

    def bind_page_table(self, table):
        """hvx_gather_bind_table: keep the page table resident on the device (``None`` unbinds)."""
        if table is None:
            self._check(self._lib.hvx_gather_bind_table(self._handle, None, 0))
            return
        table = np.ascontiguousarray(table)
        self._check(self._lib.hvx_gather_bind_table(self._handle, C.c_void_p(table.ctypes.data), len(table)))

    def read_meshlets(self, chunk, kind=0):
        """Host copy of one chunk's (descriptors, bounds) after build_meshlets."""
        base = _ffi.BUF_TRANSITION_MESHLETS if kind else _ffi.BUF_REGULAR_MESHLETS
        max_indices = self.max_transition_indices if kind else self.max_indices
        stride = (max_indices + 62) // 63
        count = int(self.read(base + 2, chunk, 1)[0])
        return self.read(base, chunk * stride, count), self.read(base + 1, chunk * stride, count)

    def read_meshes(self, kind=0, first=0, n=None, vertices_out=None, indices_out=None):
        """Packed host copy of chunks [first, first+n): (vertices, indices, ranges)."""
        if n is None:
            n = self.max_chunks - first
        ranges = np.zeros(n, dtype=RANGE_DTYPE)
        tv, ti = C.c_uint64(), C.c_uint64()
        rp = ranges.ctypes.data_as(C.POINTER(_ffi.Range))
        if vertices_out is None or indices_out is None:
            status = self._lib.hvx_read_meshes(self._handle, kind, first, n, None, 0, None, 0, rp, C.byref(tv), C.byref(ti))
            if status not in (_ffi.HVX_OK, _ffi.HVX_E_INVALID_CAPACITY):
                self._check(status)
            vertices_out = np.empty(tv.value, dtype=VERTEX_DTYPE)
            indices_out = np.empty(ti.value, dtype=np.uint32)
            if tv.value == 0 and ti.value == 0:
                return vertices_out, indices_out, ranges
        vptr = vertices_out.data_ptr() if hasattr(vertices_out, "data_ptr") else vertices_out.ctypes.data
        iptr = indices_out.data_ptr() if hasattr(indices_out, "data_ptr") else indices_out.ctypes.data
        vcap = vertices_out.numel() * vertices_out.element_size() // 32 if hasattr(vertices_out, "data_ptr") else vertices_out.size
        icap = indices_out.numel() if hasattr(indices_out, "data_ptr") else indices_out.size
        self._check(self._lib.hvx_read_meshes(self._handle, kind, first, n, C.c_void_p(vptr), vcap, C.c_void_p(iptr), icap,
                                              rp, C.byref(tv), C.byref(ti)))
        if hasattr(vertices_out, "data_ptr"):
            return tv.value, ti.value, ranges
        return vertices_out[:tv.value], indices_out[:ti.value], ranges
