"""ctypes binding of libhelio_voxel_cuda.so (include/hvx.h).

There is no fallback of any kind: if the shared library is missing or the machine has no
CUDA device, loading / context creation raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libhelio_voxel_cuda.so"

HVX_OK = 0
HVX_E_SAMPLE_COUNT = -1
HVX_E_INVALID_CAPACITY = -2
HVX_E_DEVICE_LIMIT = -3
HVX_E_TRANSITION_MASK = -4
HVX_E_INVALID_ARGUMENT = -5
HVX_E_CUDA = -6
HVX_E_BATCH_CAPACITY = -7
HVX_E_FINEST_LOD = -8
HVX_E_ADDRESS = -9
HVX_E_TOPOLOGY_EMPTY = -20
HVX_E_TOPOLOGY_DUPLICATE = -21
HVX_E_TOPOLOGY_OVERLAP = -22
HVX_E_TOPOLOGY_UNBALANCED = -23
HVX_E_TOPOLOGY_ROOT_LOD = -24
HVX_E_TOPOLOGY_MINIMUM_LOD = -25
HVX_E_TOPOLOGY_PAGE_BUDGET = -26
HVX_E_TOPOLOGY_MISSING_PARENT = -27
HVX_E_TOPOLOGY_COVERAGE = -28

HVX_E_INVALID_LIMITS = -40
HVX_E_ARITHMETIC_OVERFLOW = -41
HVX_E_NON_TRIANGLE_INDEX_COUNT = -42
HVX_E_INCOMPLETE_SURFACE_COUNTS = -43
HVX_E_PENDING_CAPACITY = -44
HVX_E_ARENA_CAPACITY = -45
HVX_E_GENERATION_CONFLICT = -46
HVX_E_RESERVATION_MISSING = -47
HVX_E_RESERVATION_MISMATCH = -48
HVX_E_DEVICE_BUFFER_LIMIT = -49

HVX_CFG_DEBUG_RECORDS = 1
HVX_CFG_FIRST_GENERATION = 2
HVX_CHUNK_UNIFORM = 1
(XPUB_VERTICES, XPUB_INDICES, XPUB_PAGE_RANGES, XPUB_COUNTERS) = range(4)
(BRICK_VERTICES, BRICK_NORMALS, BRICK_INDICES, BRICK_DESCRIPTORS, BRICK_DRAWS, BRICK_REJECTED) = range(6)
(PUB_REGULAR_VERTICES, PUB_REGULAR_INDICES, PUB_TRANSITION_VERTICES, PUB_TRANSITION_INDICES, PUB_STATES, PUB_REGULAR_DRAWS,
 PUB_TRANSITION_DRAWS, PUB_FEEDBACK) = range(8)

(BUF_SAMPLES, BUF_SLABS, BUF_REGULAR_VERTICES, BUF_REGULAR_INDICES, BUF_REGULAR_COUNTERS, BUF_REGULAR_CLASSIFY,
 BUF_REGULAR_RANGES, BUF_REGULAR_CELLS, BUF_REGULAR_OFFSETS, BUF_REGULAR_BLOCKS, BUF_TRANSITION_VERTICES,
 BUF_TRANSITION_INDICES, BUF_TRANSITION_COUNTERS, BUF_TRANSITION_RANGES, BUF_TRANSITION_CELLS,
 BUF_TRANSITION_OFFSETS, BUF_TRANSITION_BLOCKS, BUF_REGULAR_MESHLETS, BUF_REGULAR_MESHLET_BOUNDS,
 BUF_REGULAR_MESHLET_COUNTS, BUF_TRANSITION_MESHLETS, BUF_TRANSITION_MESHLET_BOUNDS,
 BUF_TRANSITION_MESHLET_COUNTS, BUF_GATHER_COUNTERS, BUF_GATHER_INDIRECT) = range(25)


class Config(C.Structure):
    _fields_ = [("edge", C.c_uint32), ("max_chunks", C.c_uint32), ("max_vertices", C.c_uint32),
                ("max_indices", C.c_uint32), ("max_transition_vertices", C.c_uint32),
                ("max_transition_indices", C.c_uint32), ("flags", C.c_uint32), ("_reserved", C.c_uint32)]


class ChunkDesc(C.Structure):
    _fields_ = [("generation", C.c_uint64), ("dirty_microbricks", C.c_uint64), ("transition_mask", C.c_uint32),
                ("cost_hint", C.c_uint32), ("flags", C.c_uint32), ("_reserved", C.c_uint32)]


class VoxelEdit(C.Structure):
    """GpuVoxelEdit (crates/helio-voxel-core/src/gpu_types.rs:47-54)."""
    _fields_ = [("volume_id", C.c_uint32), ("op_type", C.c_uint32), ("material", C.c_uint32), ("center", C.c_float * 3),
                ("radius", C.c_float), ("_pad", C.c_uint32)]


class EmissionCounters(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("required_vertices", "required_indices", "emitted_vertices", "emitted_indices",
                                          "vertex_overflow", "index_overflow", "completed", "_pad")]


class Range(C.Structure):
    _fields_ = [("first_vertex", C.c_uint32), ("vertex_count", C.c_uint32), ("first_index", C.c_uint32),
                ("index_count", C.c_uint32)]


class Page(C.Structure):
    _fields_ = [("page_xyz", C.c_int64 * 3), ("lod", C.c_uint8), ("transition_mask", C.c_uint8),
                ("_pad", C.c_uint8 * 6)]


class LodStats(C.Structure):
    _fields_ = [("pages", C.c_uint32), ("minimum_lod", C.c_uint32), ("maximum_lod", C.c_uint32),
                ("transition_faces", C.c_uint32)]


# ---- bounded extraction publisher (PV/src/extraction.rs) ------------------------------------------------
def _u32s(*names):
    return [(n, C.c_uint32) for n in names]


class ExtractionRequest(C.Structure):
    _fields_ = _u32s("page_slot", "generation_low", "generation_high", "transition_mask", "dirty_microbricks_low",
                     "dirty_microbricks_high") + [("_pad", C.c_uint32 * 2)]


class ExtractionRange(C.Structure):
    _fields_ = _u32s("first_vertex", "vertex_count", "first_index", "index_count", "first_meshlet", "meshlet_count",
                     "generation_low", "generation_high")


class ExtractionLimits(C.Structure):
    _fields_ = _u32s("max_page_slots", "max_pending_pages", "max_vertices", "max_indices", "max_meshlets")


class ExtractionPlan(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("request_bytes", "page_range_bytes", "vertex_bytes", "index_bytes",
                                          "meshlet_bytes", "counter_bytes", "total_bytes")]


class PlanetPageKey(C.Structure):
    _fields_ = [("planet_id", C.c_uint8 * 16), ("page_xyz", C.c_int64 * 3), ("lod", C.c_uint8), ("_pad", C.c_uint8 * 7)]


class SurfaceCounts(C.Structure):
    _fields_ = _u32s("vertices", "indices", "meshlets")


class ArenaSlice(C.Structure):
    _fields_ = _u32s("first", "count")


class SurfaceAllocation(C.Structure):
    _fields_ = [("vertices", ArenaSlice), ("indices", ArenaSlice), ("meshlets", ArenaSlice)]


class Reservation(C.Structure):
    _fields_ = [("key", PlanetPageKey), ("generation", C.c_uint64), ("allocation", SurfaceAllocation)]


class PublishedSurface(C.Structure):
    _fields_ = [("generation", C.c_uint64), ("allocation", SurfaceAllocation)]


class ReservationOutcome(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("detail", C.c_uint32), ("reservation", Reservation), ("current", PublishedSurface),
                ("newest_generation", C.c_uint64)]


class PublicationOutcome(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("has_replaced", C.c_uint32), ("current", PublishedSurface),
                ("replaced", PublishedSurface), ("newest_generation", C.c_uint64)]


class EvictOutcome(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("_pad", C.c_uint32), ("newest_generation", C.c_uint64)]


class ExtractionPublisherCounters(C.Structure):
    _fields_ = [("current_pages", C.c_uint64), ("pending_pages", C.c_uint64), ("used_vertices", C.c_uint32),
                ("used_indices", C.c_uint32), ("used_meshlets", C.c_uint32), ("_pad0", C.c_uint32),
                ("pending_high_water", C.c_uint64), ("vertex_high_water", C.c_uint32), ("index_high_water", C.c_uint32),
                ("meshlet_high_water", C.c_uint32), ("_pad1", C.c_uint32)] + \
               [(n, C.c_uint64) for n in ("reservations", "publications", "replacements", "cancellations", "evictions",
                                          "stale_rejected", "backpressured")]


# every symbol include/hvx.h declares; tests assert the library exports all of them
EXPORTS = [
    "hvx_create", "hvx_destroy", "hvx_last_error", "hvx_status_name", "hvx_abi_version", "hvx_get_config",
    "hvx_allocated_bytes", "hvx_set_stream", "hvx_get_stream", "hvx_synchronize", "hvx_launch_count", "hvx_debug_set_mode", "hvx_start_order", "hvx_regular_kernel_name", "hvx_selftest_edge_parameter", "hvx_selftest_inv_sqrt",
    "hvx_fill_density", "hvx_fill_slabs", "hvx_apply_edit", "hvx_extract_regular", "hvx_extract_regular_to_host", "hvx_classify_regular", "hvx_extract_transition",
    "hvx_build_meshlets", "hvx_weld_meshes", "hvx_gather_surface", "hvx_gather_bind_table", "hvx_publisher_create", "hvx_publisher_destroy", "hvx_publish_surfaces",
    "hvx_refresh_visibility", "hvx_publisher_buffer", "hvx_publisher_buffer_bytes", "hvx_publisher_read", "hvx_publisher_write",
    "hvx_buffer", "hvx_buffer_bytes", "hvx_read", "hvx_write", "hvx_read_meshes", "hvx_lod_topology",
    "hvx_horizon_plan", "hvx_partition_chunks", "hvx_chunk_cost", "hvx_copy_segments",
    "hvx_extraction_request_new", "hvx_extraction_limits_plan", "hvx_extraction_limits_validate_device",
    "hvx_extraction_gpu_range", "hvx_extraction_publisher_create", "hvx_extraction_publisher_destroy",
    "hvx_extraction_reserve", "hvx_extraction_publish", "hvx_extraction_cancel_pending", "hvx_extraction_evict",
    "hvx_extraction_current", "hvx_extraction_pending", "hvx_extraction_publisher_get_counters",
    "hvx_extraction_publisher_attach", "hvx_extraction_commit", "hvx_extraction_publisher_buffer",
    "hvx_extraction_publisher_read",
    "hvx_brick_mesher_create", "hvx_brick_mesher_destroy", "hvx_brick_extract", "hvx_brick_clear_slot", "hvx_brick_buffer",
    "hvx_brick_buffer_bytes", "hvx_brick_read",
]

_lib = None


def load() -> C.CDLL:
    """Load the CUDA extension; raises if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("HVX_LIBRARY", LIB_PATH))
    if not path.exists():
        raise RuntimeError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). helio_b200 has no CPU or PyTorch fallback.")
    L = C.CDLL(str(path))
    vp, u32p, u64p, i64p, u8p = C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_int64), \
        C.POINTER(C.c_uint8)
    L.hvx_create.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(Config)]
    L.hvx_destroy.argtypes = [vp]
    L.hvx_destroy.restype = None
    L.hvx_last_error.argtypes = [vp]
    L.hvx_last_error.restype = C.c_char_p
    L.hvx_status_name.argtypes = [C.c_int]
    L.hvx_status_name.restype = C.c_char_p
    L.hvx_abi_version.restype = C.c_uint32
    L.hvx_get_config.argtypes = [vp, C.POINTER(Config)]
    L.hvx_allocated_bytes.argtypes = [vp]
    L.hvx_allocated_bytes.restype = C.c_uint64
    L.hvx_set_stream.argtypes = [vp, vp]
    L.hvx_get_stream.argtypes = [vp]
    L.hvx_get_stream.restype = vp
    L.hvx_synchronize.argtypes = [vp]
    L.hvx_launch_count.argtypes = [vp]
    L.hvx_launch_count.restype = C.c_uint64
    L.hvx_debug_set_mode.argtypes = [vp, C.c_uint32]
    L.hvx_start_order.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    L.hvx_regular_kernel_name.argtypes = [vp, C.c_int]
    L.hvx_regular_kernel_name.restype = C.c_char_p
    L.hvx_selftest_edge_parameter.argtypes = [C.c_int, u64p, u32p]
    L.hvx_selftest_inv_sqrt.argtypes = [C.c_int, u64p, u32p]
    L.hvx_fill_density.argtypes = [vp, C.c_uint32, i64p, u8p, C.c_uint32, vp]
    L.hvx_fill_slabs.argtypes = [vp, C.c_uint32, i64p, u8p, C.c_uint32, vp]
    L.hvx_apply_edit.argtypes = [vp, C.POINTER(VoxelEdit), i64p, u8p, C.c_uint32, vp, u64p, u32p]
    L.hvx_extract_regular.argtypes = [vp, vp, C.c_uint64, C.POINTER(ChunkDesc), C.c_uint32]
    L.hvx_classify_regular.argtypes = [vp, vp, C.c_uint64, C.POINTER(ChunkDesc), C.c_uint32]
    L.hvx_extract_regular_to_host.argtypes = [vp, vp, C.c_uint64, C.POINTER(ChunkDesc), C.c_uint32, vp, C.c_uint64, vp,
                                              C.c_uint64, C.POINTER(Range), vp, u64p, u64p]
    L.hvx_extract_transition.argtypes = [vp, vp, C.c_uint64, C.POINTER(ChunkDesc), C.c_uint32]
    L.hvx_build_meshlets.argtypes = [vp, C.c_int, C.c_uint32]
    L.hvx_weld_meshes.argtypes = [vp, C.c_int, C.c_uint32]
    L.hvx_gather_surface.argtypes = [vp, vp, vp, vp, C.c_uint64, vp, C.c_uint32]
    L.hvx_gather_bind_table.argtypes = [vp, vp, C.c_uint32]
    L.hvx_publisher_create.argtypes = [vp, C.c_uint32, C.POINTER(vp)]
    L.hvx_publisher_destroy.argtypes = [vp]
    L.hvx_publisher_destroy.restype = None
    L.hvx_publish_surfaces.argtypes = [vp, vp, vp, vp, C.c_uint32]
    L.hvx_refresh_visibility.argtypes = [vp, vp]
    L.hvx_publisher_buffer.argtypes = [vp, C.c_int]
    L.hvx_publisher_buffer.restype = vp
    L.hvx_publisher_buffer_bytes.argtypes = [vp, C.c_int]
    L.hvx_publisher_buffer_bytes.restype = C.c_uint64
    L.hvx_publisher_read.argtypes = [vp, C.c_int, C.c_uint64, C.c_uint64, vp]
    L.hvx_publisher_write.argtypes = [vp, C.c_int, C.c_uint64, C.c_uint64, vp]
    L.hvx_buffer.argtypes = [vp, C.c_int]
    L.hvx_buffer.restype = vp
    L.hvx_buffer_bytes.argtypes = [vp, C.c_int]
    L.hvx_buffer_bytes.restype = C.c_uint64
    L.hvx_read.argtypes = [vp, C.c_int, C.c_uint64, C.c_uint64, vp]
    L.hvx_write.argtypes = [vp, C.c_int, C.c_uint64, C.c_uint64, vp]
    L.hvx_read_meshes.argtypes = [vp, C.c_int, C.c_uint32, C.c_uint32, vp, C.c_uint64, vp, C.c_uint64,
                                  C.POINTER(Range), u64p, u64p]
    L.hvx_lod_topology.argtypes = [C.POINTER(Page), C.c_uint32, C.c_uint32, C.POINTER(LodStats)]
    L.hvx_horizon_plan.argtypes = [i64p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(Page), u32p,
                                   C.POINTER(Page), C.POINTER(LodStats)]
    L.hvx_partition_chunks.argtypes = [u64p, C.c_uint32, C.c_uint32, u32p]
    L.hvx_copy_segments.argtypes = [C.c_int, vp, vp, vp, u64p, C.c_uint32]
    L.hvx_chunk_cost.argtypes = [C.c_uint32, C.c_uint32]
    L.hvx_chunk_cost.restype = C.c_uint64
    L.hvx_extraction_request_new.argtypes = [C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint64, C.POINTER(ExtractionRequest)]
    L.hvx_extraction_limits_plan.argtypes = [C.POINTER(ExtractionLimits), C.POINTER(ExtractionPlan)]
    L.hvx_extraction_limits_validate_device.argtypes = [C.POINTER(ExtractionLimits), C.c_uint64, C.c_uint64,
                                                        C.POINTER(C.c_char_p), u64p]
    L.hvx_extraction_gpu_range.argtypes = [C.POINTER(SurfaceAllocation), C.c_uint64, C.POINTER(ExtractionRange)]
    L.hvx_extraction_gpu_range.restype = None
    L.hvx_extraction_publisher_create.argtypes = [C.POINTER(ExtractionLimits), C.POINTER(vp)]
    L.hvx_extraction_publisher_destroy.argtypes = [vp]
    L.hvx_extraction_publisher_destroy.restype = None
    L.hvx_extraction_reserve.argtypes = [vp, C.POINTER(PlanetPageKey), C.c_uint64, C.POINTER(SurfaceCounts),
                                         C.POINTER(ReservationOutcome)]
    L.hvx_extraction_publish.argtypes = [vp, C.POINTER(Reservation), C.POINTER(PublicationOutcome)]
    L.hvx_extraction_cancel_pending.argtypes = [vp, C.POINTER(PlanetPageKey), C.c_uint64, C.POINTER(C.c_int)]
    L.hvx_extraction_evict.argtypes = [vp, C.POINTER(PlanetPageKey), C.c_uint64, C.POINTER(EvictOutcome)]
    L.hvx_extraction_current.argtypes = [vp, C.POINTER(PlanetPageKey), C.POINTER(PublishedSurface)]
    L.hvx_extraction_pending.argtypes = [vp, C.POINTER(PlanetPageKey), C.POINTER(Reservation)]
    L.hvx_extraction_publisher_get_counters.argtypes = [vp, C.POINTER(ExtractionPublisherCounters)]
    L.hvx_extraction_publisher_attach.argtypes = [vp, vp]
    L.hvx_extraction_commit.argtypes = [vp, u32p, u32p, C.POINTER(Reservation), C.c_uint32]
    L.hvx_extraction_publisher_buffer.argtypes = [vp, C.c_int]
    L.hvx_extraction_publisher_buffer.restype = vp
    L.hvx_extraction_publisher_read.argtypes = [vp, C.c_int, C.c_uint64, C.c_uint64, vp]
    L.hvx_brick_mesher_create.argtypes = [vp, C.c_uint32, C.POINTER(vp)]
    L.hvx_brick_mesher_destroy.argtypes = [vp]
    L.hvx_brick_mesher_destroy.restype = None
    L.hvx_brick_extract.argtypes = [vp, vp, C.c_uint32, vp, C.c_uint64, vp, C.c_uint32]
    L.hvx_brick_clear_slot.argtypes = [vp, C.c_uint32]
    L.hvx_brick_buffer.argtypes = [vp, C.c_int]
    L.hvx_brick_buffer.restype = vp
    L.hvx_brick_buffer_bytes.argtypes = [vp, C.c_int]
    L.hvx_brick_buffer_bytes.restype = C.c_uint64
    L.hvx_brick_read.argtypes = [vp, C.c_int, C.c_uint64, C.c_uint64, vp]
    _lib = L
    return L
