"""Bounded, generation-safe publication of extracted surfaces into variable-size arenas.

Mirrors PV/src/extraction.rs (``GpuExtractionRequest``, ``ExtractionLimits``, ``SurfaceCounts``,
``BoundedExtractionPublisher`` and its outcome / error enums).  The bookkeeping is the library's host
C++ (helio_b200/csrc/extraction_publisher.cpp); ``attach`` / ``commit`` add the device step that
moves a reserved page's mesh from its extraction slot into the bounded arenas
(``hvx_extraction_commit``).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _ffi
from .errors import HvxError, TransitionMask
from .types import PageKey

EXTRACTION_REQUEST_DTYPE = np.dtype([(n, "<u4") for n in ("page_slot", "generation_low", "generation_high",
                                                          "transition_mask", "dirty_microbricks_low",
                                                          "dirty_microbricks_high", "_pad0", "_pad1")])
EXTRACTION_RANGE_DTYPE = np.dtype([(n, "<u4") for n in ("first_vertex", "vertex_count", "first_index", "index_count",
                                                        "first_meshlet", "meshlet_count", "generation_low",
                                                        "generation_high")])
EXTRACTION_COUNTERS_DTYPE = np.dtype([(n, "<u4") for n in ("requests", "active_cells", "vertices", "indices", "meshlets",
                                                           "completed", "stale_rejected", "overflowed", "vertex_overflow",
                                                           "index_overflow", "meshlet_overflow", "_pad")])


class ExtractionError(HvxError):
    """PV/src/extraction.rs:672-698; ``kind`` is the reference variant name."""

    def __init__(self, kind, status, **fields):
        super().__init__(f"ExtractionError::{kind} {fields if fields else ''}".strip(), status)
        self.kind = kind
        for name, value in fields.items():
            setattr(self, name, value)

    def __eq__(self, other):
        return isinstance(other, ExtractionError) and self.kind == other.kind and self._fields() == other._fields()

    __hash__ = None

    def _fields(self):
        return {k: v for k, v in vars(self).items() if k not in ("kind", "status")}


_KINDS = {
    _ffi.HVX_E_INVALID_LIMITS: "InvalidLimits", _ffi.HVX_E_ARITHMETIC_OVERFLOW: "ArithmeticOverflow",
    _ffi.HVX_E_NON_TRIANGLE_INDEX_COUNT: "NonTriangleIndexCount",
    _ffi.HVX_E_INCOMPLETE_SURFACE_COUNTS: "IncompleteSurfaceCounts", _ffi.HVX_E_PENDING_CAPACITY: "PendingCapacity",
    _ffi.HVX_E_ARENA_CAPACITY: "ArenaCapacity", _ffi.HVX_E_GENERATION_CONFLICT: "GenerationConflict",
    _ffi.HVX_E_RESERVATION_MISSING: "ReservationMissing", _ffi.HVX_E_RESERVATION_MISMATCH: "ReservationMismatch",
    _ffi.HVX_E_DEVICE_BUFFER_LIMIT: "DeviceBufferLimit",
}
EXTRACTION_CAPACITIES = ("Vertices", "Indices", "Meshlets")   # ExtractionCapacity, PV/src/extraction.rs:665-670


def _raise(status, **fields):
    if status == _ffi.HVX_E_TRANSITION_MASK:
        raise TransitionMask(f"transition mask {fields.get('mask', 0):#010b} uses bits outside the six page faces",
                             status, fields.get("mask"))
    raise ExtractionError(_KINDS.get(status, f"status {status}"), status, **fields)


# ---- PODs and plain records ----------------------------------------------------------------------------
@dataclass(frozen=True)
class GpuExtractionRequest:
    """PV/src/extraction.rs:10-51."""
    page_slot: int
    generation_low: int
    generation_high: int
    transition_mask: int
    dirty_microbricks_low: int
    dirty_microbricks_high: int

    @classmethod
    def new(cls, page_slot, generation, transition_mask, dirty_microbricks):
        out = _ffi.ExtractionRequest()
        status = _ffi.load().hvx_extraction_request_new(page_slot, generation, transition_mask, dirty_microbricks, C.byref(out))
        if status != _ffi.HVX_OK:
            _raise(status, mask=transition_mask)
        return cls(out.page_slot, out.generation_low, out.generation_high, out.transition_mask,
                   out.dirty_microbricks_low, out.dirty_microbricks_high)

    def generation(self):
        return self.generation_low | (self.generation_high << 32)

    def dirty_microbricks(self):
        return self.dirty_microbricks_low | (self.dirty_microbricks_high << 32)


@dataclass(frozen=True)
class ExtractionAllocationPlan:
    request_bytes: int
    page_range_bytes: int
    vertex_bytes: int
    index_bytes: int
    meshlet_bytes: int
    counter_bytes: int
    total_bytes: int


@dataclass(frozen=True)
class ExtractionLimits:
    """PV/src/extraction.rs:111-210."""
    max_page_slots: int = 256
    max_pending_pages: int = 32
    max_vertices: int = 1_048_576
    max_indices: int = 3_145_728
    max_meshlets: int = 32_768

    @classmethod
    def new(cls, max_page_slots, max_pending_pages, max_vertices, max_indices, max_meshlets):
        limits = cls(max_page_slots, max_pending_pages, max_vertices, max_indices, max_meshlets)
        limits.allocation_plan()
        return limits

    def _c(self):
        return _ffi.ExtractionLimits(self.max_page_slots, self.max_pending_pages, self.max_vertices, self.max_indices,
                                     self.max_meshlets)

    def allocation_plan(self):
        plan = _ffi.ExtractionPlan()
        status = _ffi.load().hvx_extraction_limits_plan(C.byref(self._c()), C.byref(plan))
        if status != _ffi.HVX_OK:
            _raise(status)
        return ExtractionAllocationPlan(plan.request_bytes, plan.page_range_bytes, plan.vertex_bytes, plan.index_bytes,
                                        plan.meshlet_bytes, plan.counter_bytes, plan.total_bytes)

    def validate_device(self, max_buffer_size, max_storage_buffer_binding_size):
        """``limits`` are the two wgpu::Limits fields the reference checks."""
        name, requested = C.c_char_p(), C.c_uint64()
        status = _ffi.load().hvx_extraction_limits_validate_device(C.byref(self._c()), max_buffer_size,
                                                                   max_storage_buffer_binding_size, C.byref(name),
                                                                   C.byref(requested))
        if status == _ffi.HVX_E_DEVICE_BUFFER_LIMIT:
            _raise(status, name=name.value.decode(), requested=requested.value, max_buffer_bytes=max_buffer_size,
                   max_storage_bytes=max_storage_buffer_binding_size)
        if status != _ffi.HVX_OK:
            _raise(status)


@dataclass(frozen=True)
class SurfaceCounts:
    vertices: int = 0
    indices: int = 0
    meshlets: int = 0


@dataclass(frozen=True)
class ArenaSlice:
    first: int = 0
    count: int = 0


@dataclass(frozen=True)
class SurfaceAllocation:
    vertices: ArenaSlice = ArenaSlice()
    indices: ArenaSlice = ArenaSlice()
    meshlets: ArenaSlice = ArenaSlice()

    def counts(self):
        return SurfaceCounts(self.vertices.count, self.indices.count, self.meshlets.count)

    def gpu_range(self, generation):
        out = _ffi.ExtractionRange()
        _ffi.load().hvx_extraction_gpu_range(C.byref(_allocation_to_c(self)), generation, C.byref(out))
        return np.frombuffer(bytes(out), dtype=EXTRACTION_RANGE_DTYPE)[0]


@dataclass(frozen=True)
class PlanetPageKey:
    """helio-planet-voxel-core/src/types.rs:312-316."""
    planet: bytes
    page: PageKey

    @classmethod
    def new(cls, planet, page):
        return cls(bytes(planet), page)


@dataclass(frozen=True)
class ExtractionReservation:
    key: PlanetPageKey
    generation: int
    allocation: SurfaceAllocation


@dataclass(frozen=True)
class PublishedSurface:
    generation: int
    allocation: SurfaceAllocation


@dataclass(frozen=True)
class ReservationOutcome:
    """kind in {"Reserved", "Current", "DuplicatePending", "Stale"} (PV/src/extraction.rs:296-302)."""
    kind: str
    reservation: Optional[ExtractionReservation] = None
    current: Optional[PublishedSurface] = None
    newest_generation: Optional[int] = None


@dataclass(frozen=True)
class PublicationOutcome:
    """kind in {"Published", "Stale"} (PV/src/extraction.rs:304-313)."""
    kind: str
    current: Optional[PublishedSurface] = None
    replaced: Optional[PublishedSurface] = None
    newest_generation: Optional[int] = None


@dataclass(frozen=True)
class ExtractionEvictOutcome:
    """kind in {"Evicted", "Missing", "Stale"} (PV/src/extraction.rs:315-320)."""
    kind: str
    newest_generation: Optional[int] = None


@dataclass(frozen=True)
class ExtractionPublisherCounters:
    current_pages: int = 0
    pending_pages: int = 0
    used_vertices: int = 0
    used_indices: int = 0
    used_meshlets: int = 0
    pending_high_water: int = 0
    vertex_high_water: int = 0
    index_high_water: int = 0
    meshlet_high_water: int = 0
    reservations: int = 0
    publications: int = 0
    replacements: int = 0
    cancellations: int = 0
    evictions: int = 0
    stale_rejected: int = 0
    backpressured: int = 0


# ---- ctypes <-> records ------------------------------------------------------------------------------
def _key_to_c(key: PlanetPageKey):
    out = _ffi.PlanetPageKey()
    planet = bytes(key.planet).ljust(16, b"\0")[:16]
    out.planet_id[:] = list(planet)
    out.page_xyz[:] = key.page.page_xyz
    out.lod = key.page.lod
    return out


def _key_from_c(c):
    return PlanetPageKey(bytes(c.planet_id), PageKey(c.lod, tuple(c.page_xyz)))


def _slice_from_c(c):
    return ArenaSlice(c.first, c.count)


def _allocation_from_c(c):
    return SurfaceAllocation(_slice_from_c(c.vertices), _slice_from_c(c.indices), _slice_from_c(c.meshlets))


def _allocation_to_c(a: SurfaceAllocation):
    out = _ffi.SurfaceAllocation()
    for name in ("vertices", "indices", "meshlets"):
        getattr(out, name).first = getattr(a, name).first
        getattr(out, name).count = getattr(a, name).count
    return out


def _reservation_from_c(c):
    return ExtractionReservation(_key_from_c(c.key), c.generation, _allocation_from_c(c.allocation))


def _reservation_to_c(r: ExtractionReservation):
    out = _ffi.Reservation()
    out.key = _key_to_c(r.key)
    out.generation = r.generation
    out.allocation = _allocation_to_c(r.allocation)
    return out


def _published_from_c(c):
    return PublishedSurface(c.generation, _allocation_from_c(c.allocation))


class BoundedExtractionPublisher:
    """PV/src/extraction.rs:342-603.  A replacement reserves new ranges without freeing the current
    surface; publication swaps generations atomically and only then recycles the old ranges."""

    def __init__(self, limits: ExtractionLimits):
        self._lib = _ffi.load()
        handle = C.c_void_p()
        status = self._lib.hvx_extraction_publisher_create(C.byref(limits._c()), C.byref(handle))
        if status != _ffi.HVX_OK:
            _raise(status)
        self._handle, self._limits, self._ctx = handle, limits, None

    def close(self):
        if self._handle:
            self._lib.hvx_extraction_publisher_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def limits(self):
        return self._limits

    def current(self, key):
        out = _ffi.PublishedSurface()
        found = self._lib.hvx_extraction_current(self._handle, C.byref(_key_to_c(key)), C.byref(out))
        return _published_from_c(out) if found == 1 else None

    def pending(self, key):
        out = _ffi.Reservation()
        found = self._lib.hvx_extraction_pending(self._handle, C.byref(_key_to_c(key)), C.byref(out))
        return _reservation_from_c(out) if found == 1 else None

    def counters(self):
        out = _ffi.ExtractionPublisherCounters()
        self._lib.hvx_extraction_publisher_get_counters(self._handle, C.byref(out))
        return ExtractionPublisherCounters(**{name: getattr(out, name) for name in ExtractionPublisherCounters.__dataclass_fields__})

    def reserve(self, key, generation, counts: SurfaceCounts):
        out = _ffi.ReservationOutcome()
        status = self._lib.hvx_extraction_reserve(self._handle, C.byref(_key_to_c(key)), generation,
                                                  C.byref(_ffi.SurfaceCounts(counts.vertices, counts.indices, counts.meshlets)),
                                                  C.byref(out))
        if status == _ffi.HVX_E_NON_TRIANGLE_INDEX_COUNT:
            _raise(status, indices=counts.indices)
        if status == _ffi.HVX_E_PENDING_CAPACITY:
            _raise(status, maximum=out.detail)
        if status == _ffi.HVX_E_ARENA_CAPACITY:
            _raise(status, capacity=EXTRACTION_CAPACITIES[out.detail])
        if status == _ffi.HVX_E_GENERATION_CONFLICT:
            _raise(status, key=key, generation=generation)
        if status != _ffi.HVX_OK:
            _raise(status)
        if out.kind == 0:
            return ReservationOutcome("Reserved", reservation=_reservation_from_c(out.reservation))
        if out.kind == 1:
            return ReservationOutcome("Current", current=_published_from_c(out.current))
        if out.kind == 2:
            return ReservationOutcome("DuplicatePending", reservation=_reservation_from_c(out.reservation))
        return ReservationOutcome("Stale", newest_generation=out.newest_generation)

    def publish(self, reservation: ExtractionReservation):
        out = _ffi.PublicationOutcome()
        status = self._lib.hvx_extraction_publish(self._handle, C.byref(_reservation_to_c(reservation)), C.byref(out))
        if status == _ffi.HVX_E_RESERVATION_MISSING:
            _raise(status, key=reservation.key)
        if status == _ffi.HVX_E_RESERVATION_MISMATCH:
            _raise(status, key=reservation.key, generation=reservation.generation)
        if status != _ffi.HVX_OK:
            _raise(status)
        if out.kind == 1:
            return PublicationOutcome("Stale", newest_generation=out.newest_generation)
        return PublicationOutcome("Published", current=_published_from_c(out.current),
                                  replaced=_published_from_c(out.replaced) if out.has_replaced else None)

    def cancel_pending(self, key, generation):
        cancelled = C.c_int()
        status = self._lib.hvx_extraction_cancel_pending(self._handle, C.byref(_key_to_c(key)), generation, C.byref(cancelled))
        if status == _ffi.HVX_E_RESERVATION_MISMATCH:
            _raise(status, key=key, generation=generation)
        if status != _ffi.HVX_OK:
            _raise(status)
        return bool(cancelled.value)

    def evict(self, key, generation):
        out = _ffi.EvictOutcome()
        self._lib.hvx_extraction_evict(self._handle, C.byref(_key_to_c(key)), generation, C.byref(out))
        if out.kind == 0:
            return ExtractionEvictOutcome("Evicted")
        if out.kind == 1:
            return ExtractionEvictOutcome("Missing")
        return ExtractionEvictOutcome("Stale", newest_generation=out.newest_generation)

    # ---- device side ----------------------------------------------------------------------------------
    def attach(self, ctx):
        """Allocate the plan's bounded arenas on ``ctx``'s device (vertices, indices, page ranges, counters)."""
        status = self._lib.hvx_extraction_publisher_attach(self._handle, ctx._handle)
        if status != _ffi.HVX_OK:
            ctx._check(status)
        self._ctx = ctx

    def commit(self, chunks, page_slots, reservations):
        """Copy the regular meshes of extraction chunks ``chunks`` to their reservations' ranges and stamp
        ``page_ranges[page_slot]``; call ``publish`` for each reservation afterwards."""
        n = len(reservations)
        chunk = np.ascontiguousarray(chunks, dtype=np.uint32)
        slot = np.ascontiguousarray(page_slots, dtype=np.uint32)
        arr = (_ffi.Reservation * max(n, 1))(*[_reservation_to_c(r) for r in reservations])
        status = self._lib.hvx_extraction_commit(self._handle, chunk.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                 slot.ctypes.data_as(C.POINTER(C.c_uint32)), arr, n)
        if status != _ffi.HVX_OK:
            self._ctx._check(status)

    def read(self, buffer_id, dtype, first=0, count=None):
        dtype = np.dtype(dtype)
        total = {_ffi.XPUB_VERTICES: self._limits.max_vertices, _ffi.XPUB_INDICES: self._limits.max_indices,
                 _ffi.XPUB_PAGE_RANGES: self._limits.max_page_slots, _ffi.XPUB_COUNTERS: 1}[buffer_id]
        count = total - first if count is None else count
        out = np.empty(count, dtype=dtype)
        if count:
            status = self._lib.hvx_extraction_publisher_read(self._handle, buffer_id, first * dtype.itemsize,
                                                             count * dtype.itemsize, out.ctypes.data_as(C.c_void_p))
            if status != _ffi.HVX_OK:
                self._ctx._check(status)
        return out
