"""Host-side mirror of the reference's surface-gather interface.

Reference: ``PlanetarySurfaceRequest`` / ``GpuSurfaceGatherJob`` / ``GpuSurfaceSampler``
(PV/src/surface_sampling.rs:18-337), ``GpuPageTableEntry`` / ``GpuResidencyUniform`` /
``PageTable`` (PV/src/table.rs).  The residency subsystem that owns the atlas is out of scope; this
module only provides what a caller needs to hand an atlas + page table to ``hvx_gather_surface`` and
then run the extractors on the gathered arenas without a host round trip.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _ffi
from .context import GATHER_COUNTERS_DTYPE, Context, make_descs

PAGE_EDGE = 32
PAGE_TABLE_EMPTY, PAGE_TABLE_OCCUPIED, PAGE_TABLE_TOMBSTONE = 0, 1, 2
REGULAR_EXTRACTION_INDIRECT_OFFSETS = (0, 12, 24, 36)        # surface_sampling.rs:15
TRANSITION_EXTRACTION_INDIRECT_OFFSETS = (48, 60, 72, 84)    # surface_sampling.rs:16

PAGE_TABLE_ENTRY_DTYPE = np.dtype([("planet_id", "<u4", 4), ("relative_lod0_cell_min", "<i4", 3), ("lod", "<u4"),
                                   ("slot", "<u4"), ("generation_low", "<u4"), ("generation_high", "<u4"), ("state", "<u4")])
RESIDENCY_UNIFORM_DTYPE = np.dtype([(n, "<u4") for n in (
    "table_mask", "max_probe", "resident_pages", "atlas_tiles_x", "atlas_tiles_y", "atlas_tiles_z",
    "publication_epoch_low", "publication_epoch_high")])
GATHER_JOB_DTYPE = np.dtype([("planet_id", "<u4", 4), ("relative_lod0_cell_min", "<i4", 3), ("lod", "<u4"),
                             ("generation_low", "<u4"), ("generation_high", "<u4"), ("transition_mask", "<u4"),
                             ("target_slot", "<u4"), ("residency_epoch_low", "<u4"), ("residency_epoch_high", "<u4"),
                             ("_pad", "<u4", 2)])
assert PAGE_TABLE_ENTRY_DTYPE.itemsize == 48 and RESIDENCY_UNIFORM_DTYPE.itemsize == 32 and GATHER_JOB_DTYPE.itemsize == 64


def mix_hash(value_hash: int, value: int) -> int:
    """PV/src/table.rs:163-167."""
    mixed = ((value_hash ^ value) * 0x045D9F3B) & 0xFFFFFFFF
    return mixed ^ (mixed >> 16)


@dataclass(frozen=True)
class GpuLookupKey:
    """PV/src/table.rs:126-161."""
    planet_id: tuple
    relative_lod0_cell_min: tuple
    lod: int

    def hash(self) -> int:
        h = 0x811C9DC5
        for v in (*self.planet_id, *(int(c) & 0xFFFFFFFF for c in self.relative_lod0_cell_min), self.lod):
            h = mix_hash(h, int(v) & 0xFFFFFFFF)
        return h


class PageTableError(ValueError):
    pass


class PageTable:
    """Open-addressed page table with linear probing and tombstones (PV/src/table.rs:170-320)."""

    def __init__(self, capacity: int, max_probe: int):
        if capacity <= 0 or capacity & (capacity - 1):
            raise PageTableError(f"page table capacity {capacity} is not a power of two")
        if max_probe == 0 or max_probe > capacity:
            raise PageTableError(f"max_probe {max_probe} must be in [1, {capacity}]")
        self._entries = np.zeros(capacity, dtype=PAGE_TABLE_ENTRY_DTYPE)
        self._max_probe, self._occupied, self._tombstones = max_probe, 0, 0

    def entries(self) -> np.ndarray:
        return self._entries

    def capacity(self) -> int:
        return len(self._entries)

    def max_probe(self) -> int:
        return self._max_probe

    def occupied(self) -> int:
        return self._occupied

    def tombstones(self) -> int:
        return self._tombstones

    def _find(self, key: GpuLookupKey):
        mask = self.capacity() - 1
        start, first_tombstone = key.hash() & mask, None
        for probe in range(self._max_probe):
            index = (start + probe) & mask
            e = self._entries[index]
            state = int(e["state"])
            if state == PAGE_TABLE_EMPTY:
                return "vacant", index if first_tombstone is None else first_tombstone
            if state == PAGE_TABLE_TOMBSTONE:
                if first_tombstone is None:
                    first_tombstone = index
            elif tuple(int(v) for v in e["planet_id"]) == tuple(key.planet_id) and \
                    tuple(int(v) for v in e["relative_lod0_cell_min"]) == tuple(key.relative_lod0_cell_min) and int(e["lod"]) == key.lod:
                return "found", index
        return ("saturated", None) if first_tombstone is None else ("vacant", first_tombstone)

    def insert(self, key: GpuLookupKey, slot: int, generation: int) -> int:
        kind, index = self._find(key)
        if kind == "saturated":
            raise PageTableError(f"probe sequence of hash {key.hash():#x} is saturated after {self._max_probe} entries")
        if kind == "vacant":
            if int(self._entries[index]["state"]) == PAGE_TABLE_TOMBSTONE:
                self._tombstones -= 1
            self._occupied += 1
        self._entries[index] = (key.planet_id, key.relative_lod0_cell_min, key.lod, slot, generation & 0xFFFFFFFF,
                                generation >> 32, PAGE_TABLE_OCCUPIED)
        return index

    def remove(self, key: GpuLookupKey):
        kind, index = self._find(key)
        if kind != "found":
            return None
        removed = self._entries[index].copy()
        self._entries[index] = ((0,) * 4, (0,) * 3, 0, 0, 0, 0, PAGE_TABLE_TOMBSTONE)
        self._occupied -= 1
        self._tombstones += 1
        return removed

    def lookup(self, key: GpuLookupKey):
        kind, index = self._find(key)
        return (index, self._entries[index].copy()) if kind == "found" else None


def residency_uniform(table: PageTable, atlas_tiles, resident_pages: int, publication_epoch: int) -> np.ndarray:
    """``GpuResidencyUniform`` for a table + atlas (PV/src/table.rs:62-72)."""
    out = np.zeros(1, dtype=RESIDENCY_UNIFORM_DTYPE)
    out[0] = (table.capacity() - 1, table.max_probe(), resident_pages, *(int(t) for t in atlas_tiles),
              publication_epoch & 0xFFFFFFFF, publication_epoch >> 32)
    return out


def gather_job(planet_id, relative_lod0_cell_min, lod: int, generation: int, transition_mask: int, target_slot: int,
               residency_epoch: int) -> np.ndarray:
    """``GpuSurfaceGatherJob::new`` (PV/src/surface_sampling.rs:136-168) from already-resolved metadata."""
    job = np.zeros(1, dtype=GATHER_JOB_DTYPE)
    job[0] = (tuple(planet_id), tuple(relative_lod0_cell_min), lod, generation & 0xFFFFFFFF, generation >> 32,
              transition_mask, target_slot, residency_epoch & 0xFFFFFFFF, residency_epoch >> 32, (0, 0))
    return job


class GpuSurfaceSampler:
    """Batched ``GpuSurfaceSampler`` (PV/src/surface_sampling.rs:184-350): prepare + encode for n jobs per call.

    ``dispatch`` gathers into the context's sample / slab arenas; ``extract`` then runs the regular and
    (where a job owns faces) the transition extractor on them, taking generation and transition mask
    from the jobs -- the production call stack of SURVEY 3.2 without the 480 KB/page host round trip.
    """

    def __init__(self, context: Context):
        if context.edge != PAGE_EDGE:
            raise ValueError("the surface gather works on 32-cell residency pages; create the context with edge=32")
        self._ctx = context
        self._jobs = None

    def bind_table(self, table) -> None:
        """Upload the page table once; later ``dispatch(residency, None, ...)`` calls use the resident copy."""
        table_entries = table.entries() if isinstance(table, PageTable) else table
        self._ctx.bind_page_table(None if table_entries is None else np.ascontiguousarray(table_entries, dtype=PAGE_TABLE_ENTRY_DTYPE))

    def dispatch(self, residency, table, atlas, jobs) -> None:
        table_entries = table.entries() if isinstance(table, PageTable) else table
        self._jobs = np.ascontiguousarray(jobs, dtype=GATHER_JOB_DTYPE).copy()
        self._ctx.gather_surface(np.ascontiguousarray(residency, dtype=RESIDENCY_UNIFORM_DTYPE),
                                 None if table_entries is None else np.ascontiguousarray(table_entries, dtype=PAGE_TABLE_ENTRY_DTYPE),
                                 atlas, self._jobs)

    def counters_buffer(self, n=None) -> np.ndarray:
        n = len(self._jobs) if n is None else n
        return self._ctx.read(_ffi.BUF_GATHER_COUNTERS, 0, n).view(GATHER_COUNTERS_DTYPE)

    def indirect_buffer(self, n=None) -> np.ndarray:
        n = len(self._jobs) if n is None else n
        return self._ctx.read(_ffi.BUF_GATHER_INDIRECT, 0, n * 24).reshape(n, 8, 3)

    def regular_samples(self, job: int) -> np.ndarray:
        return self._ctx.read(_ffi.BUF_SAMPLES, job * 34 ** 3, 34 ** 3)

    def transition_samples(self, job: int) -> np.ndarray:
        return self._ctx.read(_ffi.BUF_SLABS, job * 6 * 3 * 67 * 67, 6 * 3 * 67 * 67)

    def extract(self, dirty_microbricks=(1 << 64) - 1, transition=True) -> None:
        """Regular (and transition) extraction of the gathered jobs, straight from the arenas."""
        jobs, n = self._jobs, len(self._jobs)
        generation = jobs["generation_low"].astype(np.uint64) | (jobs["generation_high"].astype(np.uint64) << np.uint64(32))
        masks = jobs["transition_mask"]
        descs = make_descs(n, generation, dirty_microbricks, masks)
        self._ctx.extract_regular(None, descs, n)
        if transition and masks.any() and self._ctx.max_transition_vertices:
            self._ctx.extract_transition(None, descs, n)
