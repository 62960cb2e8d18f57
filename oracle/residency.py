"""CPU restatement of the residency pieces the surface gather reads.  TEST INFRASTRUCTURE.

Follows, function by function:
  * GpuLookupKey::hash / mix_hash            PV/src/table.rs:148-167
  * PageTable::{find_slot, insert, remove, lookup, compact}   PV/src/table.rs:246-320
  * PageKey::{lod0_cell_span, lod0_cell_min, relative_lod0_cell_min, address_lod0_cell}
                                              crates/helio-planet-voxel-core/src/types.rs:258-310
  * PlanetarySurfaceRequest::required_pages   PV/src/surface_sampling.rs:42-123
  * canonical_cell / canonical_page_cells     PV/src/surface_sampling.rs:754-781 (the reference test's golden field)
Pure Python / numpy; the per-sample gather itself is in hvx_oracle.c (hvxo_gather_surface).
"""
from __future__ import annotations

import numpy as np

PAGE_EDGE = 32
EMPTY, OCCUPIED, TOMBSTONE = 0, 1, 2

ENTRY_DTYPE = np.dtype([("planet_id", "<u4", 4), ("relative_lod0_cell_min", "<i4", 3), ("lod", "<u4"), ("slot", "<u4"),
                        ("generation_low", "<u4"), ("generation_high", "<u4"), ("state", "<u4")])
RESIDENCY_DTYPE = np.dtype([(n, "<u4") for n in ("table_mask", "max_probe", "resident_pages", "atlas_tiles_x", "atlas_tiles_y",
                                                 "atlas_tiles_z", "publication_epoch_low", "publication_epoch_high")])
JOB_DTYPE = np.dtype([("planet_id", "<u4", 4), ("relative_lod0_cell_min", "<i4", 3), ("lod", "<u4"), ("generation_low", "<u4"),
                      ("generation_high", "<u4"), ("transition_mask", "<u4"), ("target_slot", "<u4"),
                      ("residency_epoch_low", "<u4"), ("residency_epoch_high", "<u4"), ("_pad", "<u4", 2)])
COUNTERS_DTYPE = np.dtype([(n, "<u4") for n in ("regular_samples", "transition_samples", "table_probes", "page_misses",
                                                "stale_targets", "completed")] + [("_pad", "<u4", 2)])
assert ENTRY_DTYPE.itemsize == 48 and RESIDENCY_DTYPE.itemsize == 32 and JOB_DTYPE.itemsize == 64 and COUNTERS_DTYPE.itemsize == 32

# PV/src/transvoxel_transition.rs:463-502 / surface_gather.wgsl:80-100
FACE_ORIGIN = [(0, 0, 1), (1, 0, 0), (1, 0, 0), (0, 1, 0), (0, 1, 0), (0, 0, 1)]
FACE_U = [(0, 1, 0), (0, 1, 0), (0, 0, 1), (0, 0, 1), (1, 0, 0), (1, 0, 0)]
FACE_V = [(0, 0, -1), (0, 0, 1), (-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0)]
FACE_OUT = [(-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1)]


def planet_words(planet_bytes: bytes) -> list[int]:
    """PlanetId([u8; 16]) as four little-endian words (surface_sampling.rs:141-150)."""
    assert len(planet_bytes) == 16
    return [int.from_bytes(planet_bytes[4 * i:4 * i + 4], "little") for i in range(4)]


def mix_hash(h: int, value: int) -> int:
    mixed = ((h ^ value) * 0x045D9F3B) & 0xFFFFFFFF
    return mixed ^ (mixed >> 16)


def key_hash(planet_id, relative_min, lod: int) -> int:
    h = 0x811C9DC5
    for v in list(planet_id) + [int(x) & 0xFFFFFFFF for x in relative_min] + [lod]:
        h = mix_hash(h, int(v) & 0xFFFFFFFF)
    return h


def lod0_cell_span(lod: int) -> int:
    return PAGE_EDGE << lod


def lod0_cell_min(lod: int, page_xyz) -> list[int]:
    return [int(c) * lod0_cell_span(lod) for c in page_xyz]


def relative_lod0_cell_min(lod: int, page_xyz, frame_origin) -> list[int]:
    rel = [a - int(o) for a, o in zip(lod0_cell_min(lod, page_xyz), frame_origin)]
    for r in rel:
        if not -(1 << 31) <= r < (1 << 31):
            raise OverflowError("OutsideRenderFrame")
    return rel


def address_lod0_cell(lod: int, cell_xyz) -> list[int]:
    """Page that contains an LOD0 cell (div_euclid by the page span)."""
    span = lod0_cell_span(lod)
    return [int(c) // span for c in cell_xyz]


def required_pages(lod: int, page_xyz, transition_mask: int) -> set[tuple[int, tuple[int, int, int]]]:
    """(lod, page_xyz) of every page the halo block and the enabled fine-side slabs read."""
    if transition_mask & ~0x3F:
        raise ValueError("TransitionMask")
    if lod == 0 and transition_mask:
        raise ValueError("FinestLodTransition")
    pages: set[tuple[int, tuple[int, int, int]]] = set()
    page_min = lod0_cell_min(lod, page_xyz)
    coarse = 1 << lod

    def insert_box(l, lo, hi):
        plo, phi = address_lod0_cell(l, lo), address_lod0_cell(l, hi)
        for z in range(plo[2], phi[2] + 1):
            for y in range(plo[1], phi[1] + 1):
                for x in range(plo[0], phi[0] + 1):
                    pages.add((l, (x, y, z)))

    insert_box(lod, [v - coarse for v in page_min], [v + PAGE_EDGE * coarse for v in page_min])
    if transition_mask == 0:
        return pages
    fine, span = coarse // 2, PAGE_EDGE * coarse
    for face in range(6):
        if not (transition_mask >> face) & 1:
            continue
        lo, hi = [1 << 62] * 3, [-(1 << 62)] * 3
        for u in (-1, 65):  # TRANSITION_FACE_SAMPLE_EDGE = 65
            for v in (-1, 65):
                for o in (-1, 1):
                    for a in range(3):
                        p = page_min[a] + FACE_ORIGIN[face][a] * span + FACE_U[face][a] * u * fine + FACE_V[face][a] * v * fine + \
                            FACE_OUT[face][a] * o * fine
                        lo[a], hi[a] = min(lo[a], p), max(hi[a], p)
        insert_box(lod - 1, lo, hi)
    return pages


class PageTable:
    """Open-addressed page table with tombstones (PV/src/table.rs:170-320)."""

    def __init__(self, capacity: int, max_probe: int):
        if capacity & (capacity - 1) or capacity == 0:
            raise ValueError("CapacityNotPowerOfTwo")
        if max_probe == 0 or max_probe > capacity:
            raise ValueError("InvalidMaxProbe")
        self.entries = np.zeros(capacity, dtype=ENTRY_DTYPE)
        self.max_probe, self.occupied, self.tombstones = max_probe, 0, 0

    @property
    def capacity(self) -> int:
        return len(self.entries)

    def _find_slot(self, planet_id, rel, lod):
        mask = self.capacity - 1
        start = key_hash(planet_id, rel, lod) & mask
        first_tombstone = None
        for probe in range(self.max_probe):
            index = (start + probe) & mask
            e = self.entries[index]
            if e["state"] == EMPTY:
                return "vacant", (index if first_tombstone is None else first_tombstone)
            if e["state"] == TOMBSTONE and first_tombstone is None:
                first_tombstone = index
            elif e["state"] == OCCUPIED and list(e["planet_id"]) == list(planet_id) and \
                    list(e["relative_lod0_cell_min"]) == list(rel) and e["lod"] == lod:
                return "found", index
        return ("saturated", None) if first_tombstone is None else ("vacant", first_tombstone)

    def insert(self, planet_id, rel, lod, slot, generation) -> int:
        kind, index = self._find_slot(planet_id, rel, lod)
        if kind == "saturated":
            raise RuntimeError("ProbeSaturated")
        if kind == "vacant":
            if self.entries[index]["state"] == TOMBSTONE:
                self.tombstones -= 1
            self.occupied += 1
        self.entries[index] = (planet_id, rel, lod, slot, generation & 0xFFFFFFFF, generation >> 32, OCCUPIED)
        return index

    def remove(self, planet_id, rel, lod) -> bool:
        kind, index = self._find_slot(planet_id, rel, lod)
        if kind != "found":
            return False
        self.entries[index] = ([0] * 4, [0] * 3, 0, 0, 0, 0, TOMBSTONE)
        self.occupied -= 1
        self.tombstones += 1
        return True

    def lookup(self, planet_id, rel, lod):
        kind, index = self._find_slot(planet_id, rel, lod)
        return (index, self.entries[index].copy()) if kind == "found" else None


def canonical_cell(position) -> int:
    """The reference gather test's golden field (surface_sampling.rs:772-781), as a CellWord."""
    def wrap(v):
        return ((v + (1 << 63)) & ((1 << 64) - 1)) - (1 << 63)
    mixed = wrap(wrap(position[0] * 17) + wrap(position[1] * 31) + wrap(position[2] * 43))
    density = (mixed % 30_001) - 15_000
    material = (wrap(mixed * 7) % 255) + 1
    flags = wrap(mixed * 11) % 64
    return (density & 0xFFFF) | (material << 16) | (flags << 24)


def canonical_cells_np(x, y, z) -> np.ndarray:
    """Vectorised canonical_cell for int64 coordinate arrays (no wrap occurs for |coord| < 2^50)."""
    mixed = x.astype(np.int64) * 17 + y.astype(np.int64) * 31 + z.astype(np.int64) * 43
    density = np.mod(mixed, 30_001) - 15_000
    material = np.mod(mixed * 7, 255) + 1
    flags = np.mod(mixed * 11, 64)
    return ((density & 0xFFFF) | (material << 16) | (flags << 24)).astype(np.uint32)


def canonical_page_cells(lod: int, page_xyz) -> np.ndarray:
    """32^3 words of one page, x fastest (surface_sampling.rs:754-770)."""
    mn, scale = lod0_cell_min(lod, page_xyz), 1 << lod
    r = np.arange(PAGE_EDGE, dtype=np.int64) * scale
    z, y, x = np.meshgrid(mn[2] + r, mn[1] + r, mn[0] + r, indexing="ij")
    return canonical_cells_np(x, y, z).reshape(-1)


def expected_regular(lod: int, page_xyz) -> np.ndarray:
    """What the gather must produce for the halo block of a fully resident neighbourhood (:682-699)."""
    mn, scale = lod0_cell_min(lod, page_xyz), 1 << lod
    r = (np.arange(34, dtype=np.int64) - 1) * scale
    z, y, x = np.meshgrid(mn[2] + r, mn[1] + r, mn[0] + r, indexing="ij")
    return canonical_cells_np(x, y, z).reshape(-1)


def expected_transition(lod: int, page_xyz) -> np.ndarray:
    """All six 67x67x3 fine-side slabs of a fully resident neighbourhood (:707-735)."""
    mn, scale = lod0_cell_min(lod, page_xyz), 1 << lod
    fine, span = scale // 2, PAGE_EDGE * scale
    out = np.empty(6 * 3 * 67 * 67, dtype=np.uint32)
    layer, v, u = np.meshgrid(np.arange(3, dtype=np.int64), np.arange(67, dtype=np.int64), np.arange(67, dtype=np.int64), indexing="ij")
    for face in range(6):
        pos = [mn[a] + FACE_ORIGIN[face][a] * span + FACE_U[face][a] * (u - 1) * fine + FACE_V[face][a] * (v - 1) * fine +
               FACE_OUT[face][a] * (layer - 1) * fine for a in range(3)]
        out[face * 13467:(face + 1) * 13467] = canonical_cells_np(pos[0], pos[1], pos[2]).reshape(-1)
    return out


class Atlas:
    """Linear R32Uint atlas (x fastest) + page table + residency uniform, filled page by page."""

    def __init__(self, tiles, table_capacity: int, max_probe: int, epoch: int = 1):
        self.tiles = tuple(int(t) for t in tiles)
        self.words = np.full((self.tiles[2] * 32, self.tiles[1] * 32, self.tiles[0] * 32), 0xDEADBEEF, dtype=np.uint32)
        self.table = PageTable(table_capacity, max_probe)
        self.epoch = epoch
        self.next_slot = 0
        self.resident = 0

    def upload(self, planet_id, lod, page_xyz, frame_origin, cells: np.ndarray, generation: int, slot: int | None = None) -> int:
        if slot is None:
            slot = self.next_slot
            self.next_slot += 1
        tx, ty, tz = slot % self.tiles[0], (slot // self.tiles[0]) % self.tiles[1], slot // (self.tiles[0] * self.tiles[1])
        assert tz < self.tiles[2], "atlas full"
        self.words[tz * 32:(tz + 1) * 32, ty * 32:(ty + 1) * 32, tx * 32:(tx + 1) * 32] = cells.reshape(32, 32, 32)
        self.table.insert(planet_id, relative_lod0_cell_min(lod, page_xyz, frame_origin), lod, slot, generation)
        self.resident += 1
        return slot

    def residency(self) -> np.ndarray:
        r = np.zeros(1, dtype=RESIDENCY_DTYPE)
        r[0] = (self.table.capacity - 1, self.table.max_probe, self.resident, *self.tiles, self.epoch & 0xFFFFFFFF, self.epoch >> 32)
        return r


def make_job(planet_id, lod, page_xyz, frame_origin, generation: int, transition_mask: int, slot: int, epoch: int) -> np.ndarray:
    j = np.zeros(1, dtype=JOB_DTYPE)
    j[0] = (planet_id, relative_lod0_cell_min(lod, page_xyz, frame_origin), lod, generation & 0xFFFFFFFF, generation >> 32,
            transition_mask, slot, epoch & 0xFFFFFFFF, epoch >> 32, [0, 0])
    return j
