"""Value types of the path, mirroring helio-planet-voxel-core and the PV fixtures."""
from __future__ import annotations

import enum
from dataclasses import dataclass

PAGE_EDGE = 32                      # helio-planet-voxel-core/src/types.rs:6,11
MAX_ADDRESSABLE_LOD = 57            # types.rs:18
TRANSITION_FACE_MASK = 0b00_111111  # types.rs:19


class TransitionFace(enum.IntEnum):
    """helio-planet-voxel-core/src/types.rs:22-71 (stable bit order)."""
    NegativeX = 0
    PositiveX = 1
    NegativeY = 2
    PositiveY = 3
    NegativeZ = 4
    PositiveZ = 5

    def index(self):
        return int(self)

    def bit(self):
        return 1 << int(self)

    def axis(self):
        return int(self) // 2

    def is_positive(self):
        return int(self) & 1 != 0


class ExtractionFixtureKind(enum.IntEnum):
    """PV/src/fixture.rs:9-16, plus the two synthetic bench fields (16, 17; no reference twin)."""
    Plane = 0
    Sphere = 1
    Cave = 2
    SharpCorner = 3
    ThinSlab = 4
    MaterialSeam = 5
    TerrainFbm = 16
    DenseRandom = 17


class CellWord(int):
    """helio-planet-voxel-core/src/types.rs:328-354: i16 density | material<<16 | flags<<24."""
    AIR = 0x00007FFF

    def __new__(cls, density=0, material=0, flags=0, *, raw=None):
        if raw is not None:
            return super().__new__(cls, int(raw) & 0xFFFFFFFF)
        return super().__new__(cls, (int(density) & 0xFFFF) | ((int(material) & 0xFF) << 16) | ((int(flags) & 0xFF) << 24))

    def density(self):
        d = int(self) & 0xFFFF
        return d - 0x10000 if d & 0x8000 else d

    def material(self):
        return (int(self) >> 16) & 0xFF

    def flags(self):
        return (int(self) >> 24) & 0xFF

    def is_solid(self):
        return self.density() <= 0


@dataclass(frozen=True, order=True)
class PageKey:
    """helio-planet-voxel-core/src/types.rs:231-310; ordering = (lod, page_xyz) like derive(Ord)."""
    lod: int
    page_xyz: tuple

    def __init__(self, lod, page_xyz):
        object.__setattr__(self, "lod", int(lod))
        object.__setattr__(self, "page_xyz", tuple(int(v) for v in page_xyz))

    def lod0_cell_span(self, edge=PAGE_EDGE):
        return edge << self.lod

    def lod0_cell_min(self, edge=PAGE_EDGE):
        span = self.lod0_cell_span(edge)
        return tuple(v * span for v in self.page_xyz)

    def parent(self):
        if self.lod + 1 > MAX_ADDRESSABLE_LOD:
            return None
        return PageKey(self.lod + 1, tuple(v // 2 for v in self.page_xyz))


@dataclass(frozen=True)
class GpuTransvoxelCell:
    """Decoder for PV/src/transvoxel_gpu.rs:77-132 records."""
    packed: int
    generation: int

    def case_index(self):
        return self.packed & 0xFF

    def class_index(self):
        return (self.packed >> 8) & 0xFF

    def vertex_count(self):
        return (self.packed >> 16) & 0xFF

    def triangle_count(self):
        return (self.packed >> 24) & 0x0F

    def is_valid_for(self, generation):
        return bool(self.packed & 0x80000000) and self.generation == generation


@dataclass(frozen=True)
class GpuTransvoxelTransitionCell:
    """Decoder for PV/src/transvoxel_transition_gpu.rs:67-131 records."""
    packed: int
    generation: int

    def case_index(self):
        return self.packed & 0x1FF

    def class_code(self):
        return (self.packed >> 9) & 0xFF

    def class_index(self):
        return self.class_code() & 0x7F

    def reverse_winding(self):
        return bool(self.class_code() & 0x80)

    def vertex_count(self):
        return (self.packed >> 17) & 0x0F

    def triangle_count(self):
        return (self.packed >> 21) & 0x0F

    def is_valid_for(self, generation):
        return bool(self.packed & 0x80000000) and self.generation == generation
