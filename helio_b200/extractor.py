"""Host-side mirror of the reference's extractor objects, over the C ABI.

Same names, argument order and error behaviour as

* ``TransvoxelGpuExtractorConfig`` / ``TransvoxelGpuExtractor``     PV/src/transvoxel_emit.rs:57-396
* ``TransvoxelGpuClassifier``                                       PV/src/transvoxel_gpu.rs:148-356
* ``TransvoxelGpuTransitionExtractorConfig`` / ``...Extractor``      PV/src/transvoxel_transition_gpu.rs:148-520

so the parity tests read like the reference's own (PV/tests/gpu_transvoxel*.rs).  Where the
reference takes ``(&Device, &Queue)`` these take a CUDA device ordinal at construction.
``ChunkBatchExtractor`` is the batch form the bench and the multi-GPU scheduler use.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _ffi
from .context import Context, make_descs
from .errors import InvalidExtractionCapacity, TransitionInvalidExtractionCapacity

TRANSVOXEL_SCAN_WORKGROUP_SIZE = 256  # PV/src/transvoxel_gpu.rs:11
EXTRACTION_SAMPLE_COUNT = 34 ** 3    # PV/src/fixture.rs:5-6
TRANSITION_ALL_FACE_SLAB_SAMPLE_COUNT = 6 * 3 * 67 * 67  # PV/src/transvoxel_transition.rs:22-24


@dataclass(frozen=True)
class TransvoxelGpuExtractorConfig:
    """PV/src/transvoxel_emit.rs:57-85 (defaults 32768*12 / 32768*15)."""
    max_vertices: int = 393_216
    max_indices: int = 491_520

    @classmethod
    def new(cls, max_vertices, max_indices):
        if max_vertices == 0 or max_indices == 0:
            raise InvalidExtractionCapacity(
                f"Transvoxel extraction capacities must be nonzero (vertices={max_vertices}, indices={max_indices})",
                _ffi.HVX_E_INVALID_CAPACITY, max_vertices, max_indices)
        return cls(max_vertices, max_indices)


@dataclass(frozen=True)
class TransvoxelGpuTransitionExtractorConfig:
    """PV/src/transvoxel_transition_gpu.rs:148-181 (defaults 6144*12 / 6144*36)."""
    max_vertices: int = 73_728
    max_indices: int = 221_184

    @classmethod
    def new(cls, max_vertices, max_indices):
        if max_vertices == 0 or max_indices == 0:
            raise TransitionInvalidExtractionCapacity(
                f"Transvoxel transition capacities must be nonzero (vertices={max_vertices}, indices={max_indices})",
                _ffi.HVX_E_INVALID_CAPACITY, max_vertices, max_indices)
        return cls(max_vertices, max_indices)


@dataclass(frozen=True)
class ResourceStats:
    """TransvoxelExtractorResourceStats (PV/src/transvoxel_emit.rs:87-91)."""
    buffers: int
    allocated_bytes: int


class _Single:
    """Shared plumbing of the one-page-in-flight mirrors."""

    def __init__(self, ctx):
        self._ctx = ctx

    @property
    def context(self) -> Context:
        return self._ctx

    def resource_stats(self):
        return ResourceStats(buffers=sum(1 for i in range(25) if self._ctx.buffer_bytes(i)),
                             allocated_bytes=self._ctx.allocated_bytes)

    def resize(self, _width, _height):
        """Extraction owns no surface-size-dependent resources (same no-op as the reference)."""

    def close(self):
        self._ctx.close()


class TransvoxelGpuExtractor(_Single):
    """Regular-cell extractor for one page per dispatch (the reference's shape)."""

    def __init__(self, device=0, config=None, *, edge=32, debug_records=True, first_generation=False):
        """``debug_records``: also produce the per-cell records behind cells_buffer / offsets_buffer / blocks_buffer
        (two extra launches per dispatch; the meshes come from the same kernel either way).
        ``first_generation``: diagnostics, extract with the first-generation kernel (identical output)."""
        config = config or TransvoxelGpuExtractorConfig()
        TransvoxelGpuExtractorConfig.new(config.max_vertices, config.max_indices)
        super().__init__(Context(device, edge=edge, max_chunks=1, max_vertices=config.max_vertices,
                                 max_indices=config.max_indices, debug_records=debug_records,
                                 first_generation=first_generation))
        self._config = config

    def config(self):
        return self._config

    def dispatch(self, samples, generation, dirty_microbricks, transition_mask):
        """PV/src/transvoxel_emit.rs:233-254.  ``samples``: (edge+2)^3 CellWords, numpy or torch (host/device)."""
        count = samples.numel() if hasattr(samples, "numel") else np.asarray(samples).size
        self._ctx.extract_regular(samples, make_descs(1, generation, dirty_microbricks, transition_mask), 1,
                                  sample_words=count)

    # buffer accessors (PV/src/transvoxel_emit.rs:367-385), returned as host copies
    def counters_buffer(self):
        return self._ctx.read(_ffi.BUF_REGULAR_COUNTERS, 0, 1)[0]

    def classify_counters_buffer(self):
        return self._ctx.read(_ffi.BUF_REGULAR_CLASSIFY, 0, 1)[0]

    def vertices_buffer(self, count=None):
        return self._ctx.read(_ffi.BUF_REGULAR_VERTICES, 0, self._config.max_vertices if count is None else count)

    def indices_buffer(self, count=None):
        return self._ctx.read(_ffi.BUF_REGULAR_INDICES, 0, self._config.max_indices if count is None else count)

    def cells_buffer(self):
        return self._ctx.read(_ffi.BUF_REGULAR_CELLS, 0, self._ctx.cells)

    def offsets_buffer(self):
        return self._ctx.read(_ffi.BUF_REGULAR_OFFSETS, 0, self._ctx.cells)

    def blocks_buffer(self):
        return self._ctx.read(_ffi.BUF_REGULAR_BLOCKS, 0, self._ctx.cells // 256)


class TransvoxelGpuClassifier(_Single):
    """Classification only (PV/src/transvoxel_gpu.rs:148-356)."""

    def __init__(self, device=0, *, edge=32, first_generation=False):
        super().__init__(Context(device, edge=edge, max_chunks=1, debug_records=True, first_generation=first_generation))

    def dispatch(self, samples, generation, dirty_microbricks):
        count = samples.numel() if hasattr(samples, "numel") else np.asarray(samples).size
        self._ctx.extract_regular(samples, make_descs(1, generation, dirty_microbricks, 0), 1, sample_words=count,
                                  classify_only=True)

    def output_buffer(self):
        return self._ctx.read(_ffi.BUF_REGULAR_CELLS, 0, self._ctx.cells)

    def counters_buffer(self):
        return self._ctx.read(_ffi.BUF_REGULAR_CLASSIFY, 0, 1)[0]


class TransvoxelGpuTransitionExtractor(_Single):
    """Transition-cell extractor for one coarse page per dispatch."""

    def __init__(self, device=0, config=None, *, edge=32, debug_records=True):
        config = config or TransvoxelGpuTransitionExtractorConfig(edge * edge * 6 * 12, edge * edge * 6 * 36)
        TransvoxelGpuTransitionExtractorConfig.new(config.max_vertices, config.max_indices)
        super().__init__(Context(device, edge=edge, max_chunks=1, max_vertices=1, max_indices=1,
                                 max_transition_vertices=config.max_vertices,
                                 max_transition_indices=config.max_indices, debug_records=debug_records,
                                 error_kind="transition"))
        self._config = config

    def config(self):
        return self._config

    def dispatch(self, face_slabs, transition_mask, generation):
        """PV/src/transvoxel_transition_gpu.rs:366-380 (note the argument order: mask, then generation)."""
        count = face_slabs.numel() if hasattr(face_slabs, "numel") else np.asarray(face_slabs).size
        self._ctx.extract_transition(face_slabs, make_descs(1, generation, (1 << 64) - 1, transition_mask), 1,
                                     slab_words=count)

    def counters_buffer(self):
        return self._ctx.read(_ffi.BUF_TRANSITION_COUNTERS, 0, 1)[0]

    def vertices_buffer(self, count=None):
        return self._ctx.read(_ffi.BUF_TRANSITION_VERTICES, 0, self._config.max_vertices if count is None else count)

    def indices_buffer(self, count=None):
        return self._ctx.read(_ffi.BUF_TRANSITION_INDICES, 0, self._config.max_indices if count is None else count)

    def cells_buffer(self):
        return self._ctx.read(_ffi.BUF_TRANSITION_CELLS, 0, self._ctx.transition_cells)

    def offsets_buffer(self):
        return self._ctx.read(_ffi.BUF_TRANSITION_OFFSETS, 0, self._ctx.transition_cells)

    def blocks_buffer(self):
        return self._ctx.read(_ffi.BUF_TRANSITION_BLOCKS, 0, self._ctx.transition_cells // 256)


class ChunkBatchExtractor:
    """Batch form: N chunks per dispatch into fixed-stride per-chunk slots.

    This is what the planet renderer's page queue would feed once it stops submitting one page
    per frame (docs/planetary_voxel_progress.md:107-108), and what bench.py times.
    """

    def __init__(self, device=0, *, edge=64, max_chunks=256, max_vertices=49_152, max_indices=73_728,
                 max_transition_vertices=0, max_transition_indices=0, debug_records=False, first_generation=False):
        self.ctx = Context(device, edge=edge, max_chunks=max_chunks, max_vertices=max_vertices,
                           max_indices=max_indices, max_transition_vertices=max_transition_vertices,
                           max_transition_indices=max_transition_indices, debug_records=debug_records,
                           first_generation=first_generation)

    def fill_density(self, kind, page_xyz, lod=None, out_ptr=None):
        return self.ctx.fill_density(kind, page_xyz, lod, out_ptr)

    def fill_slabs(self, kind, page_xyz, lod, out_ptr=None):
        return self.ctx.fill_slabs(kind, page_xyz, lod, out_ptr)

    def extract_regular(self, samples, n, generation=1, dirty_microbricks=(1 << 64) - 1, transition_mask=0, descs=None):
        descs = descs if descs is not None else make_descs(n, generation, dirty_microbricks, transition_mask)
        self.ctx.extract_regular(samples, descs, n)

    def extract_transition(self, slabs, n, transition_mask, generation=1, descs=None):
        descs = descs if descs is not None else make_descs(n, generation, (1 << 64) - 1, transition_mask)
        self.ctx.extract_transition(slabs, descs, n)

    def counters(self, n=None):
        return self.ctx.read(_ffi.BUF_REGULAR_COUNTERS, 0, n)

    def classify_counters(self, n=None):
        return self.ctx.read(_ffi.BUF_REGULAR_CLASSIFY, 0, n)

    def ranges(self, n=None):
        return self.ctx.read(_ffi.BUF_REGULAR_RANGES, 0, n)

    def transition_counters(self, n=None):
        return self.ctx.read(_ffi.BUF_TRANSITION_COUNTERS, 0, n)

    def chunk_mesh(self, chunk, kind=0):
        """Host copy of one chunk's (vertices, indices)."""
        v, i, _ = self.ctx.read_meshes(kind, chunk, 1)
        return v, i

    def synchronize(self):
        self.ctx.synchronize()

    def close(self):
        self.ctx.close()
