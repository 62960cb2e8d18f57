"""helio_b200 -- B200-native Transvoxel chunk extraction behind Helio's extractor API.

Only the planetary-voxel mesh-extraction hot path lives here: the C-ABI
CUDA library (csrc/, libhelio_voxel_cuda.so) and the host-side mirror of the reference's
extractor objects.  Importing the package does not load the library; creating any extractor
does, and fails loudly if it is missing -- there is no CPU fallback.
"""
from . import _ffi
from .brick import (BRICK_DRAW_DTYPE, BRICK_MESHLET_DTYPE, BRICK_META_DTYPE, DIRTY_BRICK_DTYPE, MAX_SURFACE_INDICES_PER_BRICK,
                    MAX_SURFACE_VERTS_PER_BRICK, VOXEL_MESH_BRICK_VOXEL_WORDS, VOXEL_MESH_MAX_BRICKS, VOXEL_MESH_MAX_DIRTY,
                    ActiveBrickRange, VoxelMeshExtractor, pack_brick)
from .context import (GATHER_COUNTERS_DTYPE, CELL_OFFSET_DTYPE, CELL_RECORD_DTYPE, CLASSIFY_COUNTERS_DTYPE, EMISSION_COUNTERS_DTYPE,
                      MESHLET_BOUNDS_DTYPE, MESHLET_DTYPE, RANGE_DTYPE, TERRAIN_MESHLET_BUILD_INDICES, SCAN_BLOCK_DTYPE, TRANSITION_COUNTERS_DTYPE, VERTEX_DTYPE, Context, make_descs)
from .errors import (AddressError, BatchCapacity, CudaError, DeviceLimit, FinestLodHasNoFinerNeighbor, HvxError,
                     InvalidExtractionCapacity, SampleCount, TerrainLodTopologyError, TransitionDeviceLimit,
                     TransitionInvalidExtractionCapacity, TransitionMask, TransitionSampleCount,
                     TransvoxelGpuError, TransvoxelTransitionGpuError)
from .extraction import (EXTRACTION_COUNTERS_DTYPE, EXTRACTION_RANGE_DTYPE, EXTRACTION_REQUEST_DTYPE, ArenaSlice,
                         BoundedExtractionPublisher, ExtractionAllocationPlan, ExtractionError, ExtractionEvictOutcome,
                         ExtractionLimits, ExtractionPublisherCounters, ExtractionReservation, GpuExtractionRequest,
                         PlanetPageKey, PublicationOutcome, PublishedSurface, ReservationOutcome, SurfaceAllocation,
                         SurfaceCounts)
from .extractor import (EXTRACTION_SAMPLE_COUNT, TRANSITION_ALL_FACE_SLAB_SAMPLE_COUNT,
                        TRANSVOXEL_SCAN_WORKGROUP_SIZE, ChunkBatchExtractor, ResourceStats, TransvoxelGpuClassifier,
                        TransvoxelGpuExtractor, TransvoxelGpuExtractorConfig, TransvoxelGpuTransitionExtractor,
                        TransvoxelGpuTransitionExtractorConfig)
from .gather import (GATHER_JOB_DTYPE, PAGE_TABLE_ENTRY_DTYPE, RESIDENCY_UNIFORM_DTYPE, GpuLookupKey, GpuSurfaceSampler,
                     PageTable, PageTableError, gather_job, residency_uniform)
from .lod import HorizonLodFixturePlan, TerrainLodTopology, TerrainLodTopologyStats, chunk_cost, partition_chunks, start_order
from .publish import (DRAW_INDEXED_INDIRECT_DTYPE, DRAW_PAGE_DTYPE, PAGE_META_DTYPE, SURFACE_FEEDBACK_DTYPE, SURFACE_JOB_DTYPE,
                      SURFACE_STATE_DTYPE, SurfacePublisher, max_meshlets_for_indices)
from .types import (MAX_ADDRESSABLE_LOD, PAGE_EDGE, TRANSITION_FACE_MASK, CellWord, ExtractionFixtureKind,
                    GpuTransvoxelCell, GpuTransvoxelTransitionCell, PageKey, TransitionFace)

__version__ = "0.1.0"
