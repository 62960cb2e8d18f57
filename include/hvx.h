/* hvx.h -- C ABI of libhelio_voxel_cuda.so: B200-native Transvoxel chunk extraction.
 *
 * This is the drop-in boundary for ONE path of Far-Beyond-Pulsar/Helio: the
 * planetary-voxel mesh extractor.  Every entry point names the reference
 * interface it replaces (paths relative to the reference tree; PV =
 * crates/passes/3d/helio-pass-planetary-voxel).  The Rust crate
 * `helio-voxel-cuda` (rust/helio-voxel-cuda) binds exactly
 * these symbols; so do helio_b200/ (ctypes) and include/helio_voxel_cuda.hpp.
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns HVX_OK (0) or a
 *     negative hvx_status; hvx_last_error() gives the formatted message.
 *   - one hvx_ctx = one device + one stream + one set of output arenas.  Like
 *     the reference extractor (`&self`, one page in flight) a ctx is not
 *     re-entrant: calls on one ctx must be serialised by the caller.
 *   - extraction calls are asynchronous on the ctx stream; hvx_synchronize()
 *     or any hvx_read_*() waits.
 *   - byte layouts of vertices, counters and per-cell records equal the
 *     reference PODs so Rust can bytemuck::cast_slice them.
 *   - there is no CPU fallback: without a CUDA device hvx_create fails with
 *     HVX_E_CUDA.
 *   - objects created FROM a ctx (hvx_publisher, hvx_brick_mesher, an attached
 *     hvx_extraction_publisher) borrow its device and stream: destroy them
 *     before the ctx.
 */
#ifndef HVX_H
#define HVX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HVX_ABI_VERSION 2u /* 2: hvx_chunk_desc grew flags (32 bytes) */

/* Errors.  -1..-4 mirror PV/src/transvoxel_gpu.rs:445-459 TransvoxelGpuError and
 * PV/src/transvoxel_transition_gpu.rs:712-732 TransvoxelTransitionGpuError. */
typedef enum {
    HVX_OK = 0,
    HVX_E_SAMPLE_COUNT = -1,        /* SampleCount { actual, expected } */
    HVX_E_INVALID_CAPACITY = -2,    /* InvalidExtractionCapacity */
    HVX_E_DEVICE_LIMIT = -3,        /* DeviceLimit { name, required, available } */
    HVX_E_TRANSITION_MASK = -4,     /* TransitionMask(u8) */
    HVX_E_INVALID_ARGUMENT = -5,
    HVX_E_CUDA = -6,
    HVX_E_BATCH_CAPACITY = -7,      /* n > hvx_config.max_chunks */
    HVX_E_FINEST_LOD = -8,          /* TransvoxelTransitionError::FinestLodHasNoFinerNeighbor */
    HVX_E_ADDRESS = -9,             /* AddressError::{UnsupportedLod, CoordinateOverflow} */
    /* PV/src/lod_topology.rs:369-401 TerrainLodTopologyError */
    HVX_E_TOPOLOGY_EMPTY = -20,
    HVX_E_TOPOLOGY_DUPLICATE = -21,
    HVX_E_TOPOLOGY_OVERLAP = -22,
    HVX_E_TOPOLOGY_UNBALANCED = -23,
    HVX_E_TOPOLOGY_ROOT_LOD = -24,
    HVX_E_TOPOLOGY_MINIMUM_LOD = -25,
    HVX_E_TOPOLOGY_PAGE_BUDGET = -26,
    HVX_E_TOPOLOGY_MISSING_PARENT = -27,
    HVX_E_TOPOLOGY_COVERAGE = -28,
    /* PV/src/extraction.rs:672-698 ExtractionError */
    HVX_E_INVALID_LIMITS = -40,
    HVX_E_ARITHMETIC_OVERFLOW = -41,
    HVX_E_NON_TRIANGLE_INDEX_COUNT = -42,
    HVX_E_INCOMPLETE_SURFACE_COUNTS = -43,
    HVX_E_PENDING_CAPACITY = -44,       /* PendingCapacity { maximum } */
    HVX_E_ARENA_CAPACITY = -45,         /* ArenaCapacity { capacity } */
    HVX_E_GENERATION_CONFLICT = -46,
    HVX_E_RESERVATION_MISSING = -47,
    HVX_E_RESERVATION_MISMATCH = -48,
    HVX_E_DEVICE_BUFFER_LIMIT = -49     /* DeviceBufferLimit { name, requested, .. } */
} hvx_status;

/* PV/src/extraction.rs:72-79 GpuTerrainVertex (repr(C, align(16)), 32 B). */
typedef struct {
    float position[3];
    uint32_t material;
    float normal[3];
    uint32_t flags;
} hvx_vertex;

/* PV/src/transvoxel_emit.rs:38-48 GpuTransvoxelEmissionCounters (32 B), one per chunk.
 * Capacity overflow is NOT an error: *_overflow = 1, emitted_* = 0, completed = 1 and
 * required_* is the true need (PV/tests/gpu_transvoxel_emission.rs:249-262). */
typedef struct {
    uint32_t required_vertices, required_indices;
    uint32_t emitted_vertices, emitted_indices;
    uint32_t vertex_overflow, index_overflow;
    uint32_t completed, _pad;
} hvx_emission_counters;

/* PV/src/transvoxel_gpu.rs:134-141 GpuTransvoxelClassifyCounters (16 B), one per chunk. */
typedef struct {
    uint32_t visited_cells, active_cells, vertices, triangles;
} hvx_classify_counters;

/* PV/src/transvoxel_transition_gpu.rs:133-146 GpuTransvoxelTransitionCounters (48 B). */
typedef struct {
    uint32_t active_cells, active_faces;
    uint32_t required_vertices, required_indices;
    uint32_t emitted_vertices, emitted_indices;
    uint32_t vertex_overflow, index_overflow;
    uint32_t completed, _pad[3];
} hvx_transition_counters;

/* PV/src/transvoxel_gpu.rs:77-84 GpuTransvoxelCell and
 * PV/src/transvoxel_transition_gpu.rs:67-74 GpuTransvoxelTransitionCell (16 B).
 *   regular:    case | class<<8 | nv<<16 | nt<<24 | 1<<31
 *   transition: case | class_code<<9 | nv<<17 | nt<<21 | 1<<31 */
typedef struct {
    uint32_t packed_case_class_counts, generation_low, generation_high, _pad;
} hvx_cell_record;

/* PV/src/transvoxel_emit.rs:14-21 GpuTransvoxelCellOffset: offsets RELATIVE to the cell's
 * 256-cell scan block; add hvx_scan_block.first_* for the chunk-local value. */
typedef struct {
    uint32_t first_vertex, first_index, generation_low, generation_high;
} hvx_cell_offset;

/* PV/src/transvoxel_emit.rs:29-36 GpuTransvoxelScanBlock (one per 256 cells). */
typedef struct {
    uint32_t vertex_count, index_count, first_vertex, first_index;
} hvx_scan_block;

/* PV/src/extraction.rs:81-92 GpuTerrainMeshlet (32 B) and PV/src/terrain_meshlet.rs:10-20
 * GpuTerrainMeshletBounds (48 B): fixed 63-index meshlets over a chunk's mesh. */
typedef struct {
    uint32_t first_index, index_count, first_vertex, vertex_count;
    uint32_t bounds_offset, generation_low, generation_high, _pad; /* _pad: 0 regular, 1 transition */
} hvx_meshlet;

typedef struct {
    float center[3], radius;
    float cone_apex[3], cone_cutoff;
    float cone_axis[3], _pad;
} hvx_meshlet_bounds;

#define HVX_MESHLET_INDICES 63u /* TERRAIN_MESHLET_BUILD_INDICES, PV/src/terrain_meshlet.rs:7-8 */

/* Where chunk k's mesh lives in the ctx arenas (elements, not bytes).  Slots are
 * fixed-stride (k * max_vertices, k * max_indices), like the reference's per-slot
 * banked arenas (PV/src/surface_publish.wgsl:126-220); index values are chunk-local. */
typedef struct {
    uint32_t first_vertex, vertex_count, first_index, index_count;
} hvx_range;

/* Per-chunk dispatch parameters: the variable part of
 * PV/src/transvoxel_gpu.rs:15-32 GpuTransvoxelDispatch /
 * PV/src/transvoxel_transition_gpu.rs:31-42 GpuTransvoxelTransitionDispatch. */
typedef struct {
    uint64_t generation;
    uint64_t dirty_microbricks; /* 4x4x4 microbricks of (edge/4)^3 cells; bit = mx + 4*my + 16*mz */
    uint32_t transition_mask;   /* TransitionFace bits 0..5 (-X,+X,-Y,+Y,-Z,+Z) */
    uint32_t cost_hint;         /* scheduler input, 0 = unknown: a relative cost estimate, e.g. the chunk's vertex
                                   count at its last extraction.  When any chunk of a batch carries a hint, the
                                   start order follows the hints: descending (longest first, so the kernel does not
                                   end on a lone heavy chunk); a few heavy chunks among many light ones are spread
                                   over the first three quarters of the order instead of starting all at once (the
                                   SMs that only stream then use the bandwidth the emitting ones leave).  Slots,
                                   ranges and meshes do not depend on the order. */
    uint32_t flags;             /* HVX_CHUNK_* */
    uint32_t _reserved;
} hvx_chunk_desc;

/* The producer of the samples knows this chunk holds no surface (e.g. an octree node that is all air or all rock:
 * every sample has the same sign).  Its samples are then neither uploaded (host input) nor read (device input): the
 * chunk's slot reports an empty, completed mesh.  For host input its part of the sample array need not be
 * populated.  A flag on a chunk that does cross the surface makes that chunk's mesh silently empty -- the caller's
 * contract, like the dirty mask. */
#define HVX_CHUNK_UNIFORM 1u

/* PV/src/transvoxel_emit.rs:57-85 TransvoxelGpuExtractorConfig +
 * PV/src/transvoxel_transition_gpu.rs:154-181, widened to a batch of chunks. */
typedef struct {
    uint32_t edge;                    /* 32 (the reference's PAGE_EDGE) or 64 */
    uint32_t max_chunks;              /* batch capacity (reference: 1) */
    uint32_t max_vertices;            /* per chunk, regular; reference default 393,216 */
    uint32_t max_indices;             /* per chunk, regular; reference default 491,520 */
    uint32_t max_transition_vertices; /* per chunk; reference default 73,728; 0 = transitions unused */
    uint32_t max_transition_indices;  /* per chunk; reference default 221,184 */
    uint32_t flags;                   /* HVX_CFG_* */
    uint32_t _reserved;
} hvx_config;

#define HVX_CFG_DEBUG_RECORDS 1u    /* also write per-cell records, offsets and scan blocks (two extra launches) */
#define HVX_CFG_FIRST_GENERATION 2u /* diagnostics: extract regular cells with the first-generation kernel (CTA-wide
                                       barriers instead of the decoupled warps); identical output, slower.  The stress
                                       tests use it as an independent second implementation. */

typedef struct hvx_ctx hvx_ctx;

/* ---- lifetime ---------------------------------------------------------------------- */
/* TransvoxelGpuExtractor::new / TransvoxelGpuTransitionExtractor::new
 * (PV/src/transvoxel_emit.rs:114-231, PV/src/transvoxel_transition_gpu.rs:213-364). */
int hvx_create(hvx_ctx** out, int device, const hvx_config* config);
void hvx_destroy(hvx_ctx* ctx);
/* ctx may be NULL: returns the calling thread's last creation error. */
const char* hvx_last_error(const hvx_ctx* ctx);
const char* hvx_status_name(int status);
uint32_t hvx_abi_version(void);
int hvx_get_config(const hvx_ctx* ctx, hvx_config* out);
/* resource_stats() (PV/src/transvoxel_emit.rs:87-91): device bytes currently allocated. */
uint64_t hvx_allocated_bytes(const hvx_ctx* ctx);
/* Use a caller-owned cudaStream_t (e.g. torch's current stream); NULL restores the ctx stream. */
int hvx_set_stream(hvx_ctx* ctx, void* cuda_stream);
void* hvx_get_stream(const hvx_ctx* ctx);
int hvx_synchronize(hvx_ctx* ctx);
/* The kernel hvx_extract_regular launches for this ctx (as ncu / cuobjdump print it, without the namespace), e.g.
 * "regular_extract_decoupled_kernel<Cfg<64,1,6,20>,false>"; `partial` != 0: the instantiation for batches that hold
 * partially dirty chunks.  bench.py copies it into roofline.kernel. */
const char* hvx_regular_kernel_name(const hvx_ctx* ctx, int partial);
/* number of kernels this ctx has launched so far (bench.py's gpu_launches). */
uint64_t hvx_launch_count(const hvx_ctx* ctx);
/* The order in which hvx_extract_regular starts the chunks of a batch that carries cost hints (host-only; for tests and
 * for callers that want to see the schedule): order_out[k] = index of the k-th chunk started.  Descending hint, ties in
 * index order; when a few chunks are heavy (more than four times the median hint, at most a quarter of the batch) they
 * are placed evenly over the first spread_pct per cent of the order instead of all first -- the library uses 75;
 * 0 = plain descending order. */
int hvx_start_order(const uint32_t* cost_hints, uint32_t n, uint32_t spread_pct, uint32_t* order_out);

/* Roofline probes for the regular kernel (bench tools only): 0 = normal, 1 = stream the samples and do nothing else,
 * 2 = stream + sign bits.  Modes 1 and 2 produce no meshes.  | 0x100: never split chunks across CTAs (a dispatch with
 * fewer chunks than the machine has resident CTAs normally walks z-ranges of chunks; output is identical either way).
 * | 0x200: also split the chunks of a thin last wave of a batch of a few waves (measured slower, off by default).
 * | pct << 12 (bits 12..19): share of the start order the heavy chunks of a hinted batch are spread over
 *   (0 = default 75, 255 = plain descending order). */
int hvx_debug_set_mode(hvx_ctx* ctx, uint32_t mode);

/* Device-side proof of an arithmetic shortcut (diagnostics; tests/test_gpu_regular.py): the regular kernel divides
 * d0 / (d0 - d1) without the IEEE slow path.  Checks all 2^32 pairs of i16 densities against the reference formula
 * with __fdiv_rn; *mismatches_out = pairs whose result bits differ (must be 0), *witness_out = (d0+32768)<<16 | (d1+32768)
 * of one of them. */
int hvx_selftest_edge_parameter(int device, uint64_t* mismatches_out, uint32_t* witness_out);
/* Same for the normal's 1 / sqrt(s) (inv_sqrt_rn_normal against IEEE sqrt then divide): every float s from 1e-12 (the
 * kernel's guard) to 2^40 (s is a sum of three squared i16 differences, < 2^35).  *witness_out = bits of one bad s. */
int hvx_selftest_inv_sqrt(int device, uint64_t* mismatches_out, uint32_t* witness_out);

/* ---- K1: density / SDF fill -------------------------------------------------------- */
/* ExtractionFixture::new (PV/src/fixture.rs:95-124) on the device.
 * kind: 0..5 = ExtractionFixtureKind order (plane, sphere, cave, sharp_corner, thin_slab,
 * material_seam); 16 = fBm terrain (helio-pass-sdf noise.rs terrain_sdf, Rolling), 17 = dense
 * random.  page_xyz: host [n][3]; lod: host [n] or NULL (all 0).  d_samples: device,
 * n*(edge+2)^3 words, x fastest; NULL = the ctx sample arena (hvx_buffer(HVX_BUF_SAMPLES)). */
int hvx_fill_density(hvx_ctx* ctx, uint32_t kind, const int64_t* page_xyz, const uint8_t* lod, uint32_t n,
                     uint32_t* d_samples);
/* TransvoxelTransitionFaceFixture::slab_samples for six faces
 * (PV/src/transvoxel_transition.rs:335-349): n * 6*3*(2*edge+3)^2 words; lod[i] >= 1. */
int hvx_fill_slabs(hvx_ctx* ctx, uint32_t kind, const int64_t* page_xyz, const uint8_t* lod, uint32_t n,
                   uint32_t* d_slabs);

/* ---- edits on resident samples (BASELINE config 5: the incremental-edit path) -------------------------- */
/* GpuVoxelEdit (crates/helio-voxel-core/src/gpu_types.rs:47-54), 32 bytes. */
typedef struct {
    uint32_t volume_id;
    uint32_t op_type;     /* VoxelOp (edit.rs:5-10): 1 AddSphere, 2 SubtractSphere; 0 SetBox and 3 Paint are not handled */
    uint32_t material;
    float center[3];      /* metres, in the frame of the pages' LOD0 cell grid (cell c spans [0.1 c, 0.1 (c+1)) m) */
    float radius;         /* metres */
    uint32_t _pad;
} hvx_voxel_edit;

/* Apply one sphere edit to the device-resident samples of chunks [0, n) (d_samples NULL = the ctx sample arena,
 * page_xyz / lod as in hvx_fill_density) and derive what must be re-extracted.
 *   samples:  every sample within `radius` of the centre -- halo samples included, so neighbouring chunks stay
 *             consistent -- gets  carve = clamp_i16(rint((radius - distance) / cell_m * 256)),
 *             SubtractSphere: density = max(density, carve), material 0 where the sample becomes air;
 *             AddSphere: density = min(density, -carve), material = edit.material where it becomes solid.
 *             (The reference queues edits and marks its octree, crates/helio/src/scene/voxel.rs:81-115, but has no
 *             kernel that applies them to planetary pages; this rule is ours, restated in oracle/edit.py.)
 *   dirty_out[n]: per chunk, the microbricks to re-extract: the legacy octree's sphere-vs-box rule
 *             |centre_a - box_centre_a| > half_a + r  =>  untouched  (crates/helio-voxel-core/src/octree.rs:139-173)
 *             at microbrick granularity, with r = radius + 2 cells (a vertex also reads the gradient neighbours of
 *             its cell's corners).  0 = chunk untouched.  Feed it to hvx_chunk_desc.dirty_microbricks.
 *   *touched_out: chunks whose samples were modified. */
int hvx_apply_edit(hvx_ctx* ctx, const hvx_voxel_edit* edit, const int64_t* page_xyz, const uint8_t* lod, uint32_t n,
                   uint32_t* d_samples, uint64_t* dirty_out, uint32_t* touched_out);

/* ---- K2-K4: regular cells ---------------------------------------------------------- */
/* TransvoxelGpuExtractor::dispatch (PV/src/transvoxel_emit.rs:233-254) over n chunks.
 * samples: HOST or DEVICE pointer (detected), n*(edge+2)^3 CellWords, 16-byte aligned;
 * NULL = the ctx sample arena.  sample_words is the caller's slice length and must equal
 * n*(edge+2)^3 (else HVX_E_SAMPLE_COUNT).  descs: host [n]. */
int hvx_extract_regular(hvx_ctx* ctx, const uint32_t* samples, uint64_t sample_words,
                        const hvx_chunk_desc* descs, uint32_t n);
/* TransvoxelGpuClassifier::dispatch (PV/src/transvoxel_gpu.rs:268-286): classification +
 * classify counters (+ cell records when HVX_CFG_DEBUG_RECORDS), no emission. */
int hvx_classify_regular(hvx_ctx* ctx, const uint32_t* samples, uint64_t sample_words,
                         const hvx_chunk_desc* descs, uint32_t n);
/* dispatch + the read-back the reference's callers do afterwards (map counters, copy the emitted ranges;
 * PV/tests/gpu_transvoxel_emission.rs:59-101), as ONE pipelined call: host samples are uploaded in sub-batches of
 * ~256 MiB on a copy stream while the previous sub-batch is extracted, and every sub-batch's mesh is packed and copied
 * to the host behind the kernels on a third stream.  Outputs as hvx_read_meshes (tightly packed in chunk order,
 * indices chunk-local, ranges_out[n] = packed placement) plus counters_out[n] (nullable).  Returns
 * HVX_E_INVALID_CAPACITY if the host arrays are too small (totals and ranges_out are still filled). */
int hvx_extract_regular_to_host(hvx_ctx* ctx, const uint32_t* samples, uint64_t sample_words,
                                const hvx_chunk_desc* descs, uint32_t n, hvx_vertex* vertices_out, uint64_t vertex_cap,
                                uint32_t* indices_out, uint64_t index_cap, hvx_range* ranges_out,
                                hvx_emission_counters* counters_out, uint64_t* total_vertices, uint64_t* total_indices);

/* ---- K2-K4 (T): transition cells ---------------------------------------------------- */
/* TransvoxelGpuTransitionExtractor::dispatch (PV/src/transvoxel_transition_gpu.rs:366-380).
 * slabs: HOST or DEVICE, n * 6*3*(2*edge+3)^2 words (face-major, then layer, v, u).
 * descs[i].transition_mask selects faces; bits above 0x3f -> HVX_E_TRANSITION_MASK. */
int hvx_extract_transition(hvx_ctx* ctx, const uint32_t* slabs, uint64_t slab_words,
                           const hvx_chunk_desc* descs, uint32_t n);

/* ---- meshlets (the step after extraction in the reference's pass) -------------------- */
/* build_regular / build_transition (PV/src/terrain_meshlet_build.wgsl:205-261; CPU twin
 * build_terrain_meshlets, PV/src/terrain_meshlet.rs:83-155) for chunks [0, n) of the last
 * extraction of that kind: fixed 63-index partition, unique-vertex count, AABB-centre bounding
 * sphere, normal cone.  Chunk k's meshlets start at k * ceil(max_indices / 63); chunks that
 * overflowed publish none.  kind 0 = regular, 1 = transition. */
int hvx_build_meshlets(hvx_ctx* ctx, int kind, uint32_t n);

/* ---- optional vertex-reuse output (north_star kernel 4) -------------------------------------------- */
/* The reference never shares vertices between cells: the reuse byte of the Transvoxel vertex code is not consumed
 * (PV/src/transvoxel_emit.wgsl:322-358 uses `code & 0xff` only; PV/src/transvoxel.rs:99-101 reuse() has no caller), so
 * the default output repeats every crossing edge once per cell that touches it, and parity demands exactly that.
 * hvx_weld_meshes rewrites the meshes of chunks [0, n) of the last extraction of that kind IN PLACE into the indexed
 * mesh edge-ownership reuse would give: bit-identical 32-byte vertex records (the copies of one edge) are merged, the
 * first occurrence is kept and the original order preserved; every index is redirected to the kept copy (triangle
 * order and winding unchanged); range.vertex_count and counters.emitted_vertices become the kept count
 * (required_vertices still reports what the unshared mesh needs).  Idempotent.  Off by default, no reference
 * counterpart: the oracle is oracle/weld.py.  kind 0 = regular, 1 = transition. */
int hvx_weld_meshes(hvx_ctx* ctx, int kind, uint32_t n);

/* ---- surface gather (the step before extraction in the reference's pass) ---------- */
/* GpuPageTableEntry (PV/src/table.rs:8-19): open-addressed page table of the residency atlas. */
typedef struct {
    uint32_t planet_id[4];
    int32_t relative_lod0_cell_min[3];
    uint32_t lod;
    uint32_t slot;
    uint32_t generation_low, generation_high;
    uint32_t state;                    /* 0 empty, 1 occupied, 2 tombstone */
} hvx_page_table_entry;

/* GpuResidencyUniform (PV/src/table.rs:62-72). */
typedef struct {
    uint32_t table_mask, max_probe, resident_pages;
    uint32_t atlas_tiles_x, atlas_tiles_y, atlas_tiles_z;
    uint32_t publication_epoch_low, publication_epoch_high;
} hvx_residency;

/* GpuSurfaceGatherJob (PV/src/surface_sampling.rs:121-134). */
typedef struct {
    uint32_t planet_id[4];
    int32_t relative_lod0_cell_min[3];
    uint32_t lod;
    uint32_t generation_low, generation_high;
    uint32_t transition_mask;
    uint32_t target_slot;
    uint32_t residency_epoch_low, residency_epoch_high;
    uint32_t _pad[2];
} hvx_gather_job;

/* GpuSurfaceGatherCounters (PV/src/surface_sampling.rs:171-181). */
typedef struct {
    uint32_t regular_samples, transition_samples, table_probes, page_misses;
    uint32_t stale_targets, completed, _pad[2];
} hvx_gather_counters;

/* GpuSurfaceSampler::prepare + encode (PV/src/surface_sampling.rs:309-337; gather_regular,
 * gather_transition, finalize_gather in PV/src/surface_gather.wgsl:200-264) for n jobs at once.
 * Requires edge 32 (the residency page edge).  For job k it builds the 34^3 halo block into
 * HVX_BUF_SAMPLES[k] and, for the faces in transition_mask (lod > 0), the six 67x67x3 fine-side slabs
 * into HVX_BUF_SLABS[k] from the resident page atlas, looking pages up in the open-addressed table
 * exactly like the reference (same hash, same probe sequence, missing page -> AIR 0x00007fff).
 * Counters have the reference's values (table_probes counts one lookup per gathered sample);
 * HVX_BUF_GATHER_INDIRECT[k] holds the eight DispatchIndirectArgs finalize_gather publishes.
 * A job whose residency epoch does not match the uniform gathers nothing, like the shader.
 *   table:  HOST or DEVICE, table_mask + 1 entries; NULL = the table bound with hvx_gather_bind_table (a host table is
 *           otherwise staged on every call).
 *   atlas:  HOST or DEVICE, R32Uint texels of the 3-D atlas in linear order, x fastest:
 *           (32 tiles_x) x (32 tiles_y) x (32 tiles_z) words; page slot s occupies the 32^3 tile at
 *           tile coordinates (s % tiles_x, (s / tiles_x) % tiles_y, s / (tiles_x tiles_y)).
 *   jobs:   HOST, n entries.
 * The gathered arenas are then extracted with hvx_extract_regular / hvx_extract_transition
 * (samples = NULL), with no host round trip of the 480 KB per page. */
int hvx_gather_surface(hvx_ctx* ctx, const hvx_residency* residency, const hvx_page_table_entry* table,
                       const uint32_t* atlas, uint64_t atlas_words, const hvx_gather_job* jobs, uint32_t n);
/* Keep the open-addressed page table resident on the ctx's device: copied once (HOST or DEVICE source, `entries` a power
 * of two), used by every later hvx_gather_surface that passes table = NULL, until the next bind.  The residency layer
 * republishes its table when the publication epoch advances (PV/src/table.rs:62-72): bind then, not per dispatch.
 * table = NULL or entries = 0 unbinds. */
int hvx_gather_bind_table(hvx_ctx* ctx, const hvx_page_table_entry* table, uint32_t entries);

/* ---- surface publication (the step after extraction in the reference's pass) -------- */
/* GpuSurfaceJob (PV/src/render.rs:466-480) */
typedef struct {
    uint32_t slot, transition_mask, generation_low, generation_high;
    uint32_t regular_max_vertices, regular_max_indices, transition_max_vertices, transition_max_indices;
    uint32_t regular_max_meshlets, transition_max_meshlets, _pad[2];
} hvx_surface_job;
/* GpuPageMeta (crates/helio-planet-voxel-core/src/gpu.rs:62-99) */
typedef struct {
    int32_t relative_lod0_cell_min[3];
    uint32_t lod, slot, generation_low, generation_high, transition_mask;
} hvx_page_meta;
/* GpuSurfaceState (PV/src/render.rs:505-519) */
typedef struct {
    uint32_t generation_low, generation_high, active_bank, valid;
    uint32_t regular_vertex_count, regular_index_count, transition_vertex_count, transition_index_count;
    uint32_t regular_meshlet_count, transition_meshlet_count, _pad[2];
} hvx_surface_state;
/* GpuDrawPage (PV/src/render.rs:533-544) */
typedef struct {
    int32_t relative_lod0_cell_min[3];
    uint32_t lod;
    float camera_relative_m[3];
    float lod0_cell_size_m;
    uint32_t generation_low, generation_high, transition_mask, visible;
} hvx_draw_page;
/* GpuSurfaceFeedback (PV/src/render.rs:521-531) */
typedef struct {
    uint32_t submitted_jobs, published_jobs, stale_rejections, overflow_rejections, incomplete_rejections, _pad[3];
} hvx_surface_feedback;
/* DrawIndexedIndirectArgs (PV/src/render.rs:546-554) */
typedef struct {
    uint32_t index_count, instance_count, first_index;
    int32_t base_vertex;
    uint32_t first_instance;
} hvx_draw_indexed_indirect;

/* The double-banked per-slot arenas, surface states, indirect draws and feedback the reference's render
 * pass owns (PV/src/render.rs), for `slots` residency slots.  Bank capacities are the ctx's per-chunk
 * capacities: bank b of the regular arena starts at b * max_vertices / b * max_indices, b = slot*2 + bank. */
typedef struct hvx_publisher hvx_publisher;
int hvx_publisher_create(hvx_ctx* ctx, uint32_t slots, hvx_publisher** out);
void hvx_publisher_destroy(hvx_publisher* pub);

/* copy_regular_surface + copy_transition_surface + publish_surface (PV/src/surface_publish.wgsl:125-216)
 * for n jobs of the ctx's last regular (and, if the ctx has transition capacity, transition) extraction.
 * job_chunk[i] names the extraction chunk that holds job i's meshes and counters.  Per job, exactly like
 * the shader: a job whose page metadata moved on (slot / generation mismatch) is a stale rejection; an
 * incomplete or overflowed extraction is rejected; otherwise the meshes are copied into the slot's
 * INACTIVE bank (only emitted_vertices / emitted_indices elements, not a capacity-sized dispatch), the
 * state flips to that bank with the new generation and counts, and both DrawIndexedIndirectArgs are
 * rewritten with instance_count 0 (visibility is granted by hvx_refresh_visibility).  A ctx without
 * transition capacity publishes empty, completed transition counters (what the reference's transition
 * extractor reports for mask 0).  Slots must be distinct within one call (the reference handles one job
 * per submission; two jobs for one slot in one batch would race on the bank).
 *   jobs, job_chunk: HOST, n entries.   page_metadata: HOST or DEVICE, `slots` entries. */
int hvx_publish_surfaces(hvx_publisher* pub, const hvx_surface_job* jobs, const uint32_t* job_chunk,
                         const hvx_page_meta* page_metadata, uint32_t n);
/* refresh_visibility (PV/src/surface_publish.wgsl:218-225): instance_count = valid && visible, per slot.
 *   draw_pages: HOST or DEVICE, `slots` entries. */
int hvx_refresh_visibility(hvx_publisher* pub, const hvx_draw_page* draw_pages);

typedef enum {
    HVX_PUB_REGULAR_VERTICES = 0,     /* hvx_vertex [slots*2][max_vertices] */
    HVX_PUB_REGULAR_INDICES = 1,      /* u32 [slots*2][max_indices] */
    HVX_PUB_TRANSITION_VERTICES = 2,  /* hvx_vertex [slots*2][max_transition_vertices] (if enabled) */
    HVX_PUB_TRANSITION_INDICES = 3,
    HVX_PUB_STATES = 4,               /* hvx_surface_state [slots] */
    HVX_PUB_REGULAR_DRAWS = 5,        /* hvx_draw_indexed_indirect [slots] */
    HVX_PUB_TRANSITION_DRAWS = 6,
    HVX_PUB_FEEDBACK = 7,             /* hvx_surface_feedback [1] */
    HVX_PUB_COUNT = 8
} hvx_publisher_buffer_id;
void* hvx_publisher_buffer(hvx_publisher* pub, int buffer_id);
uint64_t hvx_publisher_buffer_bytes(hvx_publisher* pub, int buffer_id);
/* synchronising copies (the states / draws / feedback are caller-visible render state) */
int hvx_publisher_read(hvx_publisher* pub, int buffer_id, uint64_t byte_offset, uint64_t bytes, void* dst);
int hvx_publisher_write(hvx_publisher* pub, int buffer_id, uint64_t byte_offset, uint64_t bytes, const void* src);

/* ---- bounded extraction publisher: variable-size arena placement ------------------ */
/* The reference's generation-safe bounded publication contract, PV/src/extraction.rs:8-704
 * (GpuExtractionRequest, GpuExtractionRange, GpuExtractionCounters, ExtractionLimits,
 * BoundedExtractionPublisher, first-fit RangeAllocator).  Host-side bookkeeping is plain C++; the one
 * device step (hvx_extraction_commit) moves a reserved page's mesh out of its fixed-stride extraction
 * slot into the bounded, tightly packed arenas at the reserved ranges. */
typedef struct {            /* GpuExtractionRequest, PV/src/extraction.rs:10-21 (32 B) */
    uint32_t page_slot, generation_low, generation_high, transition_mask;
    uint32_t dirty_microbricks_low, dirty_microbricks_high, _pad[2];
} hvx_extraction_request;
typedef struct {            /* GpuExtractionRange, PV/src/extraction.rs:54-65 (32 B) */
    uint32_t first_vertex, vertex_count, first_index, index_count;
    uint32_t first_meshlet, meshlet_count, generation_low, generation_high;
} hvx_extraction_range;
typedef struct {            /* GpuExtractionCounters, PV/src/extraction.rs:94-109 (48 B) */
    uint32_t requests, active_cells, vertices, indices, meshlets, completed, stale_rejected, overflowed;
    uint32_t vertex_overflow, index_overflow, meshlet_overflow, _pad;
} hvx_extraction_counters;
typedef struct {            /* ExtractionLimits, PV/src/extraction.rs:111-118 */
    uint32_t max_page_slots, max_pending_pages, max_vertices, max_indices, max_meshlets;
} hvx_extraction_limits;
typedef struct {            /* ExtractionAllocationPlan, PV/src/extraction.rs:212-221 */
    uint64_t request_bytes, page_range_bytes, vertex_bytes, index_bytes, meshlet_bytes, counter_bytes, total_bytes;
} hvx_extraction_plan;
typedef struct {            /* PlanetPageKey, helio-planet-voxel-core/src/types.rs:312-316 */
    uint8_t planet_id[16];
    int64_t page_xyz[3];
    uint8_t lod, _pad[7];
} hvx_planet_page_key;
typedef struct { uint32_t vertices, indices, meshlets; } hvx_surface_counts;       /* SurfaceCounts :223-228 */
typedef struct { uint32_t first, count; } hvx_arena_slice;                         /* ArenaSlice :246-250 */
typedef struct { hvx_arena_slice vertices, indices, meshlets; } hvx_surface_allocation; /* :252-257 */
typedef struct {            /* ExtractionReservation :283-288 */
    hvx_planet_page_key key;
    uint64_t generation;
    hvx_surface_allocation allocation;
} hvx_reservation;
typedef struct {            /* PublishedSurface :290-294 */
    uint64_t generation;
    hvx_surface_allocation allocation;
} hvx_published_surface;
typedef enum {              /* ReservationOutcome :296-302 */
    HVX_RESERVED = 0, HVX_RESERVE_CURRENT = 1, HVX_RESERVE_DUPLICATE_PENDING = 2, HVX_RESERVE_STALE = 3
} hvx_reservation_kind;
typedef struct {
    uint32_t kind;                       /* hvx_reservation_kind */
    uint32_t detail;                     /* on HVX_E_ARENA_CAPACITY: 0 vertices, 1 indices, 2 meshlets; on
                                            HVX_E_PENDING_CAPACITY: the maximum */
    hvx_reservation reservation;         /* RESERVED, DUPLICATE_PENDING */
    hvx_published_surface current;       /* CURRENT */
    uint64_t newest_generation;          /* STALE */
} hvx_reservation_outcome;
typedef struct {            /* PublicationOutcome :304-313; kind 0 Published, 1 Stale */
    uint32_t kind, has_replaced;
    hvx_published_surface current, replaced;
    uint64_t newest_generation;
} hvx_publication_outcome;
typedef struct {            /* ExtractionEvictOutcome :315-320; kind 0 Evicted, 1 Missing, 2 Stale */
    uint32_t kind, _pad;
    uint64_t newest_generation;
} hvx_evict_outcome;
typedef struct {            /* ExtractionPublisherCounters :322-340 */
    uint64_t current_pages, pending_pages;
    uint32_t used_vertices, used_indices, used_meshlets, _pad0;
    uint64_t pending_high_water;
    uint32_t vertex_high_water, index_high_water, meshlet_high_water, _pad1;
    uint64_t reservations, publications, replacements, cancellations, evictions, stale_rejected, backpressured;
} hvx_extraction_publisher_counters;

/* GpuExtractionRequest::new (:24-42): HVX_E_TRANSITION_MASK for bits outside the six faces. */
int hvx_extraction_request_new(uint32_t page_slot, uint64_t generation, uint32_t transition_mask,
                               uint64_t dirty_microbricks, hvx_extraction_request* out);
/* ExtractionLimits::new + allocation_plan (:121-178): HVX_E_INVALID_LIMITS / HVX_E_ARITHMETIC_OVERFLOW. */
int hvx_extraction_limits_plan(const hvx_extraction_limits* limits, hvx_extraction_plan* plan_out);
/* ExtractionLimits::validate_device (:180-203) against wgpu-style limits: HVX_E_DEVICE_BUFFER_LIMIT with
 * *name_out = the first buffer that does not fit ("terrain vertices", ...) and *requested_out its bytes. */
int hvx_extraction_limits_validate_device(const hvx_extraction_limits* limits, uint64_t max_buffer_size,
                                          uint64_t max_storage_buffer_binding_size, const char** name_out,
                                          uint64_t* requested_out);
/* SurfaceAllocation::gpu_range (:268-280). */
void hvx_extraction_gpu_range(const hvx_surface_allocation* allocation, uint64_t generation, hvx_extraction_range* out);

typedef struct hvx_extraction_publisher hvx_extraction_publisher;
/* BoundedExtractionPublisher::new (:357-367); limits are validated like ExtractionLimits::new. */
int hvx_extraction_publisher_create(const hvx_extraction_limits* limits, hvx_extraction_publisher** out);
void hvx_extraction_publisher_destroy(hvx_extraction_publisher* pub);
/* reserve (:398-479), publish (:481-518), cancel_pending (:520-539), evict (:541-566). */
int hvx_extraction_reserve(hvx_extraction_publisher* pub, const hvx_planet_page_key* key, uint64_t generation,
                           const hvx_surface_counts* counts, hvx_reservation_outcome* out);
int hvx_extraction_publish(hvx_extraction_publisher* pub, const hvx_reservation* reservation,
                           hvx_publication_outcome* out);
int hvx_extraction_cancel_pending(hvx_extraction_publisher* pub, const hvx_planet_page_key* key, uint64_t generation,
                                  int* cancelled_out);
int hvx_extraction_evict(hvx_extraction_publisher* pub, const hvx_planet_page_key* key, uint64_t generation,
                         hvx_evict_outcome* out);
/* current / pending (:373-379): return 1 and fill *out if present, 0 if not. */
int hvx_extraction_current(const hvx_extraction_publisher* pub, const hvx_planet_page_key* key,
                           hvx_published_surface* out);
int hvx_extraction_pending(const hvx_extraction_publisher* pub, const hvx_planet_page_key* key, hvx_reservation* out);
/* counters (:381-396). */
int hvx_extraction_publisher_get_counters(const hvx_extraction_publisher* pub, hvx_extraction_publisher_counters* out);

/* Device side of the contract.  attach: allocate the plan's bounded arenas on the ctx's device
 * (vertices, indices, page ranges, counters).  commit: for n reserved pages, copy the REGULAR mesh of
 * extraction chunk chunk[i] (ctx's last hvx_extract_regular) to the reserved vertex / index ranges and
 * write page_ranges[page_slot[i]] = gpu_range(generation); one launch for the whole batch, 16 bytes per
 * thread.  A reservation whose counts differ from what the chunk emitted (or a chunk that overflowed) is
 * not copied: it bumps counters.overflowed and leaves the page range untouched.  The caller then calls
 * hvx_extraction_publish (host) -- the old ranges are recycled only after that, like the reference. */
int hvx_extraction_publisher_attach(hvx_extraction_publisher* pub, hvx_ctx* ctx);
int hvx_extraction_commit(hvx_extraction_publisher* pub, const uint32_t* chunk, const uint32_t* page_slot,
                          const hvx_reservation* reservations, uint32_t n);
typedef enum {
    HVX_XPUB_VERTICES = 0,    /* hvx_vertex [max_vertices] */
    HVX_XPUB_INDICES = 1,     /* u32 [max_indices] */
    HVX_XPUB_PAGE_RANGES = 2, /* hvx_extraction_range [max_page_slots] */
    HVX_XPUB_COUNTERS = 3,    /* hvx_extraction_counters [1] */
    HVX_XPUB_COUNT = 4
} hvx_extraction_publisher_buffer_id;
void* hvx_extraction_publisher_buffer(hvx_extraction_publisher* pub, int buffer_id);
int hvx_extraction_publisher_read(hvx_extraction_publisher* pub, int buffer_id, uint64_t byte_offset, uint64_t bytes,
                                  void* dst);

/* ---- legacy 8^3-brick marching cubes (helio-voxel-core bricks) ------------------------ */
/* The extraction half of VoxelMeshPass (crates/passes/3d/helio-pass-voxel-mesh/src/lib.rs:153-726,
 * shaders/voxel_surface_extract.wgsl) for a batch of dirty bricks.  PODs as in the reference: */
typedef struct { uint32_t data_offset, occupancy; } hvx_brick_meta;          /* GpuBrickMeta, helio-voxel-core/src/gpu_types.rs:18-22;
                                                                                data_offset in WORDS into voxel_data */
typedef struct {            /* DirtyBrick, helio-pass-voxel-mesh/src/lib.rs:64-69 (32 B) */
    uint32_t brick_slot, volume_id, _pad[2];
    float origin_size[4];   /* xyz = world origin, w = voxel size */
} hvx_dirty_brick;
typedef struct {            /* GpuBrickMeshlet, helio-voxel-core/src/gpu_types.rs:25-33 (32 B) */
    uint32_t vertex_offset, index_offset, vertex_count, index_count, brick_index, volume_id, _pad[2];
} hvx_brick_meshlet;
#define HVX_BRICK_VOXEL_WORDS 183u   /* padded 9x9x9 bytes, VOXEL_MESH_BRICK_VOXEL_WORDS */
#define HVX_BRICK_MAX_ENTRIES 2048u  /* MAX_SURFACE_VERTS_PER_BRICK == MAX_SURFACE_INDICES_PER_BRICK */

/* Owns vertex_buf / normal_buf (float4 [max_bricks][2048]), index_buf (u32 [max_bricks][2048]), descriptors and
 * indirect draws (per brick slot), like VoxelMeshPass::new (reference capacity: 1,024 bricks). */
typedef struct hvx_brick_mesher hvx_brick_mesher;
int hvx_brick_mesher_create(hvx_ctx* ctx, uint32_t max_bricks, hvx_brick_mesher** out);
void hvx_brick_mesher_destroy(hvx_brick_mesher* mesher);
/* The compute step of VoxelMeshPass::execute (lib.rs:653-667) for n_dirty bricks: marching cubes over each brick's
 * padded 9^3 block -> edge-midpoint vertices (xyz, material), central-difference normals, one index per vertex, the
 * brick's GpuBrickMeshlet descriptor and DrawIndexedIndirect.  Within a brick, cells are emitted in linear order
 * (the reference's order depends on its atomics; this is one of its possible outcomes).  A cell that would pass
 * 2,048 entries is dropped, counts are clamped -- the reference's overflow rule.
 *   meta:   n_meta entries (indexed by brick slot);  dirty: n_dirty entries -- both HOST or both DEVICE;
 *   voxels: HOST or DEVICE, n_words words.
 * Host lists: HVX_E_BATCH_CAPACITY if a slot >= max_bricks or >= n_meta, HVX_E_SAMPLE_COUNT if a brick's 183
 * words do not lie inside voxels, before anything is launched.  Device lists: the kernel makes the same checks,
 * skips the entry and counts it in HVX_BRICK_REJECTED. */
int hvx_brick_extract(hvx_brick_mesher* mesher, const hvx_brick_meta* meta, uint32_t n_meta, const uint32_t* voxels,
                      uint64_t n_words, const hvx_dirty_brick* dirty, uint32_t n_dirty);
/* clear_brick_slot (lib.rs:590-615): zero the slot's indirect draw. */
int hvx_brick_clear_slot(hvx_brick_mesher* mesher, uint32_t brick_slot);
typedef enum {
    HVX_BRICK_VERTICES = 0,     /* float4 [max_bricks][2048]: xyz world position, w = material */
    HVX_BRICK_NORMALS = 1,      /* float4 [max_bricks][2048] */
    HVX_BRICK_INDICES = 2,      /* u32 [max_bricks][2048], brick-local */
    HVX_BRICK_DESCRIPTORS = 3,  /* hvx_brick_meshlet [max_bricks] */
    HVX_BRICK_DRAWS = 4,        /* hvx_draw_indexed_indirect [max_bricks] */
    HVX_BRICK_REJECTED = 5,     /* u32 [1]: dirty entries of DEVICE-resident lists skipped by the kernel's bounds check */
    HVX_BRICK_BUF_COUNT = 6
} hvx_brick_buffer_id;
void* hvx_brick_buffer(hvx_brick_mesher* mesher, int buffer_id);
uint64_t hvx_brick_buffer_bytes(hvx_brick_mesher* mesher, int buffer_id);
int hvx_brick_read(hvx_brick_mesher* mesher, int buffer_id, uint64_t byte_offset, uint64_t bytes, void* dst);

/* ---- outputs ------------------------------------------------------------------------ */
typedef enum {
    HVX_BUF_SAMPLES = 0,            /* u32  [max_chunks][(edge+2)^3]          (lazy) */
    HVX_BUF_SLABS = 1,               /* u32  [max_chunks][6*3*(2*edge+3)^2]    (lazy) */
    HVX_BUF_REGULAR_VERTICES = 2,    /* hvx_vertex [max_chunks][max_vertices]            vertices_buffer() */
    HVX_BUF_REGULAR_INDICES = 3,     /* u32  [max_chunks][max_indices]                   indices_buffer()  */
    HVX_BUF_REGULAR_COUNTERS = 4,    /* hvx_emission_counters [max_chunks]               counters_buffer() */
    HVX_BUF_REGULAR_CLASSIFY = 5,    /* hvx_classify_counters [max_chunks]  classifier.counters_buffer() */
    HVX_BUF_REGULAR_RANGES = 6,      /* hvx_range [max_chunks] */
    HVX_BUF_REGULAR_CELLS = 7,       /* hvx_cell_record [max_chunks][edge^3]   (debug)  output_buffer()  */
    HVX_BUF_REGULAR_OFFSETS = 8,     /* hvx_cell_offset [max_chunks][edge^3]   (debug)  offsets_buffer() */
    HVX_BUF_REGULAR_BLOCKS = 9,      /* hvx_scan_block  [max_chunks][edge^3/256] (debug) blocks_buffer() */
    HVX_BUF_TRANSITION_VERTICES = 10,
    HVX_BUF_TRANSITION_INDICES = 11,
    HVX_BUF_TRANSITION_COUNTERS = 12, /* hvx_transition_counters [max_chunks] */
    HVX_BUF_TRANSITION_RANGES = 13,
    HVX_BUF_TRANSITION_CELLS = 14,    /* [max_chunks][6*edge^2] (debug) cells_buffer() */
    HVX_BUF_TRANSITION_OFFSETS = 15,
    HVX_BUF_TRANSITION_BLOCKS = 16,   /* [max_chunks][6*edge^2/256] */
    HVX_BUF_REGULAR_MESHLETS = 17,         /* hvx_meshlet        [max_chunks][ceil(max_indices/63)]  (lazy) */
    HVX_BUF_REGULAR_MESHLET_BOUNDS = 18,   /* hvx_meshlet_bounds [max_chunks][ceil(max_indices/63)]  (lazy) */
    HVX_BUF_REGULAR_MESHLET_COUNTS = 19,   /* u32 [max_chunks] */
    HVX_BUF_TRANSITION_MESHLETS = 20,
    HVX_BUF_TRANSITION_MESHLET_BOUNDS = 21,
    HVX_BUF_TRANSITION_MESHLET_COUNTS = 22,
    HVX_BUF_GATHER_COUNTERS = 23,          /* hvx_gather_counters [max_chunks]                 (lazy) */
    HVX_BUF_GATHER_INDIRECT = 24,          /* u32 [max_chunks][8][3]  DispatchIndirectArgs     (lazy) */
    HVX_BUF_COUNT = 25
} hvx_buffer_id;

/* Device pointer / size of a ctx arena (allocating it if lazy); NULL / 0 if unavailable. */
void* hvx_buffer(hvx_ctx* ctx, int buffer_id);
uint64_t hvx_buffer_bytes(hvx_ctx* ctx, int buffer_id);
/* Synchronising device->host copy of [byte_offset, byte_offset + bytes) of an arena. */
int hvx_read(hvx_ctx* ctx, int buffer_id, uint64_t byte_offset, uint64_t bytes, void* host_dst);
/* Host->device copy into an arena (samples / slabs upload for device-resident workflows). */
int hvx_write(hvx_ctx* ctx, int buffer_id, uint64_t byte_offset, uint64_t bytes, const void* host_src);
/* Gather the emitted meshes of chunks [first, first+n) into tightly packed host arrays in chunk
 * order (vertices then indices, index values stay chunk-local).  kind 0 = regular, 1 = transition.
 * ranges_out[n] receives the packed placement.  *_cap in elements; returns HVX_E_INVALID_CAPACITY
 * if too small (ranges_out is still filled so the caller can size the buffers). */
int hvx_read_meshes(hvx_ctx* ctx, int kind, uint32_t first, uint32_t n, hvx_vertex* vertices_out,
                    uint64_t vertex_cap, uint32_t* indices_out, uint64_t index_cap, hvx_range* ranges_out,
                    uint64_t* total_vertices, uint64_t* total_indices);

/* ---- host-side LOD scheduler input (no GPU involved) --------------------------------- */
/* PV/src/lod_topology.rs.  hvx_page is PageKey + the derived coarse-owned transition mask. */
typedef struct {
    int64_t page_xyz[3];
    uint8_t lod;
    uint8_t transition_mask;
    uint8_t _pad[6];
} hvx_page;

typedef struct {
    uint32_t pages, minimum_lod, maximum_lod, transition_faces;
} hvx_lod_stats;

/* TerrainLodTopology::new (PV/src/lod_topology.rs:27-100): sorts pages into PageKey order
 * (lod, then xyz), validates and fills transition_mask. */
int hvx_lod_topology(hvx_page* pages, uint32_t n, uint32_t edge, hvx_lod_stats* stats);
/* HorizonLodFixturePlan::build_with_minimum_lod (PV/src/lod_topology.rs:169-217).
 * out[max_pages]; *n_out pages written in PageKey order with masks; root_out optional. */
int hvx_horizon_plan(const int64_t focus_lod0_cell[3], uint32_t root_lod, uint32_t minimum_lod,
                     uint32_t max_pages, uint32_t edge, hvx_page* out, uint32_t* n_out, hvx_page* root_out,
                     hvx_lod_stats* stats);
/* The interleave step of the optional multi-GPU mesh gather (helio_b200/distributed.py): n independent
 * device-to-device segment copies in ONE launch on `cuda_stream` of `device`.  segments: HOST, n triples
 * (source word, destination word, words) of 32-bit words into src / dst.  Segments must not overlap in dst. */
int hvx_copy_segments(int device, void* cuda_stream, const uint32_t* d_src, uint32_t* d_dst, const uint64_t* segments,
                      uint32_t n);
/* Static multi-GPU partition: LPT-greedy assignment of chunks to ranks by cost;
 * owner[i] in [0, ranks).  Deterministic (ties -> lower chunk index, lower rank). */
int hvx_partition_chunks(const uint64_t* cost, uint32_t n, uint32_t ranks, uint32_t* owner);
/* Cost model used by the scheduler: bytes a chunk moves (samples + slabs of masked faces). */
uint64_t hvx_chunk_cost(uint32_t edge, uint32_t transition_mask);

#ifdef __cplusplus
}
#endif
#endif /* HVX_H */
