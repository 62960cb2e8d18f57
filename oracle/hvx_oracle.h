/* hvx_oracle.h -- CPU ORACLE for the Transvoxel chunk-extraction hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() may load it.
 * The shipped library (libhelio_voxel_cuda.so) never links, loads or calls it.
 *
 * It is a plain-C restatement of the reference's CPU extractor.  The reference
 * is Rust and no Rust toolchain exists in this image, so the reference itself
 * cannot be executed here; the oracle is pinned instead against every known
 * answer the reference's own tests and docs hold for this path (table audit
 * fingerprint, per-fixture vertex/triangle counts, CellWord packing, the
 * secondary-position case, overflow contract, transition winding ...), see
 * tests/test_oracle_golden.py and SURVEY.md section 8(c).
 *
 * Parity status: integer outputs (case words, ranges, indices, materials,
 * flags, counters) are PINNED at edge 32.  Float outputs are pinned by the
 * reference only to its own tolerances (position 1e-5, normal 2e-4); the exact
 * bit patterns follow the reference's CPU formulae operation by operation.
 * Edge 64, the fBm density quantisation and the dense-random field have no
 * counterpart in the reference: "parity unpinned" for those, they are pinned
 * only by this oracle's generalisation over the edge length.
 *
 * All citations are relative to /root/reference/ ; PV =
 * crates/passes/3d/helio-pass-planetary-voxel.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (see oracle/Makefile).
 * -ffp-contract=off matters: Rust never fuses a*b+c, gcc would under -march=native.
 */
#ifndef HVX_ORACLE_H
#define HVX_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* PV/src/extraction.rs:72-79 GpuTerrainVertex, repr(C, align(16)), 32 bytes. */
typedef struct {
    float position[3];
    uint32_t material;
    float normal[3];
    uint32_t flags;
} hvxo_vertex;

/* PV/src/fixture.rs:9-16 ExtractionFixtureKind (same order). */
enum {
    HVXO_FIELD_PLANE = 0,
    HVXO_FIELD_SPHERE = 1,
    HVXO_FIELD_CAVE = 2,
    HVXO_FIELD_SHARP_CORNER = 3,
    HVXO_FIELD_THIN_SLAB = 4,
    HVXO_FIELD_MATERIAL_SEAM = 5,
    /* ours: no reference counterpart, parity unpinned */
    HVXO_FIELD_TERRAIN_FBM = 16,
    HVXO_FIELD_DENSE_RANDOM = 17
};

/* PV/src/transvoxel.rs:248-261 TransvoxelTableAudit. */
typedef struct {
    uint32_t regular_cases, transition_cases;
    uint32_t regular_vertices, regular_triangles;
    uint32_t transition_vertices, transition_triangles;
    uint32_t max_regular_vertices, max_regular_triangles;
    uint32_t max_transition_vertices, max_transition_triangles;
    uint64_t fingerprint;
} hvxo_table_audit;

/* PV/src/fixture.rs:74-81 ExtractionFixtureMetrics. */
typedef struct {
    uint32_t solid_samples, air_samples, active_cells, _pad;
    uint64_t active_microbrick_mask;
    uint64_t fingerprint;
} hvxo_fixture_metrics;

/* helio-planet-voxel-core/src/types.rs:335-337 CellWord::new. */
uint32_t hvxo_cellword(int16_t density, uint8_t material, uint8_t flags);

/* PV/src/transvoxel.rs:263-379 validate_transvoxel_tables; returns 0 when every
 * case is index safe, else a negative code. */
int hvxo_validate_tables(hvxo_table_audit* out);
const char* hvxo_table_revision(void);
/* PV/src/transvoxel.rs:201-238: raw per-case topology, for table tests.
 * kind 0 regular / 1 transition; triangles already have the inverse flip applied. */
int hvxo_case_topology(int kind, uint32_t case_index, uint32_t* class_index, uint32_t* reverse,
                       uint32_t* vertex_count, uint32_t* triangle_count, uint16_t codes[12],
                       uint8_t triangles[36]);

/* PV/src/fixture.rs:43-71 sample_canonical (+ our two extra fields). */
uint32_t hvxo_sample_canonical(int kind, const int64_t position[3], uint32_t lod);
/* PV/src/fixture.rs:95-124 ExtractionFixture::new, generalised over edge.
 * samples: (edge+2)^3 words, x fastest then y then z. */
int hvxo_fixture_fill(int kind, int edge, uint32_t lod, const int64_t page_xyz[3], uint32_t* samples);
/* PV/src/fixture.rs:169-197 measure. */
int hvxo_fixture_metrics_of(int edge, const uint32_t* samples, hvxo_fixture_metrics* out);
/* PV/src/transvoxel_transition.rs:335-349 slab_samples for all six faces
 * (PV/tests/gpu_transvoxel_transitions.rs:389-403 layout): 6*(2*edge+3)^2*3 words. */
int hvxo_slab_fill(int kind, int edge, uint32_t lod, const int64_t page_xyz[3], uint32_t* slabs);

/* crates/passes/3d/helio-pass-sdf/src/noise.rs:139-208 terrain_sdf, Rolling style,
 * TerrainConfig::rolling() (terrain.rs:40-51). */
float hvxo_terrain_sdf_rolling(float x, float y, float z);

/* Counter layouts:
 *   classify[4]  PV/src/transvoxel_gpu.rs:134-141  visited, active, vertices, triangles
 *   emission[8]  PV/src/transvoxel_emit.rs:38-48   required_v, required_i, emitted_v, emitted_i,
 *                                                   vertex_overflow, index_overflow, completed, pad
 *   transition[12] PV/src/transvoxel_transition_gpu.rs:133-146 active_cells, active_faces,
 *                  required_v, required_i, emitted_v, emitted_i, v_overflow, i_overflow, completed, pad*3
 */

/* PV/tests/gpu_transvoxel_emission.rs:272-344 expected_mesh +
 * PV/tests/gpu_transvoxel.rs:150-186 expected_classification, generalised over edge.
 *   cell_words  : edge^3 * 4 u32  (GpuTransvoxelCell; untouched for non-dirty cells)  or NULL
 *   cell_ranges : edge^3 * 2 u32  (first_vertex, first_index; untouched for non-dirty) or NULL
 * vertices/indices receive min(required, cap) entries; the counters carry the
 * all-or-nothing overflow contract against max_vertices / max_indices. */
int hvxo_extract_regular(int edge, const uint32_t* samples, uint64_t generation,
                         uint64_t dirty_microbricks, uint32_t transition_mask,
                         uint32_t max_vertices, uint32_t max_indices,
                         hvxo_vertex* vertices, uint32_t vertex_cap,
                         uint32_t* indices, uint32_t index_cap,
                         uint32_t* cell_words, uint32_t* cell_ranges,
                         uint32_t classify[4], uint32_t emission[8]);

/* PV/src/transvoxel_transition.rs:189-270,366-397 +
 * PV/tests/gpu_transvoxel_transitions.rs:351-387, consuming the six-face slab
 * block exactly as the GPU extractor does (gradients from the slab halo).
 *   cell_words  : 6*edge^2 * 4 u32 (GpuTransvoxelTransitionCell; zeroed for inactive faces) or NULL
 *   cell_ranges : 6*edge^2 * 2 u32 or NULL */
int hvxo_extract_transition(int edge, const uint32_t* slabs, uint32_t transition_mask,
                            uint64_t generation, uint32_t max_vertices, uint32_t max_indices,
                            hvxo_vertex* vertices, uint32_t vertex_cap,
                            uint32_t* indices, uint32_t index_cap,
                            uint32_t* cell_words, uint32_t* cell_ranges, uint32_t counters[12]);

/* The reference's own route: evaluate the analytic field around every face
 * sample (PV/src/transvoxel_transition.rs:313-329,399-424) instead of reading a
 * slab.  Used to prove the slab-halo gradient derivation is exact.  One face. */
int hvxo_extract_transition_face_analytic(int kind, int edge, uint32_t lod,
                                          const int64_t page_xyz[3], int face,
                                          hvxo_vertex* vertices, uint32_t vertex_cap,
                                          uint32_t* indices, uint32_t index_cap,
                                          uint32_t* vertex_count, uint32_t* index_count);

/* Batch driver for the CPU baseline: fill (optional) + regular extraction of n
 * chunks with OpenMP `schedule(dynamic)` over chunks.  Returns total cells.
 * totals[0..3] = sum required vertices, sum required indices, active cells, chunks with output. */
int64_t hvxo_batch_regular(int kind, int edge, uint32_t lod, const int64_t* page_xyz /* n*3 */,
                           uint32_t n, int do_fill, const uint32_t* samples_or_null,
                           int threads, uint64_t totals[4]);
/* OpenMP fill of n chunks (setup helper for the CPU baseline; same result as hvxo_fixture_fill). */
int hvxo_batch_fill(int kind, int edge, uint32_t lod, const int64_t* page_xyz, uint32_t n, int threads, uint32_t* out);
int hvxo_max_threads(void);

/* PV/src/extraction.rs:81-92 GpuTerrainMeshlet and PV/src/terrain_meshlet.rs:10-20 GpuTerrainMeshletBounds. */
typedef struct {
    uint32_t first_index, index_count, first_vertex, vertex_count, bounds_offset, generation_low, generation_high, _pad;
} hvxo_meshlet;
typedef struct {
    float center[3], radius, cone_apex[3], cone_cutoff, cone_axis[3], _pad;
} hvxo_meshlet_bounds;

/* build_terrain_meshlets (PV/src/terrain_meshlet.rs:83-155): fixed 63-index partition, AABB-centre
 * bounding sphere, normal cone.  Returns the meshlet count (<= capacity) or a negative error:
 * -1 IncompleteTriangle, -2 IndexOutOfBounds, -3 NonFinitePosition, -4 capacity. */
int hvxo_build_meshlets(const hvxo_vertex* vertices, uint32_t vertex_count, const uint32_t* indices,
                        uint32_t index_count, uint32_t first_index, uint32_t first_vertex, uint32_t first_bounds,
                        uint64_t generation, uint32_t flags, hvxo_meshlet* meshlets, hvxo_meshlet_bounds* bounds,
                        uint32_t capacity);

#ifdef __cplusplus
}
#endif
#endif

/* ---- surface gather: literal restatement of PV/src/surface_gather.wgsl ------------- */
typedef struct {
    uint32_t planet_id[4];
    int32_t relative_lod0_cell_min[3];
    uint32_t lod, slot, generation_low, generation_high, state;
} hvxo_page_table_entry; /* GpuPageTableEntry, PV/src/table.rs:8-19 */
typedef struct {
    uint32_t table_mask, max_probe, resident_pages, atlas_tiles_x, atlas_tiles_y, atlas_tiles_z;
    uint32_t publication_epoch_low, publication_epoch_high;
} hvxo_residency; /* GpuResidencyUniform, PV/src/table.rs:62-72 */
typedef struct {
    uint32_t planet_id[4];
    int32_t relative_lod0_cell_min[3];
    uint32_t lod, generation_low, generation_high, transition_mask, target_slot;
    uint32_t residency_epoch_low, residency_epoch_high, _pad[2];
} hvxo_gather_job; /* GpuSurfaceGatherJob, PV/src/surface_sampling.rs:121-134 */
typedef struct {
    uint32_t regular_samples, transition_samples, table_probes, page_misses, stale_targets, completed, _pad[2];
} hvxo_gather_counters; /* GpuSurfaceGatherCounters, PV/src/surface_sampling.rs:171-181 */

/* GpuLookupKey::hash (PV/src/table.rs:148-160) */
uint32_t hvxo_page_hash(const uint32_t planet_id[4], const int32_t relative_min[3], uint32_t lod);
/* gather_regular + gather_transition + finalize_gather for one job.  regular: 34^3 words,
 * transition: 6*67*67*3 words (faces outside the mask are left untouched), indirect: 24 words
 * (written only when the job completes).  One table walk per sample, exactly like the shader. */
void hvxo_gather_surface(const hvxo_residency* residency, const hvxo_page_table_entry* table, const uint32_t* atlas,
                         const hvxo_gather_job* job, uint32_t* regular, uint32_t* transition,
                         hvxo_gather_counters* counters, uint32_t* indirect);

/* ---- surface publication: literal restatement of PV/src/surface_publish.wgsl ---------- */
typedef struct {
    uint32_t slot, transition_mask, generation_low, generation_high;
    uint32_t regular_max_vertices, regular_max_indices, transition_max_vertices, transition_max_indices;
    uint32_t regular_max_meshlets, transition_max_meshlets, _pad[2];
} hvxo_surface_job; /* GpuSurfaceJob, PV/src/render.rs:466-480 */
typedef struct {
    int32_t relative_lod0_cell_min[3];
    uint32_t lod, slot, generation_low, generation_high, transition_mask;
} hvxo_page_meta; /* GpuPageMeta, crates/helio-planet-voxel-core/src/gpu.rs:62-99 */
typedef struct {
    uint32_t generation_low, generation_high, active_bank, valid;
    uint32_t regular_vertex_count, regular_index_count, transition_vertex_count, transition_index_count;
    uint32_t regular_meshlet_count, transition_meshlet_count, _pad[2];
} hvxo_surface_state; /* GpuSurfaceState, PV/src/render.rs:505-519 */
typedef struct {
    uint32_t submitted_jobs, published_jobs, stale_rejections, overflow_rejections, incomplete_rejections, _pad[3];
} hvxo_surface_feedback; /* GpuSurfaceFeedback, PV/src/render.rs:521-531 */
typedef struct {
    uint32_t index_count, instance_count, first_index;
    int32_t base_vertex;
    uint32_t first_instance;
} hvxo_draw_args; /* DrawIndexedIndirectArgs, PV/src/render.rs:546-554 */

/* One job: copy_regular_surface, copy_transition_surface, publish_surface (surface_publish.wgsl:125-216).
 * regular_counters: 8 words (GpuTransvoxelEmissionCounters), transition_counters: 12 words
 * (GpuTransvoxelTransitionCounters).  Source meshes start at element 0; destination arenas are the
 * double-banked per-slot arenas (any of the four mesh pointers may be NULL to skip that copy). */
void hvxo_publish_surface(const hvxo_surface_job* job, const hvxo_page_meta* page_metadata, const uint32_t* regular_counters,
                          const uint32_t* transition_counters, const hvxo_vertex* src_vertices, const uint32_t* src_indices,
                          const hvxo_vertex* src_tvertices, const uint32_t* src_tindices, hvxo_surface_state* states,
                          hvxo_vertex* vertices, uint32_t* indices, hvxo_vertex* tvertices, uint32_t* tindices,
                          hvxo_draw_args* regular_draws, hvxo_draw_args* transition_draws, hvxo_surface_feedback* feedback);
/* refresh_visibility (surface_publish.wgsl:218-225); visible[i] is GpuDrawPage.visible of slot i */
void hvxo_refresh_visibility(uint32_t slots, const hvxo_surface_state* states, const uint32_t* visible,
                             hvxo_draw_args* regular_draws, hvxo_draw_args* transition_draws);
