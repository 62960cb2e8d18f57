// hvx_kernels.h -- launch interface between the C-ABI layer (hvx_api.cu) and the kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hvx.h"

namespace hvx {

// Device copy of one chunk's dispatch parameters (hvx_chunk_desc).
struct ChunkDesc {
    uint64_t generation;
    uint64_t dirty_microbricks;
    uint32_t transition_mask;
    uint32_t cost_hint;
    uint32_t flags;
    uint32_t _reserved;
};

enum : uint32_t {
    MODE_EXTRACT = 0,
    MODE_CLASSIFY = 1,
    MODE_STREAM_ONLY = 2,
    MODE_BITS_ONLY = 3,
};

// One z-range of a chunk (a work item of the split walk): steps [s_first, s_last] of the chunk's NSLAB - 1 steps,
// part `part` of `parts`.  The parts of a chunk are consecutive items in ascending order.
struct SplitItem {
    uint32_t chunk;
    uint8_t s_first, s_last, part, parts;
};

struct RegularParams {
    const uint32_t* samples;  // [n][(E+2)^3]
    const ChunkDesc* descs;   // [n] device
    const uint32_t* order;    // nullable, [n_work] device: the k-th chunk (split walk: item) to start (descending cost hints;
                              // chunks flagged uniform are left out)
    const uint32_t* skipped;  // [n_skipped] device: chunks of this launch that are not walked (uniform / nothing dirty); the
    uint32_t n_skipped;       // decoupled kernel writes their empty records itself (ids relative to this launch)
    const SplitItem* items;   // nullable, device: the split walk's work items; ids in `order` / tickets then name items
    uint4* item_totals;       // [items] device: (vertices, indices, active cells, generation) of every part: the look-back state
    uint32_t split_generation;  // tag of this dispatch's look-back entries (never 0, never repeated while the ctx lives)
    uint32_t n_chunks;        // chunks addressed by this launch (ids are < n_chunks)
    uint32_t n_work;          // chunks it actually walks (== n_chunks when order is NULL)
    uint32_t chunk_base;      // batch index of this launch's chunk 0 (a sub-batch of a pipelined dispatch); every pointer
                              // below is already offset, only the ranges record absolute slots
    uint32_t mode;
    uint32_t debug_flags;  // diagnostics: bit 0 disables the classification fast-reject
    uint32_t any_partial;  // some chunk of the batch is partially dirty (picks the slab-skipping instantiation)
    uint32_t first_generation;  // HVX_CFG_FIRST_GENERATION: run the first-generation kernel (cross-check only)
    uint32_t max_vertices, max_indices;  // per-chunk slot capacity
    hvx_vertex* vertices;                // [n][max_vertices]
    uint32_t* indices;                   // [n][max_indices]
    hvx_emission_counters* counters;     // [n]
    hvx_classify_counters* classify;     // [n]
    hvx_range* ranges;                   // [n]
    hvx_cell_record* cells;              // debug, nullable: [n][E^3]
    hvx_cell_offset* offsets;            // debug, nullable
    hvx_scan_block* blocks;              // debug, nullable: [n][E^3/256]
    uint32_t* work_counter;              // dynamic chunk queue, zeroed before launch
};

struct TransitionParams {
    const uint32_t* slabs;  // [n][6][3][(2E+3)^2]
    const ChunkDesc* descs;
    uint32_t n_chunks;
    uint32_t max_vertices, max_indices;
    hvx_vertex* vertices;
    uint32_t* indices;
    hvx_transition_counters* counters;
    hvx_range* ranges;
    hvx_cell_record* cells;    // debug, nullable: [n][6*E^2]
    hvx_cell_offset* offsets;  // debug, nullable
    hvx_scan_block* blocks;    // debug, nullable: [n][6*E^2/256]
    uint32_t* work_counter;
};

struct FillParams {
    uint32_t kind;
    uint32_t n_chunks;
    const int64_t* page_xyz;  // [n][3] device
    const uint8_t* lod;       // [n] device
    uint32_t* out;
    // fBm terrain only (kind 16, sample fill): shared height maps of the batch's distinct (x, z, lod)
    // columns, heights[col][(E+2)][(E+2)], and each chunk's column; NULL = compute per chunk
    const float* heights;
    const uint32_t* col_index;
};

struct EditParams {
    uint32_t op;              // 1 AddSphere, 2 SubtractSphere (GpuVoxelEdit::op_type)
    uint32_t material;
    float center[3];          // metres
    float radius;             // metres
    uint32_t n_touched;
    const uint32_t* ids;      // [n_touched] device: chunks whose sample block meets the sphere
    const int64_t* page_xyz;  // [n][3] device
    const uint8_t* lod;       // [n] device
    uint32_t* samples;        // [n][(E+2)^3]
};

struct GatherParams {
    uint32_t n_jobs;
    uint32_t job_base;                 // first job of this launch (grids tile the batch in steps of 65,535 jobs)
    hvx_residency residency;
    const hvx_page_table_entry* table;
    const uint32_t* atlas;
    const hvx_gather_job* jobs;
    uint32_t* samples;                 // [n][34^3]
    uint32_t* slabs;                   // [n][6*3*67^2]
    hvx_gather_counters* counters;     // [n], zeroed before the launch
    uint32_t* indirect;                // [n][24]
};

struct PublishParams {
    uint32_t n_jobs, slots;
    uint32_t job_base;
    uint32_t has_transition;
    uint32_t src_max_vertices, src_max_indices, src_max_tvertices, src_max_tindices;  // extraction slot strides
    const hvx_surface_job* jobs;
    const uint32_t* job_chunk;
    const hvx_page_meta* meta;
    const hvx_emission_counters* regular_counters;
    const hvx_transition_counters* transition_counters;
    const hvx_vertex* src_vertices;
    const uint32_t* src_indices;
    const hvx_vertex* src_tvertices;
    const uint32_t* src_tindices;
    hvx_surface_state* states;
    hvx_vertex* vertices;
    uint32_t* indices;
    hvx_vertex* tvertices;
    uint32_t* tindices;
    hvx_draw_indexed_indirect* regular_draws;
    hvx_draw_indexed_indirect* transition_draws;
    hvx_surface_feedback* feedback;
    const hvx_draw_page* draw_pages;
};

// hvx_extraction_commit: one reserved page of the bounded extraction publisher (PV/src/extraction.rs:283-288)
struct CommitJob {
    uint32_t chunk, page_slot;
    hvx_extraction_range range;  // reserved placement + generation (SurfaceAllocation::gpu_range)
};

struct CommitParams {
    uint32_t n_jobs;
    uint32_t job_base;
    uint32_t src_max_vertices, src_max_indices;  // extraction slot strides
    const CommitJob* jobs;
    const hvx_emission_counters* regular_counters;
    const hvx_vertex* src_vertices;
    const uint32_t* src_indices;
    hvx_vertex* vertices;                        // bounded arenas
    uint32_t* indices;
    hvx_extraction_range* page_ranges;
    hvx_extraction_counters* counters;
};

// hvx_brick_extract: legacy 8^3-brick marching cubes (helio-pass-voxel-mesh)
struct BrickParams {
    uint32_t n_dirty;
    uint32_t slot_limit;  // min(max_bricks, n_meta)
    uint64_t n_words;
    uint32_t* rejected;   // entries skipped by the in-kernel bounds check
    const hvx_brick_meta* meta;
    const uint32_t* voxels;
    const hvx_dirty_brick* dirty;
    float4* vertices;
    float4* normals;
    uint32_t* indices;
    hvx_brick_meshlet* descriptors;
    hvx_draw_indexed_indirect* draws;
};

struct MeshletParams {
    uint32_t n_chunks;
    uint32_t chunk_base;
    uint32_t transition;                  // 0 regular, 1 transition
    uint32_t max_vertices, max_indices;   // per-chunk slot capacity of the source arenas
    uint32_t max_meshlets;                // ceil(max_indices / 63)
    const hvx_vertex* vertices;
    const uint32_t* indices;
    const hvx_emission_counters* regular_counters;
    const hvx_transition_counters* transition_counters;
    const ChunkDesc* descs;
    hvx_meshlet* meshlets;                // [n][max_meshlets]
    hvx_meshlet_bounds* bounds;           // [n][max_meshlets]
    uint32_t* meshlet_counts;             // [n]
};

// hvx_weld_meshes: merge bit-identical vertex records of each chunk's mesh in place (mesh_weld.cu)
struct WeldParams {
    uint32_t n_chunks;
    uint32_t max_vertices, max_indices;   // per-chunk slot capacity of the arenas
    hvx_vertex* vertices;
    uint32_t* indices;
    hvx_range* ranges;                    // vertex_count is rewritten
    void* counters;                       // hvx_emission_counters or hvx_transition_counters, [n]
    uint32_t counter_stride, emitted_vertices_offset;
    uint32_t* scratch;                    // [ctas][scratch_words_per_cta]: table | representative | destination
    uint32_t table_words;                 // power of two >= 2 * max_vertices
    uint32_t scratch_words_per_cta;       // table_words + 2 * max_vertices
    uint32_t* work_counter;               // [0] tickets, [2] CTAs that have left (rearm_work_counter)
};

struct DeviceInfo {
    int ordinal;
    int sm_count;
    int max_smem_optin;
};

// Each returns the number of kernels launched (>= 1) or a negative cudaError-mapped status.
cudaError_t launch_regular(int edge, const RegularParams& p, const DeviceInfo& dev, cudaStream_t stream);
// per-cell records, block-relative offsets and scan blocks (HVX_CFG_DEBUG_RECORDS), after the extraction
cudaError_t launch_regular_records(int edge, const RegularParams& p, cudaStream_t stream);
// empty, completed records for chunks flagged HVX_CHUNK_UNIFORM (ids = batch indices)
cudaError_t launch_uniform_records(int edge, const RegularParams& p, const uint32_t* ids, uint32_t n, cudaStream_t stream);
cudaError_t launch_transition(int edge, const TransitionParams& p, const DeviceInfo& dev, cudaStream_t stream);
cudaError_t launch_fill_samples(int edge, const FillParams& p, const DeviceInfo& dev, cudaStream_t stream);
cudaError_t launch_copy_segments(const uint32_t* src, uint32_t* dst, const uint64_t* d_segments, uint32_t n, cudaStream_t stream);
cudaError_t launch_edit_sphere(int edge, const EditParams& p, cudaStream_t stream);
cudaError_t launch_fill_slabs(int edge, const FillParams& p, const DeviceInfo& dev, cudaStream_t stream);
cudaError_t launch_terrain_heights(int edge, const long long* col_xz, const uint8_t* col_lod, uint32_t n_cols, float* heights,
                                   cudaStream_t stream);
cudaError_t launch_meshlets(const MeshletParams& p, const DeviceInfo& dev, cudaStream_t stream);
cudaError_t launch_weld(const WeldParams& p, const DeviceInfo& dev, uint32_t ctas, cudaStream_t stream);
cudaError_t launch_gather(const GatherParams& p, const DeviceInfo& dev, cudaStream_t stream);
cudaError_t launch_publish(const PublishParams& p, const DeviceInfo& dev, cudaStream_t stream);
cudaError_t launch_visibility(const PublishParams& p, const DeviceInfo& dev, cudaStream_t stream);
cudaError_t launch_bricks(const BrickParams& p, const DeviceInfo& dev, cudaStream_t stream);
cudaError_t launch_commit(const CommitParams& p, const DeviceInfo& dev, cudaStream_t stream);
// Packs per-chunk slots into a dense staging arena (for hvx_read_meshes).
cudaError_t launch_pack(const hvx_vertex* vertices, const uint32_t* indices, const hvx_range* slot_ranges,
                        const hvx_range* packed_ranges, uint32_t n, hvx_vertex* out_vertices,
                        uint32_t* out_indices, const DeviceInfo& dev, cudaStream_t stream);

const char* regular_kernel_name(int edge, bool first_generation, bool partial);
// shared-memory footprint of the regular kernel (for resource reporting / tests)
size_t regular_smem_bytes(int edge);

}  // namespace hvx
