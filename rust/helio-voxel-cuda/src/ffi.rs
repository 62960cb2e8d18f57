//! Raw `extern "C"` declarations: a line-for-line image of include/hvx.h.
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

#[repr(C)]
pub struct hvx_ctx { _private: [u8; 0] }

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct hvx_config {
    pub edge: u32,
    pub max_chunks: u32,
    pub max_vertices: u32,
    pub max_indices: u32,
    pub max_transition_vertices: u32,
    pub max_transition_indices: u32,
    pub flags: u32,
    pub _reserved: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct hvx_chunk_desc {
    pub generation: u64,
    pub dirty_microbricks: u64,
    pub transition_mask: u32,
    pub cost_hint: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default, PartialEq, Eq)]
pub struct hvx_range { pub first_vertex: u32, pub vertex_count: u32, pub first_index: u32, pub index_count: u32 }

#[repr(C)]
pub struct hvx_publisher { _private: [u8; 0] }

/// `GpuSurfaceJob` (PV/src/render.rs:466-480).
#[repr(C, align(16))]
#[derive(Clone, Copy, Debug, Default, PartialEq, Eq)]
pub struct hvx_surface_job {
    pub slot: u32, pub transition_mask: u32, pub generation_low: u32, pub generation_high: u32,
    pub regular_max_vertices: u32, pub regular_max_indices: u32, pub transition_max_vertices: u32, pub transition_max_indices: u32,
    pub regular_max_meshlets: u32, pub transition_max_meshlets: u32, pub _pad: [u32; 2],
}

/// `GpuPageMeta` (crates/helio-planet-voxel-core/src/gpu.rs:62-99).
#[repr(C, align(16))]
#[derive(Clone, Copy, Debug, Default, PartialEq, Eq)]
pub struct hvx_page_meta {
    pub relative_lod0_cell_min: [i32; 3], pub lod: u32, pub slot: u32, pub generation_low: u32, pub generation_high: u32,
    pub transition_mask: u32,
}

/// `GpuSurfaceState` (PV/src/render.rs:505-519).
#[repr(C, align(16))]
#[derive(Clone, Copy, Debug, Default, PartialEq, Eq)]
pub struct hvx_surface_state {
    pub generation_low: u32, pub generation_high: u32, pub active_bank: u32, pub valid: u32,
    pub regular_vertex_count: u32, pub regular_index_count: u32, pub transition_vertex_count: u32, pub transition_index_count: u32,
    pub regular_meshlet_count: u32, pub transition_meshlet_count: u32, pub _pad: [u32; 2],
}

/// `GpuDrawPage` (PV/src/render.rs:533-544).
#[repr(C, align(16))]
#[derive(Clone, Copy, Debug, Default, PartialEq)]
pub struct hvx_draw_page {
    pub relative_lod0_cell_min: [i32; 3], pub lod: u32, pub camera_relative_m: [f32; 3], pub lod0_cell_size_m: f32,
    pub generation_low: u32, pub generation_high: u32, pub transition_mask: u32, pub visible: u32,
}

/// `GpuSurfaceFeedback` (PV/src/render.rs:521-531).
#[repr(C, align(16))]
#[derive(Clone, Copy, Debug, Default, PartialEq, Eq)]
pub struct hvx_surface_feedback {
    pub submitted_jobs: u32, pub published_jobs: u32, pub stale_rejections: u32, pub overflow_rejections: u32,
    pub incomplete_rejections: u32, pub _pad: [u32; 3],
}

/// `DrawIndexedIndirectArgs` (PV/src/render.rs:546-554).
#[repr(C)]
#[derive(Clone, Copy, Debug, Default, PartialEq, Eq)]
pub struct hvx_draw_indexed_indirect {
    pub index_count: u32, pub instance_count: u32, pub first_index: u32, pub base_vertex: i32, pub first_instance: u32,
}

/// `GpuPageTableEntry` (PV/src/table.rs:8-19) -- bytemuck-castable from the reference's own type.
#[repr(C, align(16))]
#[derive(Clone, Copy, Debug, Default, PartialEq, Eq)]
pub struct hvx_page_table_entry {
    pub planet_id: [u32; 4],
    pub relative_lod0_cell_min: [i32; 3],
    pub lod: u32,
    pub slot: u32,
    pub generation_low: u32,
    pub generation_high: u32,
    pub state: u32,
}

/// `GpuResidencyUniform` (PV/src/table.rs:62-72).
#[repr(C, align(16))]
#[derive(Clone, Copy, Debug, Default, PartialEq, Eq)]
pub struct hvx_residency {
    pub table_mask: u32,
    pub max_probe: u32,
    pub resident_pages: u32,
    pub atlas_tiles_x: u32,
    pub atlas_tiles_y: u32,
    pub atlas_tiles_z: u32,
    pub publication_epoch_low: u32,
    pub publication_epoch_high: u32,
}

/// `GpuSurfaceGatherJob` (PV/src/surface_sampling.rs:121-134).
#[repr(C, align(16))]
#[derive(Clone, Copy, Debug, Default, PartialEq, Eq)]
pub struct hvx_gather_job {
    pub planet_id: [u32; 4],
    pub relative_lod0_cell_min: [i32; 3],
    pub lod: u32,
    pub generation_low: u32,
    pub generation_high: u32,
    pub transition_mask: u32,
    pub target_slot: u32,
    pub residency_epoch_low: u32,
    pub residency_epoch_high: u32,
    pub _pad: [u32; 2],
}

/// `GpuSurfaceGatherCounters` (PV/src/surface_sampling.rs:171-181).
#[repr(C, align(16))]
#[derive(Clone, Copy, Debug, Default, PartialEq, Eq)]
pub struct hvx_gather_counters {
    pub regular_samples: u32,
    pub transition_samples: u32,
    pub table_probes: u32,
    pub page_misses: u32,
    pub stale_targets: u32,
    pub completed: u32,
    pub _pad: [u32; 2],
}

pub const HVX_OK: c_int = 0;
pub const HVX_E_SAMPLE_COUNT: c_int = -1;
pub const HVX_E_INVALID_CAPACITY: c_int = -2;
pub const HVX_E_DEVICE_LIMIT: c_int = -3;
pub const HVX_E_TRANSITION_MASK: c_int = -4;
pub const HVX_CFG_DEBUG_RECORDS: u32 = 1;

pub const HVX_BUF_REGULAR_VERTICES: c_int = 2;
pub const HVX_BUF_REGULAR_INDICES: c_int = 3;
pub const HVX_BUF_REGULAR_COUNTERS: c_int = 4;
pub const HVX_BUF_REGULAR_CLASSIFY: c_int = 5;
pub const HVX_BUF_REGULAR_RANGES: c_int = 6;
pub const HVX_BUF_REGULAR_CELLS: c_int = 7;
pub const HVX_BUF_REGULAR_OFFSETS: c_int = 8;
pub const HVX_BUF_REGULAR_BLOCKS: c_int = 9;
pub const HVX_BUF_TRANSITION_VERTICES: c_int = 10;
pub const HVX_BUF_TRANSITION_INDICES: c_int = 11;
pub const HVX_BUF_TRANSITION_COUNTERS: c_int = 12;
pub const HVX_BUF_TRANSITION_CELLS: c_int = 14;

extern "C" {
    pub fn hvx_create(out: *mut *mut hvx_ctx, device: c_int, config: *const hvx_config) -> c_int;
    pub fn hvx_destroy(ctx: *mut hvx_ctx);
    pub fn hvx_last_error(ctx: *const hvx_ctx) -> *const c_char;
    pub fn hvx_allocated_bytes(ctx: *const hvx_ctx) -> u64;
    pub fn hvx_synchronize(ctx: *mut hvx_ctx) -> c_int;
    pub fn hvx_fill_density(ctx: *mut hvx_ctx, kind: u32, page_xyz: *const i64, lod: *const u8, n: u32, d_samples: *mut u32) -> c_int;
    pub fn hvx_fill_slabs(ctx: *mut hvx_ctx, kind: u32, page_xyz: *const i64, lod: *const u8, n: u32, d_slabs: *mut u32) -> c_int;
    pub fn hvx_extract_regular(ctx: *mut hvx_ctx, samples: *const u32, sample_words: u64, descs: *const hvx_chunk_desc, n: u32) -> c_int;
    pub fn hvx_classify_regular(ctx: *mut hvx_ctx, samples: *const u32, sample_words: u64, descs: *const hvx_chunk_desc, n: u32) -> c_int;
    pub fn hvx_extract_transition(ctx: *mut hvx_ctx, slabs: *const u32, slab_words: u64, descs: *const hvx_chunk_desc, n: u32) -> c_int;
    pub fn hvx_build_meshlets(ctx: *mut hvx_ctx, kind: c_int, n: u32) -> c_int;
    pub fn hvx_publisher_create(ctx: *mut hvx_ctx, slots: u32, out: *mut *mut hvx_publisher) -> c_int;
    pub fn hvx_publisher_destroy(publisher: *mut hvx_publisher);
    pub fn hvx_publish_surfaces(publisher: *mut hvx_publisher, jobs: *const hvx_surface_job, job_chunk: *const u32,
                                page_metadata: *const hvx_page_meta, n: u32) -> c_int;
    pub fn hvx_refresh_visibility(publisher: *mut hvx_publisher, draw_pages: *const hvx_draw_page) -> c_int;
    pub fn hvx_publisher_buffer(publisher: *mut hvx_publisher, buffer_id: c_int) -> *mut c_void;
    pub fn hvx_publisher_buffer_bytes(publisher: *mut hvx_publisher, buffer_id: c_int) -> u64;
    pub fn hvx_publisher_read(publisher: *mut hvx_publisher, buffer_id: c_int, byte_offset: u64, bytes: u64, dst: *mut c_void) -> c_int;
    pub fn hvx_publisher_write(publisher: *mut hvx_publisher, buffer_id: c_int, byte_offset: u64, bytes: u64, src: *const c_void) -> c_int;
    pub fn hvx_gather_surface(ctx: *mut hvx_ctx, residency: *const hvx_residency, table: *const hvx_page_table_entry,
                              atlas: *const u32, atlas_words: u64, jobs: *const hvx_gather_job, n: u32) -> c_int;
    pub fn hvx_buffer(ctx: *mut hvx_ctx, buffer_id: c_int) -> *mut c_void;
    pub fn hvx_buffer_bytes(ctx: *mut hvx_ctx, buffer_id: c_int) -> u64;
    pub fn hvx_read(ctx: *mut hvx_ctx, buffer_id: c_int, byte_offset: u64, bytes: u64, host_dst: *mut c_void) -> c_int;
    pub fn hvx_read_meshes(ctx: *mut hvx_ctx, kind: c_int, first: u32, n: u32, vertices_out: *mut c_void, vertex_cap: u64,
                           indices_out: *mut u32, index_cap: u64, ranges_out: *mut hvx_range,
                           total_vertices: *mut u64, total_indices: *mut u64) -> c_int;
}

// ---- legacy 8^3-brick marching cubes (include/hvx.h, "legacy 8^3-brick marching cubes") ----------------
// PODs are the reference's own: GpuBrickMeta / GpuBrickMeshlet (helio-voxel-core/src/gpu_types.rs:18-33),
// DirtyBrick (helio-pass-voxel-mesh/src/lib.rs:64-69).
#[repr(C)] #[derive(Clone, Copy, Default)] pub struct hvx_brick_meta { pub data_offset: u32, pub occupancy: u32 }
#[repr(C)] #[derive(Clone, Copy, Default)] pub struct hvx_dirty_brick { pub brick_slot: u32, pub volume_id: u32, pub _pad: [u32; 2], pub origin_size: [f32; 4] }
#[repr(C)] #[derive(Clone, Copy, Default)] pub struct hvx_brick_meshlet { pub vertex_offset: u32, pub index_offset: u32, pub vertex_count: u32, pub index_count: u32, pub brick_index: u32, pub volume_id: u32, pub _pad: [u32; 2] }
#[repr(C)] pub struct hvx_brick_mesher { _private: [u8; 0] }
pub const HVX_BRICK_VERTICES: c_int = 0;
pub const HVX_BRICK_NORMALS: c_int = 1;
pub const HVX_BRICK_INDICES: c_int = 2;
pub const HVX_BRICK_DESCRIPTORS: c_int = 3;
pub const HVX_BRICK_DRAWS: c_int = 4;
pub const HVX_BRICK_REJECTED: c_int = 5;

extern "C" {
    pub fn hvx_brick_mesher_create(ctx: *mut hvx_ctx, max_bricks: u32, out: *mut *mut hvx_brick_mesher) -> c_int;
    pub fn hvx_brick_mesher_destroy(mesher: *mut hvx_brick_mesher);
    pub fn hvx_brick_extract(mesher: *mut hvx_brick_mesher, meta: *const hvx_brick_meta, n_meta: u32, voxels: *const u32,
                             n_words: u64, dirty: *const hvx_dirty_brick, n_dirty: u32) -> c_int;
    pub fn hvx_brick_clear_slot(mesher: *mut hvx_brick_mesher, brick_slot: u32) -> c_int;
    pub fn hvx_brick_buffer(mesher: *mut hvx_brick_mesher, buffer_id: c_int) -> *mut c_void;
    pub fn hvx_brick_buffer_bytes(mesher: *mut hvx_brick_mesher, buffer_id: c_int) -> u64;
    pub fn hvx_brick_read(mesher: *mut hvx_brick_mesher, buffer_id: c_int, byte_offset: u64, bytes: u64, dst: *mut c_void) -> c_int;
}
