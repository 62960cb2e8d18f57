//! Safe wrapper with the reference's names: swap
//! `helio_pass_planetary_voxel::{TransvoxelGpuExtractor, TransvoxelGpuClassifier, TransvoxelGpuTransitionExtractor}`
//! (PV/src/render.rs:579-580 holds the extractor fields) for the types below.  Where the reference takes
//! `(&wgpu::Device, &wgpu::Queue)` these take a CUDA device ordinal at construction.
//!
//! * one page per dispatch, like the reference: [`TransvoxelGpuExtractor`], [`TransvoxelGpuClassifier`],
//!   [`TransvoxelGpuTransitionExtractor`].  Meshes always come from the production (decoupled) kernel; the per-cell
//!   record buffers the reference exposes (`output_buffer` / `offsets_buffer` / `blocks_buffer`) are produced by two
//!   extra launches and only when the extractor was built `with_records`.
//! * the batch form the page queue feeds once it stops submitting one page per frame: [`ChunkBatchExtractor`]
//!   (`prepare` / `encode` split like PV/src/transvoxel_emit.rs:256-330: `prepare` validates and stages the
//!   descriptors, `encode` queues the work on the ctx stream, `read_*` wait), plus the steps before and after it in
//!   the reference's pass: [`GpuSurfaceSampler`] (gather), meshlets, [`SurfacePublisher`] (publication).
//! * every fallible call returns `Result`; nothing panics on a CUDA error.
//!
//! Source only: the authoring image has no Rust toolchain.  `ffi` is generated from
//! include/hvx.h by tools/gen_rust_ffi.py and checked against the header by tests/test_ffi_consistency.py.
pub mod extraction;
pub mod ffi;

use bytemuck::{Pod, Zeroable};
use core::ffi::{c_int, c_void, CStr};
use helio_planet_voxel_core::CellWord;

pub use helio_pass_planetary_voxel_types::*;

/// Re-declared PODs: byte-identical to PV/src/extraction.rs:72-79, PV/src/transvoxel_emit.rs:14-48 and
/// PV/src/transvoxel_gpu.rs:77-141, so a Helio build can instead `pub use helio_pass_planetary_voxel::{...}`.
pub mod helio_pass_planetary_voxel_types {
    use super::*;
    #[repr(C, align(16))]
    #[derive(Clone, Copy, Debug, Default, PartialEq, Pod, Zeroable)]
    pub struct GpuTerrainVertex { pub position: [f32; 3], pub material: u32, pub normal: [f32; 3], pub flags: u32 }
    #[repr(C, align(16))]
    #[derive(Clone, Copy, Debug, Default, PartialEq, Eq, Pod, Zeroable)]
    pub struct GpuTransvoxelEmissionCounters {
        pub required_vertices: u32, pub required_indices: u32, pub emitted_vertices: u32, pub emitted_indices: u32,
        pub vertex_overflow: u32, pub index_overflow: u32, pub completed: u32, pub _pad: u32,
    }
    impl GpuTransvoxelEmissionCounters { pub const fn overflowed(self) -> bool { self.vertex_overflow != 0 || self.index_overflow != 0 } }
    #[repr(C, align(16))]
    #[derive(Clone, Copy, Debug, Default, PartialEq, Eq, Pod, Zeroable)]
    pub struct GpuTransvoxelClassifyCounters { pub visited_cells: u32, pub active_cells: u32, pub vertices: u32, pub triangles: u32 }
    #[repr(C, align(16))]
    #[derive(Clone, Copy, Debug, Default, PartialEq, Eq, Pod, Zeroable)]
    pub struct GpuTransvoxelCell { pub packed_case_class_counts: u32, pub generation_low: u32, pub generation_high: u32, pub _pad: u32 }
    #[repr(C, align(16))]
    #[derive(Clone, Copy, Debug, Default, PartialEq, Eq, Pod, Zeroable)]
    pub struct GpuTransvoxelCellOffset { pub first_vertex: u32, pub first_index: u32, pub generation_low: u32, pub generation_high: u32 }
    #[repr(C, align(16))]
    #[derive(Clone, Copy, Debug, Default, PartialEq, Eq, Pod, Zeroable)]
    pub struct GpuTransvoxelScanBlock { pub vertex_count: u32, pub index_count: u32, pub first_vertex: u32, pub first_index: u32 }
    #[repr(C, align(16))]
    #[derive(Clone, Copy, Debug, Default, PartialEq, Eq, Pod, Zeroable)]
    pub struct GpuTransvoxelTransitionCounters {
        pub active_cells: u32, pub active_faces: u32, pub required_vertices: u32, pub required_indices: u32,
        pub emitted_vertices: u32, pub emitted_indices: u32, pub vertex_overflow: u32, pub index_overflow: u32,
        pub completed: u32, pub _pad: [u32; 3],
    }
}

/// PV/src/transvoxel_gpu.rs:445-459 (+ the transition twin's `TransitionMask`).
#[derive(Clone, Debug, PartialEq, Eq, thiserror::Error)]
pub enum TransvoxelGpuError {
    #[error("Transvoxel classification received {actual} samples; expected {expected}")]
    SampleCount { actual: usize, expected: usize },
    #[error("Transvoxel extraction capacities must be nonzero (vertices={max_vertices}, indices={max_indices})")]
    InvalidExtractionCapacity { max_vertices: u32, max_indices: u32 },
    #[error("{0}")]
    DeviceLimit(String),
    #[error("transition mask {0:#010b} uses bits outside the six page faces")]
    TransitionMask(u8),
    #[error("batch of {requested} chunks exceeds the extractor's capacity {capacity}")]
    BatchCapacity { requested: usize, capacity: u32 },
    #[error("the extractor was built without per-cell records; use `with_records`")]
    RecordsDisabled,
    #[error("CUDA extraction failed ({status}): {message}")]
    Backend { status: i32, message: String },
}

#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub struct TransvoxelGpuExtractorConfig { pub max_vertices: u32, pub max_indices: u32 }
impl TransvoxelGpuExtractorConfig {
    pub fn new(max_vertices: u32, max_indices: u32) -> Result<Self, TransvoxelGpuError> {
        if max_vertices == 0 || max_indices == 0 {
            return Err(TransvoxelGpuError::InvalidExtractionCapacity { max_vertices, max_indices });
        }
        Ok(Self { max_vertices, max_indices })
    }
}
impl Default for TransvoxelGpuExtractorConfig {
    fn default() -> Self { Self { max_vertices: 393_216, max_indices: 491_520 } }
}

const PAGE_SAMPLES: usize = 34 * 34 * 34;             // EXTRACTION_SAMPLE_COUNT, PV/src/fixture.rs:5-6
const PAGE_CELLS: usize = 32 * 32 * 32;
const SLAB_SAMPLES: usize = 6 * 3 * 67 * 67;          // TRANSITION_ALL_FACE_SLAB_SAMPLE_COUNT

/// One `hvx_ctx`: a device, a stream, fixed-stride output arenas.
struct Ctx { raw: *mut ffi::hvx_ctx, config: ffi::hvx_config }
unsafe impl Send for Ctx {}
unsafe impl Sync for Ctx {} // calls are serialised by `&mut self` / the pass's single queue, like the reference
impl Drop for Ctx { fn drop(&mut self) { unsafe { ffi::hvx_destroy(self.raw) } } }

impl Ctx {
    fn new(cuda_device: i32, config: ffi::hvx_config) -> Result<Self, TransvoxelGpuError> {
        let mut raw = core::ptr::null_mut();
        let status = unsafe { ffi::hvx_create(&mut raw, cuda_device, &config) };
        if status != ffi::HVX_OK {
            return Err(Self::error_of(core::ptr::null(), status, 0, 0, 0, &config));
        }
        Ok(Self { raw, config })
    }
    fn error_of(ctx: *const ffi::hvx_ctx, status: c_int, actual: usize, expected: usize, mask: u8, config: &ffi::hvx_config) -> TransvoxelGpuError {
        let message = unsafe { CStr::from_ptr(ffi::hvx_last_error(ctx)) }.to_string_lossy().into_owned();
        match status {
            ffi::HVX_E_SAMPLE_COUNT => TransvoxelGpuError::SampleCount { actual, expected },
            ffi::HVX_E_INVALID_CAPACITY => TransvoxelGpuError::InvalidExtractionCapacity {
                max_vertices: config.max_vertices, max_indices: config.max_indices },
            ffi::HVX_E_DEVICE_LIMIT => TransvoxelGpuError::DeviceLimit(message),
            ffi::HVX_E_TRANSITION_MASK => TransvoxelGpuError::TransitionMask(mask),
            ffi::HVX_E_BATCH_CAPACITY => TransvoxelGpuError::BatchCapacity { requested: actual, capacity: config.max_chunks },
            _ => TransvoxelGpuError::Backend { status, message },
        }
    }
    fn check(&self, status: c_int, actual: usize, expected: usize, mask: u8) -> Result<(), TransvoxelGpuError> {
        if status == ffi::HVX_OK { Ok(()) } else { Err(Self::error_of(self.raw, status, actual, expected, mask, &self.config)) }
    }
    /// Synchronising read of `count` elements of an arena, starting at element `first`.
    fn read<T: Pod>(&self, buffer: c_int, first: usize, count: usize) -> Result<Vec<T>, TransvoxelGpuError> {
        let mut out = vec![T::zeroed(); count];
        let size = core::mem::size_of::<T>();
        let status = unsafe { ffi::hvx_read(self.raw, buffer, (first * size) as u64, (count * size) as u64, out.as_mut_ptr() as *mut c_void) };
        self.check(status, 0, 0, 0).map(|_| out)
    }
    fn has_records(&self) -> bool { self.config.flags & ffi::HVX_CFG_DEBUG_RECORDS != 0 }
}

fn desc(generation: u64, dirty_microbricks: u64, transition_mask: u8) -> ffi::hvx_chunk_desc {
    ffi::hvx_chunk_desc { generation, dirty_microbricks, transition_mask: transition_mask as u32, ..Default::default() }
}

/// Drop-in for `helio_pass_planetary_voxel::TransvoxelGpuExtractor` (PV/src/transvoxel_emit.rs:92-396).
pub struct TransvoxelGpuExtractor { ctx: Ctx, config: TransvoxelGpuExtractorConfig }

impl TransvoxelGpuExtractor {
    /// Meshes and counters only (what the render pass consumes): one kernel launch per dispatch.
    pub fn new(cuda_device: i32, config: TransvoxelGpuExtractorConfig) -> Result<Self, TransvoxelGpuError> {
        Self::build(cuda_device, config, 0)
    }
    /// Also keep the reference's per-cell buffers (`cells`, `offsets`, `blocks`): two extra launches per dispatch.
    pub fn with_records(cuda_device: i32, config: TransvoxelGpuExtractorConfig) -> Result<Self, TransvoxelGpuError> {
        Self::build(cuda_device, config, ffi::HVX_CFG_DEBUG_RECORDS)
    }
    fn build(cuda_device: i32, config: TransvoxelGpuExtractorConfig, flags: u32) -> Result<Self, TransvoxelGpuError> {
        let config = TransvoxelGpuExtractorConfig::new(config.max_vertices, config.max_indices)?;
        let raw = ffi::hvx_config { edge: 32, max_chunks: 1, max_vertices: config.max_vertices,
            max_indices: config.max_indices, flags, ..Default::default() };
        Ok(Self { ctx: Ctx::new(cuda_device, raw)?, config })
    }

    /// PV/src/transvoxel_emit.rs:233-254
    pub fn dispatch(&self, samples: &[CellWord], generation: u64, dirty_microbricks: u64, transition_mask: u8)
        -> Result<(), TransvoxelGpuError> {
        let d = desc(generation, dirty_microbricks, transition_mask);
        let words: &[u32] = bytemuck::cast_slice(samples);
        let status = unsafe { ffi::hvx_extract_regular(self.ctx.raw, words.as_ptr(), words.len() as u64, &d, 1) };
        self.ctx.check(status, samples.len(), PAGE_SAMPLES, transition_mask)
    }

    pub fn counters(&self) -> Result<GpuTransvoxelEmissionCounters, TransvoxelGpuError> {
        Ok(self.ctx.read(ffi::HVX_BUF_REGULAR_COUNTERS, 0, 1)?[0])
    }
    pub fn classify_counters(&self) -> Result<GpuTransvoxelClassifyCounters, TransvoxelGpuError> {
        Ok(self.ctx.read(ffi::HVX_BUF_REGULAR_CLASSIFY, 0, 1)?[0])
    }
    pub fn vertices(&self, count: usize) -> Result<Vec<GpuTerrainVertex>, TransvoxelGpuError> { self.ctx.read(ffi::HVX_BUF_REGULAR_VERTICES, 0, count) }
    pub fn indices(&self, count: usize) -> Result<Vec<u32>, TransvoxelGpuError> { self.ctx.read(ffi::HVX_BUF_REGULAR_INDICES, 0, count) }
    /// `output_buffer()` / `offsets_buffer()` / `blocks_buffer()` of the reference, as host copies.
    pub fn cells(&self) -> Result<Vec<GpuTransvoxelCell>, TransvoxelGpuError> {
        if !self.ctx.has_records() { return Err(TransvoxelGpuError::RecordsDisabled); }
        self.ctx.read(ffi::HVX_BUF_REGULAR_CELLS, 0, PAGE_CELLS)
    }
    pub fn offsets(&self) -> Result<Vec<GpuTransvoxelCellOffset>, TransvoxelGpuError> {
        if !self.ctx.has_records() { return Err(TransvoxelGpuError::RecordsDisabled); }
        self.ctx.read(ffi::HVX_BUF_REGULAR_OFFSETS, 0, PAGE_CELLS)
    }
    pub fn blocks(&self) -> Result<Vec<GpuTransvoxelScanBlock>, TransvoxelGpuError> {
        if !self.ctx.has_records() { return Err(TransvoxelGpuError::RecordsDisabled); }
        self.ctx.read(ffi::HVX_BUF_REGULAR_BLOCKS, 0, PAGE_CELLS / 256)
    }
    /// Device pointers of the arenas, for CUDA<->Vulkan external-memory interop with wgpu.
    pub fn vertices_device_ptr(&self) -> *const c_void { unsafe { ffi::hvx_buffer(self.ctx.raw, ffi::HVX_BUF_REGULAR_VERTICES) } }
    pub fn indices_device_ptr(&self) -> *const c_void { unsafe { ffi::hvx_buffer(self.ctx.raw, ffi::HVX_BUF_REGULAR_INDICES) } }
    pub const fn config(&self) -> TransvoxelGpuExtractorConfig { self.config }
    pub fn allocated_bytes(&self) -> u64 { unsafe { ffi::hvx_allocated_bytes(self.ctx.raw) } }
    pub fn resize(&mut self, _width: u32, _height: u32) {}
}

/// Drop-in for `TransvoxelGpuClassifier` (PV/src/transvoxel_gpu.rs:148-356): classification and counters, no emission.
pub struct TransvoxelGpuClassifier { ctx: Ctx }

impl TransvoxelGpuClassifier {
    pub fn new(cuda_device: i32) -> Result<Self, TransvoxelGpuError> {
        let raw = ffi::hvx_config { edge: 32, max_chunks: 1, max_vertices: 1, max_indices: 1,
            flags: ffi::HVX_CFG_DEBUG_RECORDS, ..Default::default() };
        Ok(Self { ctx: Ctx::new(cuda_device, raw)? })
    }
    /// PV/src/transvoxel_gpu.rs:268-286
    pub fn dispatch(&self, samples: &[CellWord], generation: u64, dirty_microbricks: u64) -> Result<(), TransvoxelGpuError> {
        let d = desc(generation, dirty_microbricks, 0);
        let words: &[u32] = bytemuck::cast_slice(samples);
        let status = unsafe { ffi::hvx_classify_regular(self.ctx.raw, words.as_ptr(), words.len() as u64, &d, 1) };
        self.ctx.check(status, samples.len(), PAGE_SAMPLES, 0)
    }
    pub fn output(&self) -> Result<Vec<GpuTransvoxelCell>, TransvoxelGpuError> { self.ctx.read(ffi::HVX_BUF_REGULAR_CELLS, 0, PAGE_CELLS) }
    pub fn counters(&self) -> Result<GpuTransvoxelClassifyCounters, TransvoxelGpuError> { Ok(self.ctx.read(ffi::HVX_BUF_REGULAR_CLASSIFY, 0, 1)?[0]) }
}

/// Drop-in for `TransvoxelGpuTransitionExtractor` (PV/src/transvoxel_transition_gpu.rs:190-520).
pub struct TransvoxelGpuTransitionExtractor { ctx: Ctx, config: TransvoxelGpuExtractorConfig }

impl TransvoxelGpuTransitionExtractor {
    pub fn new(cuda_device: i32, config: TransvoxelGpuExtractorConfig) -> Result<Self, TransvoxelGpuError> {
        let config = TransvoxelGpuExtractorConfig::new(config.max_vertices, config.max_indices)?;
        let raw = ffi::hvx_config { edge: 32, max_chunks: 1, max_vertices: 1, max_indices: 1,
            max_transition_vertices: config.max_vertices, max_transition_indices: config.max_indices, ..Default::default() };
        Ok(Self { ctx: Ctx::new(cuda_device, raw)?, config })
    }

    /// PV/src/transvoxel_transition_gpu.rs:366-380 (mask before generation, like the reference)
    pub fn dispatch(&self, face_slabs: &[CellWord], transition_mask: u8, generation: u64) -> Result<(), TransvoxelGpuError> {
        let d = desc(generation, u64::MAX, transition_mask);
        let words: &[u32] = bytemuck::cast_slice(face_slabs);
        let status = unsafe { ffi::hvx_extract_transition(self.ctx.raw, words.as_ptr(), words.len() as u64, &d, 1) };
        self.ctx.check(status, face_slabs.len(), SLAB_SAMPLES, transition_mask)
    }
    pub fn counters(&self) -> Result<GpuTransvoxelTransitionCounters, TransvoxelGpuError> {
        Ok(self.ctx.read(ffi::HVX_BUF_TRANSITION_COUNTERS, 0, 1)?[0])
    }
    pub fn vertices(&self, count: usize) -> Result<Vec<GpuTerrainVertex>, TransvoxelGpuError> { self.ctx.read(ffi::HVX_BUF_TRANSITION_VERTICES, 0, count) }
    pub fn indices(&self, count: usize) -> Result<Vec<u32>, TransvoxelGpuError> { self.ctx.read(ffi::HVX_BUF_TRANSITION_INDICES, 0, count) }
    pub const fn config(&self) -> TransvoxelGpuExtractorConfig { self.config }
    pub fn resize(&mut self, _width: u32, _height: u32) {}
}

/// Per-chunk dispatch parameters of a batch (`GpuTransvoxelDispatch`, PV/src/transvoxel_gpu.rs:15-32, per page).
#[derive(Clone, Copy, Debug, Default, PartialEq, Eq)]
pub struct ChunkRequest {
    pub generation: u64,
    pub dirty_microbricks: u64,
    pub transition_mask: u8,
    /// scheduler input (e.g. the chunk's vertex count last time): the batch starts its heaviest chunks first
    pub cost_hint: u32,
    /// the producer knows the chunk holds no surface: not uploaded, not read (`HVX_CHUNK_UNIFORM`)
    pub uniform: bool,
}

/// A batch staged by [`ChunkBatchExtractor::prepare`], queued by `encode` (PV/src/transvoxel_emit.rs:256-330).
pub struct PreparedBatch { descs: Vec<ffi::hvx_chunk_desc> }
impl PreparedBatch { pub fn len(&self) -> usize { self.descs.len() } pub fn is_empty(&self) -> bool { self.descs.is_empty() } }

/// Packed host copy of a batch's meshes: vertices / indices back to back in chunk order, index values chunk-local.
#[derive(Clone, Debug, Default)]
pub struct BatchMeshes {
    pub vertices: Vec<GpuTerrainVertex>,
    pub indices: Vec<u32>,
    pub ranges: Vec<ffi::hvx_range>,
    pub counters: Vec<GpuTransvoxelEmissionCounters>,
}

/// N chunks per dispatch into fixed-stride per-chunk slots, edge 32 (the reference's page) or 64.
pub struct ChunkBatchExtractor { ctx: Ctx }

impl ChunkBatchExtractor {
    pub fn new(cuda_device: i32, edge: u32, max_chunks: u32, regular: TransvoxelGpuExtractorConfig,
               transition: Option<TransvoxelGpuExtractorConfig>) -> Result<Self, TransvoxelGpuError> {
        let t = transition.unwrap_or(TransvoxelGpuExtractorConfig { max_vertices: 0, max_indices: 0 });
        let raw = ffi::hvx_config { edge, max_chunks, max_vertices: regular.max_vertices, max_indices: regular.max_indices,
            max_transition_vertices: t.max_vertices, max_transition_indices: t.max_indices, ..Default::default() };
        Ok(Self { ctx: Ctx::new(cuda_device, raw)? })
    }
    pub fn edge(&self) -> u32 { self.ctx.config.edge }
    fn sample_words(&self) -> usize { let s = self.ctx.config.edge as usize + 2; s * s * s }

    /// Validation and descriptor staging, no device work (the reference's `prepare`).
    pub fn prepare(&self, requests: &[ChunkRequest]) -> Result<PreparedBatch, TransvoxelGpuError> {
        if requests.len() > self.ctx.config.max_chunks as usize {
            return Err(TransvoxelGpuError::BatchCapacity { requested: requests.len(), capacity: self.ctx.config.max_chunks });
        }
        let mut descs = Vec::with_capacity(requests.len());
        for r in requests {
            if r.transition_mask & !0x3f != 0 { return Err(TransvoxelGpuError::TransitionMask(r.transition_mask)); }
            descs.push(ffi::hvx_chunk_desc { generation: r.generation, dirty_microbricks: r.dirty_microbricks,
                transition_mask: r.transition_mask as u32, cost_hint: r.cost_hint,
                flags: if r.uniform { ffi::HVX_CHUNK_UNIFORM } else { 0 }, _reserved: 0 });
        }
        Ok(PreparedBatch { descs })
    }
    /// Queue the regular extraction of a prepared batch (the reference's `encode`).  `samples`: host CellWords
    /// (uploaded in sub-batches beside the kernels) or `None` for the ctx sample arena (after `fill` / `gather`).
    pub fn encode(&self, batch: &PreparedBatch, samples: Option<&[CellWord]>) -> Result<(), TransvoxelGpuError> {
        let n = batch.descs.len();
        let expected = n * self.sample_words();
        let (ptr, words) = match samples { Some(s) => { let w: &[u32] = bytemuck::cast_slice(s); (w.as_ptr(), w.len()) } None => (core::ptr::null(), expected) };
        let status = unsafe { ffi::hvx_extract_regular(self.ctx.raw, ptr, words as u64, batch.descs.as_ptr(), n as u32) };
        self.ctx.check(status, words, expected, 0)
    }
    /// `prepare` + `encode` + the packed read-back in ONE pipelined call (hvx_extract_regular_to_host).
    pub fn extract_to_host(&self, batch: &PreparedBatch, samples: Option<&[CellWord]>, vertex_capacity: usize, index_capacity: usize)
        -> Result<BatchMeshes, TransvoxelGpuError> {
        let n = batch.descs.len();
        let expected = n * self.sample_words();
        let (ptr, words) = match samples { Some(s) => { let w: &[u32] = bytemuck::cast_slice(s); (w.as_ptr(), w.len()) } None => (core::ptr::null(), expected) };
        let mut out = BatchMeshes { vertices: vec![GpuTerrainVertex::zeroed(); vertex_capacity], indices: vec![0u32; index_capacity],
            ranges: vec![ffi::hvx_range::default(); n], counters: vec![GpuTransvoxelEmissionCounters::zeroed(); n] };
        let (mut tv, mut ti) = (0u64, 0u64);
        let status = unsafe { ffi::hvx_extract_regular_to_host(self.ctx.raw, ptr, words as u64, batch.descs.as_ptr(), n as u32,
            out.vertices.as_mut_ptr() as *mut ffi::hvx_vertex, vertex_capacity as u64, out.indices.as_mut_ptr(), index_capacity as u64,
            out.ranges.as_mut_ptr(), out.counters.as_mut_ptr() as *mut ffi::hvx_emission_counters, &mut tv, &mut ti) };
        self.ctx.check(status, words, expected, 0)?;
        out.vertices.truncate(tv as usize);
        out.indices.truncate(ti as usize);
        Ok(out)
    }
    /// Transition extraction of chunks `[0, n)`; `slabs`: host CellWords or `None` for the ctx slab arena.
    pub fn encode_transition(&self, batch: &PreparedBatch, slabs: Option<&[CellWord]>) -> Result<(), TransvoxelGpuError> {
        let n = batch.descs.len();
        let w = 2 * self.ctx.config.edge as usize + 3;
        let expected = n * 18 * w * w;
        let (ptr, words) = match slabs { Some(s) => { let v: &[u32] = bytemuck::cast_slice(s); (v.as_ptr(), v.len()) } None => (core::ptr::null(), expected) };
        let status = unsafe { ffi::hvx_extract_transition(self.ctx.raw, ptr, words as u64, batch.descs.as_ptr(), n as u32) };
        let mask = batch.descs.iter().map(|d| d.transition_mask).find(|m| m & !0x3f != 0).unwrap_or(0) as u8;
        self.ctx.check(status, words, expected, mask)
    }
    /// ExtractionFixture::new on the device (PV/src/fixture.rs:95-124) for `pages` = `[x, y, z]` per chunk.
    pub fn fill(&self, kind: u32, pages: &[[i64; 3]], lods: Option<&[u8]>) -> Result<(), TransvoxelGpuError> {
        let status = unsafe { ffi::hvx_fill_density(self.ctx.raw, kind, pages.as_ptr() as *const i64,
            lods.map_or(core::ptr::null(), |l| l.as_ptr()), pages.len() as u32, core::ptr::null_mut()) };
        self.ctx.check(status, pages.len(), 0, 0)
    }
    /// 63-index meshlets + bounds over the last extraction (PV/src/terrain_meshlet_build.wgsl:205-261).
    pub fn build_meshlets(&self, n: u32, transition: bool) -> Result<(), TransvoxelGpuError> {
        let status = unsafe { ffi::hvx_build_meshlets(self.ctx.raw, transition as c_int, n) };
        self.ctx.check(status, n as usize, 0, 0)
    }
    /// One edit of the legacy edit ring (`GpuVoxelEdit`, crates/helio-voxel-core/src/gpu_types.rs:47-54) applied to the
    /// resident samples of `pages`; returns each chunk's dirty microbricks (the mask `prepare` takes) and how many chunks
    /// the edit touched (sphere-vs-box rule of `OctreeNode::mark_sphere_dirty`, octree.rs:139-173).
    pub fn apply_edit(&self, edit: &ffi::hvx_voxel_edit, pages: &[[i64; 3]], lods: Option<&[u8]>) -> Result<(Vec<u64>, u32), TransvoxelGpuError> {
        let mut dirty = vec![0u64; pages.len()];
        let mut touched = 0u32;
        let lod_ptr = lods.map_or(core::ptr::null(), |l| l.as_ptr());
        let status = unsafe { ffi::hvx_apply_edit(self.ctx.raw, edit, pages.as_ptr() as *const i64, lod_ptr, pages.len() as u32,
                                                   core::ptr::null_mut(), dirty.as_mut_ptr(), &mut touched) };
        self.ctx.check(status, pages.len(), 0, 0).map(|_| (dirty, touched))
    }
    /// Optional vertex-reuse output (off by default; the reference shares no vertices): merges the bit-identical vertex
    /// records of chunks `[0, n)` of the last extraction in place, indices follow, ranges / emitted_vertices shrink.
    pub fn weld_meshes(&self, n: u32, transition: bool) -> Result<(), TransvoxelGpuError> {
        let status = unsafe { ffi::hvx_weld_meshes(self.ctx.raw, transition as c_int, n) };
        self.ctx.check(status, n as usize, 0, 0)
    }
    pub fn counters(&self, n: usize) -> Result<Vec<GpuTransvoxelEmissionCounters>, TransvoxelGpuError> { self.ctx.read(ffi::HVX_BUF_REGULAR_COUNTERS, 0, n) }
    pub fn transition_counters(&self, n: usize) -> Result<Vec<GpuTransvoxelTransitionCounters>, TransvoxelGpuError> { self.ctx.read(ffi::HVX_BUF_TRANSITION_COUNTERS, 0, n) }
    pub fn ranges(&self, n: usize) -> Result<Vec<ffi::hvx_range>, TransvoxelGpuError> {
        let mut out = vec![ffi::hvx_range::default(); n];
        let status = unsafe { ffi::hvx_read(self.ctx.raw, ffi::HVX_BUF_REGULAR_RANGES, 0, (n * 16) as u64, out.as_mut_ptr() as *mut c_void) };
        self.ctx.check(status, 0, 0, 0).map(|_| out)
    }
    pub fn synchronize(&self) -> Result<(), TransvoxelGpuError> { self.ctx.check(unsafe { ffi::hvx_synchronize(self.ctx.raw) }, 0, 0, 0) }
    /// Device pointer of an arena (an `HVX_BUF_` id of the `ffi` module), for external-memory interop.
    pub fn device_ptr(&self, buffer: c_int) -> *mut c_void { unsafe { ffi::hvx_buffer(self.ctx.raw, buffer) } }
}

/// `GpuSurfaceSampler` (PV/src/surface_sampling.rs:309-337): halo blocks + fine-side slabs of a batch of jobs, gathered
/// from the resident page atlas into the batch extractor's arenas (then `encode(.., None)`).
pub struct GpuSurfaceSampler<'a> { batch: &'a ChunkBatchExtractor }
impl<'a> GpuSurfaceSampler<'a> {
    pub fn new(batch: &'a ChunkBatchExtractor) -> Self { Self { batch } }
    /// `table` / `atlas`: host slices or device memory (detected by the library); `jobs`: host.
    pub fn dispatch(&self, residency: &ffi::hvx_residency, table: &[ffi::hvx_page_table_entry], atlas: *const u32, atlas_words: u64,
                    jobs: &[ffi::hvx_gather_job]) -> Result<(), TransvoxelGpuError> {
        let table_ptr = if table.is_empty() { core::ptr::null() } else { table.as_ptr() };  // empty: the table bound with bind_page_table
        let status = unsafe { ffi::hvx_gather_surface(self.batch.ctx.raw, residency, table_ptr, atlas, atlas_words, jobs.as_ptr(), jobs.len() as u32) };
        self.batch.ctx.check(status, jobs.len(), 0, 0)
    }
    /// Upload the residency layer's page table once per publication (PV/src/table.rs:62-72); `dispatch` then takes an
    /// empty `table` slice and reads the bound one.  An empty slice here unbinds.
    pub fn bind_page_table(&self, table: &[ffi::hvx_page_table_entry]) -> Result<(), TransvoxelGpuError> {
        let ptr = if table.is_empty() { core::ptr::null() } else { table.as_ptr() };
        let status = unsafe { ffi::hvx_gather_bind_table(self.batch.ctx.raw, ptr, table.len() as u32) };
        self.batch.ctx.check(status, table.len(), 0, 0)
    }
    pub fn counters(&self, n: usize) -> Result<Vec<ffi::hvx_gather_counters>, TransvoxelGpuError> {
        let mut out = vec![ffi::hvx_gather_counters::default(); n];
        let status = unsafe { ffi::hvx_read(self.batch.ctx.raw, ffi::HVX_BUF_GATHER_COUNTERS, 0, (n * 32) as u64, out.as_mut_ptr() as *mut c_void) };
        self.batch.ctx.check(status, 0, 0, 0).map(|_| out)
    }
}

/// The double-banked per-slot arenas, surface states, indirect draws and feedback of the render pass
/// (PV/src/surface_publish.wgsl:104-225, PODs PV/src/render.rs:466-554).
pub struct SurfacePublisher<'a> { raw: *mut ffi::hvx_publisher, batch: &'a ChunkBatchExtractor }
impl<'a> SurfacePublisher<'a> {
    pub fn new(batch: &'a ChunkBatchExtractor, slots: u32) -> Result<Self, TransvoxelGpuError> {
        let mut raw = core::ptr::null_mut();
        let status = unsafe { ffi::hvx_publisher_create(batch.ctx.raw, slots, &mut raw) };
        batch.ctx.check(status, slots as usize, 0, 0).map(|_| Self { raw, batch })
    }
    /// copy_regular_surface + copy_transition_surface + publish_surface for a batch of jobs (distinct slots).
    pub fn publish(&self, jobs: &[ffi::hvx_surface_job], job_chunk: &[u32], page_metadata: &[ffi::hvx_page_meta]) -> Result<(), TransvoxelGpuError> {
        assert_eq!(jobs.len(), job_chunk.len());
        let status = unsafe { ffi::hvx_publish_surfaces(self.raw, jobs.as_ptr(), job_chunk.as_ptr(), page_metadata.as_ptr(), jobs.len() as u32) };
        self.batch.ctx.check(status, jobs.len(), 0, 0)
    }
    pub fn refresh_visibility(&self, draw_pages: &[ffi::hvx_draw_page]) -> Result<(), TransvoxelGpuError> {
        self.batch.ctx.check(unsafe { ffi::hvx_refresh_visibility(self.raw, draw_pages.as_ptr()) }, draw_pages.len(), 0, 0)
    }
    pub fn feedback(&self) -> Result<ffi::hvx_surface_feedback, TransvoxelGpuError> {
        let mut out = ffi::hvx_surface_feedback::default();
        let status = unsafe { ffi::hvx_publisher_read(self.raw, ffi::HVX_PUB_FEEDBACK, 0, 32, &mut out as *mut _ as *mut c_void) };
        self.batch.ctx.check(status, 0, 0, 0).map(|_| out)
    }
    /// Device pointer of a publisher buffer (an `HVX_PUB_` id of the `ffi` module): what wgpu imports as vertex / index / indirect buffers.
    pub fn device_ptr(&self, buffer: c_int) -> *mut c_void { unsafe { ffi::hvx_publisher_buffer(self.raw, buffer) } }
}
impl Drop for SurfacePublisher<'_> { fn drop(&mut self) { unsafe { ffi::hvx_publisher_destroy(self.raw) } } }
