//! Safe wrapper with the reference's names: swap
//! `helio_pass_planetary_voxel::{TransvoxelGpuExtractor, TransvoxelGpuTransitionExtractor}`
//! (PV/src/render.rs:579-580 holds the two fields) for the types below.  Where the reference takes
//! `(&wgpu::Device, &wgpu::Queue)` these take a CUDA device ordinal at construction.
pub mod ffi;
pub mod extraction;

use bytemuck::{Pod, Zeroable};
use core::ffi::{c_int, c_void, CStr};
use helio_planet_voxel_core::CellWord;

pub use helio_pass_planetary_voxel_types::*;

/// Re-declared PODs: byte-identical to PV/src/extraction.rs:72-79 and PV/src/transvoxel_emit.rs:38-48,
/// so a Helio build can instead `pub use helio_pass_planetary_voxel::{GpuTerrainVertex, ...}`.
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

struct Ctx(*mut ffi::hvx_ctx);
unsafe impl Send for Ctx {}
unsafe impl Sync for Ctx {} // calls are serialised by `&mut self` / the pass's single queue, like the reference
impl Drop for Ctx { fn drop(&mut self) { unsafe { ffi::hvx_destroy(self.0) } } }

fn error(ctx: *const ffi::hvx_ctx, status: c_int, samples: usize, expected: usize, mask: u8,
         config: TransvoxelGpuExtractorConfig) -> TransvoxelGpuError {
    let message = unsafe { CStr::from_ptr(ffi::hvx_last_error(ctx)) }.to_string_lossy().into_owned();
    match status {
        ffi::HVX_E_SAMPLE_COUNT => TransvoxelGpuError::SampleCount { actual: samples, expected },
        ffi::HVX_E_INVALID_CAPACITY => TransvoxelGpuError::InvalidExtractionCapacity {
            max_vertices: config.max_vertices, max_indices: config.max_indices },
        ffi::HVX_E_DEVICE_LIMIT => TransvoxelGpuError::DeviceLimit(message),
        ffi::HVX_E_TRANSITION_MASK => TransvoxelGpuError::TransitionMask(mask),
        _ => TransvoxelGpuError::Backend { status, message },
    }
}

/// Drop-in for `helio_pass_planetary_voxel::TransvoxelGpuExtractor` (PV/src/transvoxel_emit.rs:92-396).
pub struct TransvoxelGpuExtractor { ctx: Ctx, config: TransvoxelGpuExtractorConfig }

impl TransvoxelGpuExtractor {
    pub fn new(cuda_device: i32, config: TransvoxelGpuExtractorConfig) -> Result<Self, TransvoxelGpuError> {
        let config = TransvoxelGpuExtractorConfig::new(config.max_vertices, config.max_indices)?;
        let raw = ffi::hvx_config { edge: 32, max_chunks: 1, max_vertices: config.max_vertices,
            max_indices: config.max_indices, flags: ffi::HVX_CFG_DEBUG_RECORDS, ..Default::default() };
        let mut ctx = core::ptr::null_mut();
        let status = unsafe { ffi::hvx_create(&mut ctx, cuda_device, &raw) };
        if status != ffi::HVX_OK { return Err(error(core::ptr::null(), status, 0, 0, 0, config)); }
        Ok(Self { ctx: Ctx(ctx), config })
    }

    /// PV/src/transvoxel_emit.rs:233-254
    pub fn dispatch(&self, samples: &[CellWord], generation: u64, dirty_microbricks: u64, transition_mask: u8)
        -> Result<(), TransvoxelGpuError> {
        let desc = ffi::hvx_chunk_desc { generation, dirty_microbricks, transition_mask: transition_mask as u32, cost_hint: 0 };
        let words: &[u32] = bytemuck::cast_slice(samples);
        let status = unsafe { ffi::hvx_extract_regular(self.ctx.0, words.as_ptr(), words.len() as u64, &desc, 1) };
        if status != ffi::HVX_OK { return Err(error(self.ctx.0, status, samples.len(), 34 * 34 * 34, transition_mask, self.config)); }
        Ok(())
    }

    fn read<T: Pod>(&self, buffer: c_int, count: usize) -> Vec<T> {
        let mut out = vec![T::zeroed(); count];
        let bytes = (count * core::mem::size_of::<T>()) as u64;
        let status = unsafe { ffi::hvx_read(self.ctx.0, buffer, 0, bytes, out.as_mut_ptr() as *mut c_void) };
        assert_eq!(status, ffi::HVX_OK, "hvx_read failed");
        out
    }
    pub fn counters(&self) -> GpuTransvoxelEmissionCounters { self.read(ffi::HVX_BUF_REGULAR_COUNTERS, 1)[0] }
    pub fn vertices(&self, count: usize) -> Vec<GpuTerrainVertex> { self.read(ffi::HVX_BUF_REGULAR_VERTICES, count) }
    pub fn indices(&self, count: usize) -> Vec<u32> { self.read(ffi::HVX_BUF_REGULAR_INDICES, count) }
    /// Device pointers of the arenas, for CUDA<->Vulkan external-memory interop with wgpu.
    pub fn vertices_device_ptr(&self) -> *const c_void { unsafe { ffi::hvx_buffer(self.ctx.0, ffi::HVX_BUF_REGULAR_VERTICES) } }
    pub fn indices_device_ptr(&self) -> *const c_void { unsafe { ffi::hvx_buffer(self.ctx.0, ffi::HVX_BUF_REGULAR_INDICES) } }
    pub const fn config(&self) -> TransvoxelGpuExtractorConfig { self.config }
    pub fn allocated_bytes(&self) -> u64 { unsafe { ffi::hvx_allocated_bytes(self.ctx.0) } }
    pub fn resize(&mut self, _width: u32, _height: u32) {}
}

/// Drop-in for `TransvoxelGpuTransitionExtractor` (PV/src/transvoxel_transition_gpu.rs:190-520).
pub struct TransvoxelGpuTransitionExtractor { ctx: Ctx, config: TransvoxelGpuExtractorConfig }

impl TransvoxelGpuTransitionExtractor {
    pub fn new(cuda_device: i32, config: TransvoxelGpuExtractorConfig) -> Result<Self, TransvoxelGpuError> {
        let config = TransvoxelGpuExtractorConfig::new(config.max_vertices, config.max_indices)?;
        let raw = ffi::hvx_config { edge: 32, max_chunks: 1, max_vertices: 1, max_indices: 1,
            max_transition_vertices: config.max_vertices, max_transition_indices: config.max_indices,
            flags: ffi::HVX_CFG_DEBUG_RECORDS, ..Default::default() };
        let mut ctx = core::ptr::null_mut();
        let status = unsafe { ffi::hvx_create(&mut ctx, cuda_device, &raw) };
        if status != ffi::HVX_OK { return Err(error(core::ptr::null(), status, 0, 0, 0, config)); }
        Ok(Self { ctx: Ctx(ctx), config })
    }

    /// PV/src/transvoxel_transition_gpu.rs:366-380 (mask before generation, like the reference)
    pub fn dispatch(&self, face_slabs: &[CellWord], transition_mask: u8, generation: u64) -> Result<(), TransvoxelGpuError> {
        let desc = ffi::hvx_chunk_desc { generation, dirty_microbricks: u64::MAX, transition_mask: transition_mask as u32, cost_hint: 0 };
        let words: &[u32] = bytemuck::cast_slice(face_slabs);
        let status = unsafe { ffi::hvx_extract_transition(self.ctx.0, words.as_ptr(), words.len() as u64, &desc, 1) };
        if status != ffi::HVX_OK { return Err(error(self.ctx.0, status, face_slabs.len(), 6 * 3 * 67 * 67, transition_mask, self.config)); }
        Ok(())
    }
    pub fn counters(&self) -> GpuTransvoxelTransitionCounters {
        let mut out = GpuTransvoxelTransitionCounters::default();
        let status = unsafe { ffi::hvx_read(self.ctx.0, ffi::HVX_BUF_TRANSITION_COUNTERS, 0, 48, &mut out as *mut _ as *mut c_void) };
        assert_eq!(status, ffi::HVX_OK);
        out
    }
    pub const fn config(&self) -> TransvoxelGpuExtractorConfig { self.config }
    pub fn resize(&mut self, _width: u32, _height: u32) {}
}
