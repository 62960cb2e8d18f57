//! Bounded extraction publisher over the C ABI (include/hvx.h, "bounded extraction publisher").
//! Same surface as `helio_pass_planetary_voxel::extraction::BoundedExtractionPublisher`
//! (PV/src/extraction.rs:342-603); the bookkeeping runs in the library's host C++, and
//! `attach` / `commit` add the device step (copy into the reserved ranges of the bounded arenas).
//! The PODs and prototypes come from the generated `ffi` module.
//! Source only: this image has no Rust toolchain.
use core::ffi::c_int;

pub use crate::ffi::{
    hvx_arena_slice, hvx_evict_outcome, hvx_extraction_limits, hvx_extraction_plan, hvx_extraction_publisher,
    hvx_extraction_publisher_counters, hvx_planet_page_key, hvx_publication_outcome, hvx_published_surface, hvx_reservation,
    hvx_reservation_outcome, hvx_surface_allocation, hvx_surface_counts,
};
use crate::ffi::{
    hvx_ctx, hvx_extraction_cancel_pending, hvx_extraction_commit, hvx_extraction_evict, hvx_extraction_publish,
    hvx_extraction_publisher_attach, hvx_extraction_publisher_create, hvx_extraction_publisher_destroy,
    hvx_extraction_publisher_get_counters, hvx_extraction_reserve,
};

/// Owns the C handle; method names follow the reference (`reserve`, `publish`, `cancel_pending`, `evict`).
pub struct BoundedExtractionPublisher(*mut hvx_extraction_publisher);
unsafe impl Send for BoundedExtractionPublisher {}

impl BoundedExtractionPublisher {
    pub fn new(limits: hvx_extraction_limits) -> Result<Self, c_int> {
        let mut raw = core::ptr::null_mut();
        match unsafe { hvx_extraction_publisher_create(&limits, &mut raw) } { 0 => Ok(Self(raw)), status => Err(status) }
    }
    pub fn reserve(&mut self, key: hvx_planet_page_key, generation: u64, counts: hvx_surface_counts) -> Result<hvx_reservation_outcome, (c_int, u32)> {
        let mut out = hvx_reservation_outcome::default();
        match unsafe { hvx_extraction_reserve(self.0, &key, generation, &counts, &mut out) } { 0 => Ok(out), status => Err((status, out.detail)) }
    }
    pub fn publish(&mut self, reservation: hvx_reservation) -> Result<hvx_publication_outcome, c_int> {
        let mut out = hvx_publication_outcome::default();
        match unsafe { hvx_extraction_publish(self.0, &reservation, &mut out) } { 0 => Ok(out), status => Err(status) }
    }
    pub fn cancel_pending(&mut self, key: hvx_planet_page_key, generation: u64) -> Result<bool, c_int> {
        let mut cancelled = 0;
        match unsafe { hvx_extraction_cancel_pending(self.0, &key, generation, &mut cancelled) } { 0 => Ok(cancelled != 0), status => Err(status) }
    }
    pub fn evict(&mut self, key: hvx_planet_page_key, generation: u64) -> hvx_evict_outcome {
        let mut out = hvx_evict_outcome::default();
        unsafe { hvx_extraction_evict(self.0, &key, generation, &mut out) };
        out
    }
    pub fn counters(&self) -> hvx_extraction_publisher_counters {
        let mut out = hvx_extraction_publisher_counters::default();
        unsafe { hvx_extraction_publisher_get_counters(self.0, &mut out) };
        out
    }
    /// Device side: bounded arenas on `ctx`'s GPU, then one launch per batch of reserved pages.
    pub unsafe fn attach(&mut self, ctx: *mut hvx_ctx) -> c_int { hvx_extraction_publisher_attach(self.0, ctx) }
    pub fn commit(&mut self, chunk: &[u32], page_slot: &[u32], reservations: &[hvx_reservation]) -> c_int {
        assert!(chunk.len() == reservations.len() && page_slot.len() == reservations.len());
        unsafe { hvx_extraction_commit(self.0, chunk.as_ptr(), page_slot.as_ptr(), reservations.as_ptr(), reservations.len() as u32) }
    }
}

impl Drop for BoundedExtractionPublisher {
    fn drop(&mut self) { unsafe { hvx_extraction_publisher_destroy(self.0) } }
}
