//! Bounded extraction publisher over the C ABI (include/hvx.h, "bounded extraction publisher").
//! Same surface as `helio_pass_planetary_voxel::extraction::BoundedExtractionPublisher`
//! (PV/src/extraction.rs:342-603); the bookkeeping runs in the library's host C++, and
//! `attach` / `commit` add the device step (copy into the reserved ranges of the bounded arenas).
//! Source only: this image has no Rust toolchain.
use core::ffi::{c_char, c_int, c_void};

use crate::ffi::hvx_ctx;

#[repr(C)] #[derive(Clone, Copy, Default, PartialEq, Eq, Debug)]
pub struct hvx_extraction_limits { pub max_page_slots: u32, pub max_pending_pages: u32, pub max_vertices: u32, pub max_indices: u32, pub max_meshlets: u32 }
#[repr(C)] #[derive(Clone, Copy, Default, PartialEq, Eq, Debug)]
pub struct hvx_extraction_plan { pub request_bytes: u64, pub page_range_bytes: u64, pub vertex_bytes: u64, pub index_bytes: u64, pub meshlet_bytes: u64, pub counter_bytes: u64, pub total_bytes: u64 }
#[repr(C)] #[derive(Clone, Copy, Default, PartialEq, Eq, Debug)]
pub struct hvx_planet_page_key { pub planet_id: [u8; 16], pub page_xyz: [i64; 3], pub lod: u8, pub _pad: [u8; 7] }
#[repr(C)] #[derive(Clone, Copy, Default, PartialEq, Eq, Debug)]
pub struct hvx_surface_counts { pub vertices: u32, pub indices: u32, pub meshlets: u32 }
#[repr(C)] #[derive(Clone, Copy, Default, PartialEq, Eq, Debug)]
pub struct hvx_arena_slice { pub first: u32, pub count: u32 }
#[repr(C)] #[derive(Clone, Copy, Default, PartialEq, Eq, Debug)]
pub struct hvx_surface_allocation { pub vertices: hvx_arena_slice, pub indices: hvx_arena_slice, pub meshlets: hvx_arena_slice }
#[repr(C)] #[derive(Clone, Copy, Default, PartialEq, Eq, Debug)]
pub struct hvx_reservation { pub key: hvx_planet_page_key, pub generation: u64, pub allocation: hvx_surface_allocation }
#[repr(C)] #[derive(Clone, Copy, Default, PartialEq, Eq, Debug)]
pub struct hvx_published_surface { pub generation: u64, pub allocation: hvx_surface_allocation }
#[repr(C)] #[derive(Clone, Copy, Default, Debug)]
pub struct hvx_reservation_outcome { pub kind: u32, pub detail: u32, pub reservation: hvx_reservation, pub current: hvx_published_surface, pub newest_generation: u64 }
#[repr(C)] #[derive(Clone, Copy, Default, Debug)]
pub struct hvx_publication_outcome { pub kind: u32, pub has_replaced: u32, pub current: hvx_published_surface, pub replaced: hvx_published_surface, pub newest_generation: u64 }
#[repr(C)] #[derive(Clone, Copy, Default, Debug)]
pub struct hvx_evict_outcome { pub kind: u32, pub _pad: u32, pub newest_generation: u64 }
#[repr(C)] #[derive(Clone, Copy, Default, PartialEq, Eq, Debug)]
pub struct hvx_extraction_publisher_counters {
    pub current_pages: u64, pub pending_pages: u64, pub used_vertices: u32, pub used_indices: u32, pub used_meshlets: u32, pub _pad0: u32,
    pub pending_high_water: u64, pub vertex_high_water: u32, pub index_high_water: u32, pub meshlet_high_water: u32, pub _pad1: u32,
    pub reservations: u64, pub publications: u64, pub replacements: u64, pub cancellations: u64, pub evictions: u64,
    pub stale_rejected: u64, pub backpressured: u64,
}
#[repr(C)] pub struct hvx_extraction_publisher { _private: [u8; 0] }

pub const HVX_E_PENDING_CAPACITY: c_int = -44;
pub const HVX_E_ARENA_CAPACITY: c_int = -45;

extern "C" {
    pub fn hvx_extraction_limits_plan(limits: *const hvx_extraction_limits, plan_out: *mut hvx_extraction_plan) -> c_int;
    pub fn hvx_extraction_limits_validate_device(limits: *const hvx_extraction_limits, max_buffer_size: u64,
        max_storage_buffer_binding_size: u64, name_out: *mut *const c_char, requested_out: *mut u64) -> c_int;
    pub fn hvx_extraction_publisher_create(limits: *const hvx_extraction_limits, out: *mut *mut hvx_extraction_publisher) -> c_int;
    pub fn hvx_extraction_publisher_destroy(publisher: *mut hvx_extraction_publisher);
    pub fn hvx_extraction_reserve(publisher: *mut hvx_extraction_publisher, key: *const hvx_planet_page_key, generation: u64,
        counts: *const hvx_surface_counts, out: *mut hvx_reservation_outcome) -> c_int;
    pub fn hvx_extraction_publish(publisher: *mut hvx_extraction_publisher, reservation: *const hvx_reservation,
        out: *mut hvx_publication_outcome) -> c_int;
    pub fn hvx_extraction_cancel_pending(publisher: *mut hvx_extraction_publisher, key: *const hvx_planet_page_key,
        generation: u64, cancelled_out: *mut c_int) -> c_int;
    pub fn hvx_extraction_evict(publisher: *mut hvx_extraction_publisher, key: *const hvx_planet_page_key, generation: u64,
        out: *mut hvx_evict_outcome) -> c_int;
    pub fn hvx_extraction_current(publisher: *const hvx_extraction_publisher, key: *const hvx_planet_page_key,
        out: *mut hvx_published_surface) -> c_int;
    pub fn hvx_extraction_pending(publisher: *const hvx_extraction_publisher, key: *const hvx_planet_page_key,
        out: *mut hvx_reservation) -> c_int;
    pub fn hvx_extraction_publisher_get_counters(publisher: *const hvx_extraction_publisher,
        out: *mut hvx_extraction_publisher_counters) -> c_int;
    pub fn hvx_extraction_publisher_attach(publisher: *mut hvx_extraction_publisher, ctx: *mut hvx_ctx) -> c_int;
    pub fn hvx_extraction_commit(publisher: *mut hvx_extraction_publisher, chunk: *const u32, page_slot: *const u32,
        reservations: *const hvx_reservation, n: u32) -> c_int;
    pub fn hvx_extraction_publisher_buffer(publisher: *mut hvx_extraction_publisher, buffer_id: c_int) -> *mut c_void;
    pub fn hvx_extraction_publisher_read(publisher: *mut hvx_extraction_publisher, buffer_id: c_int, byte_offset: u64,
        bytes: u64, dst: *mut c_void) -> c_int;
}

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
