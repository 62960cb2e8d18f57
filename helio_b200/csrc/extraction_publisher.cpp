// extraction_publisher.cpp -- host side of the bounded, generation-safe publication contract.
//
// Follows PV/src/extraction.rs:
//   GpuExtractionRequest::new                     :24-42
//   ExtractionLimits::new / allocation_plan       :121-178,  validate_device :180-203
//   SurfaceCounts::validate                       :230-243
//   BoundedExtractionPublisher                    :342-603   (reserve, publish, cancel_pending, evict)
//   RangeAllocator (first fit, coalescing free)   :605-664
// Pure host code, no CUDA: the device step (copy into the reserved ranges) lives in hvx_api.cu
// (hvx_extraction_commit) and reaches this state through extraction_publisher.h.
#include "extraction_publisher.h"

#include <algorithm>
#include <cstring>
#include <new>

namespace hvx {

ArenaAllocator::ArenaAllocator(uint32_t capacity) : capacity_(capacity) { free_.push_back({0u, capacity}); }

// first fit: the lowest free range that is large enough, carved from its front
bool ArenaAllocator::reserve(uint32_t count, hvx_arena_slice* out) {
    *out = {0u, 0u};
    if (count == 0) return true;
    for (size_t i = 0; i < free_.size(); ++i) {
        if (free_[i].count < count) continue;
        *out = {free_[i].first, count};
        free_[i].first += count;
        free_[i].count -= count;
        if (free_[i].count == 0) free_.erase(free_.begin() + static_cast<std::ptrdiff_t>(i));
        return true;
    }
    return false;
}

// sorted insert, then merge with the neighbours it touches
void ArenaAllocator::release(hvx_arena_slice released) {
    if (released.count == 0) return;
    auto at = std::lower_bound(free_.begin(), free_.end(), released.first,
                               [](const hvx_arena_slice& r, uint32_t first) { return r.first < first; });
    at = free_.insert(at, released);
    size_t i = static_cast<size_t>(at - free_.begin());
    if (i > 0) --i;
    while (i + 1 < free_.size()) {
        const uint32_t end = free_[i].first + free_[i].count;
        if (end < free_[i + 1].first) {
            ++i;
            continue;
        }
        free_[i].count += free_[i + 1].count;
        free_.erase(free_.begin() + static_cast<std::ptrdiff_t>(i + 1));
    }
}

uint32_t ArenaAllocator::used() const {
    uint32_t free_total = 0;
    for (const auto& r : free_) free_total += r.count;
    return capacity_ - free_total;
}

namespace {

bool same_slice(const hvx_arena_slice& a, const hvx_arena_slice& b) { return a.first == b.first && a.count == b.count; }
bool same_allocation(const hvx_surface_allocation& a, const hvx_surface_allocation& b) {
    return same_slice(a.vertices, b.vertices) && same_slice(a.indices, b.indices) && same_slice(a.meshlets, b.meshlets);
}
bool same_reservation(const hvx_reservation& a, const hvx_reservation& b) {
    return PageKeyLess::compare(a.key, b.key) == 0 && a.generation == b.generation &&
           same_allocation(a.allocation, b.allocation);
}

bool checked_bytes(uint32_t count, uint64_t size, uint64_t* out) { return !__builtin_mul_overflow(static_cast<uint64_t>(count), size, out); }

int limits_plan(const hvx_extraction_limits& l, hvx_extraction_plan* plan) {
    if (l.max_page_slots == 0 || l.max_pending_pages == 0 || l.max_pending_pages > l.max_page_slots ||
        l.max_vertices == 0 || l.max_indices == 0 || l.max_meshlets == 0)
        return HVX_E_INVALID_LIMITS;
    hvx_extraction_plan p;
    if (!checked_bytes(l.max_pending_pages, sizeof(hvx_extraction_request), &p.request_bytes) ||
        !checked_bytes(l.max_page_slots, sizeof(hvx_extraction_range), &p.page_range_bytes) ||
        !checked_bytes(l.max_vertices, sizeof(hvx_vertex), &p.vertex_bytes) ||
        !checked_bytes(l.max_indices, sizeof(uint32_t), &p.index_bytes) ||
        !checked_bytes(l.max_meshlets, sizeof(hvx_meshlet), &p.meshlet_bytes))
        return HVX_E_ARITHMETIC_OVERFLOW;
    p.counter_bytes = sizeof(hvx_extraction_counters);
    uint64_t total = 0;
    for (uint64_t bytes : {p.request_bytes, p.page_range_bytes, p.vertex_bytes, p.index_bytes, p.meshlet_bytes, p.counter_bytes})
        if (__builtin_add_overflow(total, bytes, &total)) return HVX_E_ARITHMETIC_OVERFLOW;
    p.total_bytes = total;
    if (plan) *plan = p;
    return HVX_OK;
}

int validate_counts(const hvx_surface_counts& c) {
    if (c.indices % 3u != 0) return HVX_E_NON_TRIANGLE_INDEX_COUNT;
    const bool empty = c.vertices == 0 && c.indices == 0 && c.meshlets == 0;
    if ((c.vertices == 0 || c.indices == 0) && !empty) return HVX_E_INCOMPLETE_SURFACE_COUNTS;
    if (c.meshlets == 0 && c.indices != 0) return HVX_E_INCOMPLETE_SURFACE_COUNTS;
    return HVX_OK;
}

uint64_t saturating_inc(uint64_t v) { return v == UINT64_MAX ? v : v + 1; }

}  // namespace

int PageKeyLess::compare(const hvx_planet_page_key& a, const hvx_planet_page_key& b) {
    // derive(Ord) on PlanetPageKey { planet: PlanetId([u8; 16]), page: PageKey { lod, page_xyz } }
    const int planet = std::memcmp(a.planet_id, b.planet_id, 16);
    if (planet != 0) return planet < 0 ? -1 : 1;
    if (a.lod != b.lod) return a.lod < b.lod ? -1 : 1;
    for (int axis = 0; axis < 3; ++axis)
        if (a.page_xyz[axis] != b.page_xyz[axis]) return a.page_xyz[axis] < b.page_xyz[axis] ? -1 : 1;
    return 0;
}

ExtractionPublisher::ExtractionPublisher(const hvx_extraction_limits& limits)
    : limits_(limits), vertices_(limits.max_vertices), indices_(limits.max_indices), meshlets_(limits.max_meshlets) {
    std::memset(&counters_, 0, sizeof(counters_));
}

const ExtractionPublisher::PageState* ExtractionPublisher::find(const hvx_planet_page_key& key) const {
    auto it = pages_.find(key);
    return it == pages_.end() ? nullptr : &it->second;
}

size_t ExtractionPublisher::pending_pages() const {
    size_t n = 0;
    for (const auto& kv : pages_) n += kv.second.has_pending ? 1 : 0;
    return n;
}

void ExtractionPublisher::counters(hvx_extraction_publisher_counters* out) const {
    *out = counters_;
    uint64_t current = 0;
    for (const auto& kv : pages_) current += kv.second.has_current ? 1 : 0;
    out->current_pages = current;
    out->pending_pages = pending_pages();
    out->used_vertices = vertices_.used();
    out->used_indices = indices_.used();
    out->used_meshlets = meshlets_.used();
}

void ExtractionPublisher::release(const hvx_surface_allocation& a) {
    vertices_.release(a.vertices);
    indices_.release(a.indices);
    meshlets_.release(a.meshlets);
}

void ExtractionPublisher::refresh_high_water() {
    hvx_extraction_publisher_counters now;
    counters(&now);
    counters_.pending_high_water = std::max(counters_.pending_high_water, now.pending_pages);
    counters_.vertex_high_water = std::max(counters_.vertex_high_water, now.used_vertices);
    counters_.index_high_water = std::max(counters_.index_high_water, now.used_indices);
    counters_.meshlet_high_water = std::max(counters_.meshlet_high_water, now.used_meshlets);
}

int ExtractionPublisher::reserve(const hvx_planet_page_key& key, uint64_t generation, const hvx_surface_counts& counts,
                                 hvx_reservation_outcome* out) {
    std::memset(out, 0, sizeof(*out));
    if (int rc = validate_counts(counts)) return rc;
    if (const PageState* st = find(key); st && st->has_current) {
        if (generation < st->current.generation) {
            counters_.stale_rejected = saturating_inc(counters_.stale_rejected);
            out->kind = HVX_RESERVE_STALE;
            out->newest_generation = st->current.generation;
            return HVX_OK;
        }
        if (generation == st->current.generation) {
            out->kind = HVX_RESERVE_CURRENT;
            out->current = st->current;
            return HVX_OK;
        }
    }
    if (const PageState* st = find(key); st && st->has_pending) {
        const hvx_reservation pending = st->pending;
        if (generation < pending.generation) {
            counters_.stale_rejected = saturating_inc(counters_.stale_rejected);
            out->kind = HVX_RESERVE_STALE;
            out->newest_generation = pending.generation;
            return HVX_OK;
        }
        if (generation == pending.generation) {
            const hvx_surface_allocation& a = pending.allocation;
            if (a.vertices.count != counts.vertices || a.indices.count != counts.indices || a.meshlets.count != counts.meshlets)
                return HVX_E_GENERATION_CONFLICT;
            out->kind = HVX_RESERVE_DUPLICATE_PENDING;
            out->reservation = pending;
            return HVX_OK;
        }
        int cancelled = 0;
        if (int rc = cancel_pending(key, pending.generation, &cancelled)) return rc;  // a newer request supersedes it
    }
    if (pending_pages() >= limits_.max_pending_pages) {
        counters_.backpressured = saturating_inc(counters_.backpressured);
        out->detail = limits_.max_pending_pages;
        return HVX_E_PENDING_CAPACITY;
    }
    hvx_surface_allocation a;
    auto capacity_error = [&](uint32_t which) {
        counters_.backpressured = saturating_inc(counters_.backpressured);
        out->detail = which;
        return HVX_E_ARENA_CAPACITY;
    };
    if (!vertices_.reserve(counts.vertices, &a.vertices)) return capacity_error(0);
    if (!indices_.reserve(counts.indices, &a.indices)) {
        vertices_.release(a.vertices);
        return capacity_error(1);
    }
    if (!meshlets_.reserve(counts.meshlets, &a.meshlets)) {
        indices_.release(a.indices);
        vertices_.release(a.vertices);
        return capacity_error(2);
    }
    hvx_reservation r;
    std::memset(&r, 0, sizeof(r));
    r.key = key;
    r.generation = generation;
    r.allocation = a;
    PageState& st = pages_[key];
    st.pending = r;
    st.has_pending = true;
    counters_.reservations = saturating_inc(counters_.reservations);
    refresh_high_water();
    out->kind = HVX_RESERVED;
    out->reservation = r;
    return HVX_OK;
}

int ExtractionPublisher::publish(const hvx_reservation& reservation, hvx_publication_outcome* out) {
    std::memset(out, 0, sizeof(*out));
    auto it = pages_.find(reservation.key);
    if (it == pages_.end()) return HVX_E_RESERVATION_MISSING;
    PageState& st = it->second;
    bool any = false;
    uint64_t newest = 0;
    if (st.has_current) {
        newest = st.current.generation;
        any = true;
    }
    if (st.has_pending && (!any || st.pending.generation > newest)) {
        newest = st.pending.generation;
        any = true;
    }
    if (any && reservation.generation < newest) {
        counters_.stale_rejected = saturating_inc(counters_.stale_rejected);
        out->kind = 1;
        out->newest_generation = newest;
        return HVX_OK;
    }
    if (!st.has_pending || !same_reservation(st.pending, reservation)) return HVX_E_RESERVATION_MISMATCH;
    st.has_pending = false;
    hvx_published_surface current;
    current.generation = reservation.generation;
    current.allocation = reservation.allocation;
    if (st.has_current) {
        out->has_replaced = 1;
        out->replaced = st.current;
        release(st.current.allocation);  // only now are the old ranges recycled
        counters_.replacements = saturating_inc(counters_.replacements);
    }
    st.current = current;
    st.has_current = true;
    counters_.publications = saturating_inc(counters_.publications);
    out->kind = 0;
    out->current = current;
    return HVX_OK;
}

int ExtractionPublisher::cancel_pending(const hvx_planet_page_key& key, uint64_t generation, int* cancelled) {
    *cancelled = 0;
    auto it = pages_.find(key);
    if (it == pages_.end() || !it->second.has_pending) return HVX_OK;
    PageState& st = it->second;
    if (st.pending.generation != generation) return HVX_E_RESERVATION_MISMATCH;
    st.has_pending = false;
    release(st.pending.allocation);
    counters_.cancellations = saturating_inc(counters_.cancellations);
    if (!st.has_current) pages_.erase(it);
    *cancelled = 1;
    return HVX_OK;
}

void ExtractionPublisher::evict(const hvx_planet_page_key& key, uint64_t generation, hvx_evict_outcome* out) {
    std::memset(out, 0, sizeof(*out));
    auto it = pages_.find(key);
    if (it == pages_.end()) {
        out->kind = 1;
        return;
    }
    const PageState st = it->second;
    uint64_t newest = 0;
    if (st.has_current) newest = st.current.generation;
    if (st.has_pending) newest = std::max(newest, st.pending.generation);
    if (generation < newest) {
        counters_.stale_rejected = saturating_inc(counters_.stale_rejected);
        out->kind = 2;
        out->newest_generation = newest;
        return;
    }
    pages_.erase(it);
    if (st.has_current) release(st.current.allocation);
    if (st.has_pending) release(st.pending.allocation);
    counters_.evictions = saturating_inc(counters_.evictions);
    out->kind = 0;
}

}  // namespace hvx

// ---- C ABI (host-only entry points; the device ones are in hvx_api.cu) ---------------------------------
extern "C" {

int hvx_extraction_request_new(uint32_t page_slot, uint64_t generation, uint32_t transition_mask, uint64_t dirty_microbricks,
                               hvx_extraction_request* out) {
    if (!out) return HVX_E_INVALID_ARGUMENT;
    if (transition_mask & ~0x3fu) return HVX_E_TRANSITION_MASK;
    std::memset(out, 0, sizeof(*out));
    out->page_slot = page_slot;
    out->generation_low = static_cast<uint32_t>(generation);
    out->generation_high = static_cast<uint32_t>(generation >> 32);
    out->transition_mask = transition_mask;
    out->dirty_microbricks_low = static_cast<uint32_t>(dirty_microbricks);
    out->dirty_microbricks_high = static_cast<uint32_t>(dirty_microbricks >> 32);
    return HVX_OK;
}

int hvx_extraction_limits_plan(const hvx_extraction_limits* limits, hvx_extraction_plan* plan_out) {
    if (!limits) return HVX_E_INVALID_ARGUMENT;
    return hvx::limits_plan(*limits, plan_out);
}

int hvx_extraction_limits_validate_device(const hvx_extraction_limits* limits, uint64_t max_buffer_size,
                                          uint64_t max_storage_buffer_binding_size, const char** name_out,
                                          uint64_t* requested_out) {
    if (!limits) return HVX_E_INVALID_ARGUMENT;
    hvx_extraction_plan plan;
    if (int rc = hvx::limits_plan(*limits, &plan)) return rc;
    const struct {
        const char* name;
        uint64_t bytes;
    } buffers[] = {{"extraction requests", plan.request_bytes}, {"page extraction ranges", plan.page_range_bytes},
                   {"terrain vertices", plan.vertex_bytes},     {"terrain indices", plan.index_bytes},
                   {"terrain meshlets", plan.meshlet_bytes},    {"extraction counters", plan.counter_bytes}};
    for (const auto& b : buffers) {
        if (b.bytes > max_buffer_size || b.bytes > max_storage_buffer_binding_size) {
            if (name_out) *name_out = b.name;
            if (requested_out) *requested_out = b.bytes;
            return HVX_E_DEVICE_BUFFER_LIMIT;
        }
    }
    return HVX_OK;
}

void hvx_extraction_gpu_range(const hvx_surface_allocation* a, uint64_t generation, hvx_extraction_range* out) {
    if (!a || !out) return;
    *out = hvx::gpu_range(*a, generation);
}

int hvx_extraction_publisher_create(const hvx_extraction_limits* limits, hvx_extraction_publisher** out) {
    if (!limits || !out) return HVX_E_INVALID_ARGUMENT;
    *out = nullptr;
    if (int rc = hvx::limits_plan(*limits, nullptr)) return rc;
    auto* pub = new (std::nothrow) hvx_extraction_publisher(*limits);
    if (!pub) return HVX_E_INVALID_ARGUMENT;
    *out = pub;
    return HVX_OK;
}

void hvx_extraction_publisher_destroy(hvx_extraction_publisher* pub) {
    if (!pub) return;
    if (pub->release_device) pub->release_device(pub);
    delete pub;
}

int hvx_extraction_reserve(hvx_extraction_publisher* pub, const hvx_planet_page_key* key, uint64_t generation,
                           const hvx_surface_counts* counts, hvx_reservation_outcome* out) {
    if (!pub || !key || !counts || !out) return HVX_E_INVALID_ARGUMENT;
    return pub->host.reserve(*key, generation, *counts, out);
}

int hvx_extraction_publish(hvx_extraction_publisher* pub, const hvx_reservation* reservation, hvx_publication_outcome* out) {
    if (!pub || !reservation || !out) return HVX_E_INVALID_ARGUMENT;
    return pub->host.publish(*reservation, out);
}

int hvx_extraction_cancel_pending(hvx_extraction_publisher* pub, const hvx_planet_page_key* key, uint64_t generation,
                                  int* cancelled_out) {
    if (!pub || !key) return HVX_E_INVALID_ARGUMENT;
    int cancelled = 0;
    const int rc = pub->host.cancel_pending(*key, generation, &cancelled);
    if (cancelled_out) *cancelled_out = cancelled;
    return rc;
}

int hvx_extraction_evict(hvx_extraction_publisher* pub, const hvx_planet_page_key* key, uint64_t generation,
                         hvx_evict_outcome* out) {
    if (!pub || !key || !out) return HVX_E_INVALID_ARGUMENT;
    pub->host.evict(*key, generation, out);
    return HVX_OK;
}

int hvx_extraction_current(const hvx_extraction_publisher* pub, const hvx_planet_page_key* key, hvx_published_surface* out) {
    if (!pub || !key) return HVX_E_INVALID_ARGUMENT;
    const auto* st = pub->host.find(*key);
    if (!st || !st->has_current) return 0;
    if (out) *out = st->current;
    return 1;
}

int hvx_extraction_pending(const hvx_extraction_publisher* pub, const hvx_planet_page_key* key, hvx_reservation* out) {
    if (!pub || !key) return HVX_E_INVALID_ARGUMENT;
    const auto* st = pub->host.find(*key);
    if (!st || !st->has_pending) return 0;
    if (out) *out = st->pending;
    return 1;
}

int hvx_extraction_publisher_get_counters(const hvx_extraction_publisher* pub, hvx_extraction_publisher_counters* out) {
    if (!pub || !out) return HVX_E_INVALID_ARGUMENT;
    pub->host.counters(out);
    return HVX_OK;
}

}  // extern "C"
