// extraction_publisher.h -- internal: host state of the bounded extraction publisher
// (PV/src/extraction.rs:342-664), shared by extraction_publisher.cpp (bookkeeping) and hvx_api.cu
// (device arenas + commit).
#pragma once

#include <cstdint>
#include <map>
#include <vector>

#include "../../include/hvx.h"

namespace hvx {

// RangeAllocator (PV/src/extraction.rs:605-664): first fit over a sorted, coalesced free list.
class ArenaAllocator {
public:
    explicit ArenaAllocator(uint32_t capacity);
    bool reserve(uint32_t count, hvx_arena_slice* out);
    void release(hvx_arena_slice released);
    uint32_t used() const;

private:
    uint32_t capacity_;
    std::vector<hvx_arena_slice> free_;
};

struct PageKeyLess {
    static int compare(const hvx_planet_page_key& a, const hvx_planet_page_key& b);
    bool operator()(const hvx_planet_page_key& a, const hvx_planet_page_key& b) const { return compare(a, b) < 0; }
};

inline hvx_extraction_range gpu_range(const hvx_surface_allocation& a, uint64_t generation) {
    hvx_extraction_range r;
    r.first_vertex = a.vertices.first;
    r.vertex_count = a.vertices.count;
    r.first_index = a.indices.first;
    r.index_count = a.indices.count;
    r.first_meshlet = a.meshlets.first;
    r.meshlet_count = a.meshlets.count;
    r.generation_low = static_cast<uint32_t>(generation);
    r.generation_high = static_cast<uint32_t>(generation >> 32);
    return r;
}

class ExtractionPublisher {
public:
    struct PageState {
        hvx_published_surface current;
        hvx_reservation pending;
        bool has_current = false, has_pending = false;
    };

    explicit ExtractionPublisher(const hvx_extraction_limits& limits);
    const hvx_extraction_limits& limits() const { return limits_; }
    const PageState* find(const hvx_planet_page_key& key) const;
    void counters(hvx_extraction_publisher_counters* out) const;
    int reserve(const hvx_planet_page_key& key, uint64_t generation, const hvx_surface_counts& counts,
                hvx_reservation_outcome* out);
    int publish(const hvx_reservation& reservation, hvx_publication_outcome* out);
    int cancel_pending(const hvx_planet_page_key& key, uint64_t generation, int* cancelled);
    void evict(const hvx_planet_page_key& key, uint64_t generation, hvx_evict_outcome* out);

private:
    size_t pending_pages() const;
    void release(const hvx_surface_allocation& a);
    void refresh_high_water();

    hvx_extraction_limits limits_;
    ArenaAllocator vertices_, indices_, meshlets_;
    std::map<hvx_planet_page_key, PageState, PageKeyLess> pages_;
    hvx_extraction_publisher_counters counters_;
};

}  // namespace hvx

// The opaque C handle: host bookkeeping + (after hvx_extraction_publisher_attach) device arenas.
struct hvx_extraction_publisher {
    explicit hvx_extraction_publisher(const hvx_extraction_limits& limits) : host(limits) {}
    hvx::ExtractionPublisher host;
    hvx_ctx* ctx = nullptr;
    void* buf[HVX_XPUB_COUNT] = {nullptr, nullptr, nullptr, nullptr};
    uint64_t bytes[HVX_XPUB_COUNT] = {0, 0, 0, 0};
    void* d_jobs = nullptr;  // grow-only staging of commit jobs
    uint64_t d_jobs_bytes = 0;
    void (*release_device)(hvx_extraction_publisher*) = nullptr;
};
