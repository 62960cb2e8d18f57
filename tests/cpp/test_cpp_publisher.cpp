// CPU: the C++ mirror of BoundedExtractionPublisher replays two of the reference's unit tests
// (PV/src/extraction.rs:775-800 replacement_keeps_current_ranges_until_atomic_publication,
//  :834-853 capacity_failure_rolls_back_partial_ranges_and_reuses_freed_space).  No GPU involved.
#include <cstdio>
#include <cstring>

#include "helio_voxel_cuda.hpp"

using namespace helio_voxel_cuda;

#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) {                                                     \
            std::printf("FAILED %s:%d %s\n", __FILE__, __LINE__, #cond);   \
            return 1;                                                      \
        }                                                                  \
    } while (0)

static hvx_planet_page_key key(int64_t index) {
    hvx_planet_page_key k{};
    std::memset(k.planet_id, 1, sizeof(k.planet_id));
    k.page_xyz[0] = index;
    return k;
}

static hvx_surface_counts counts(uint32_t vertices, uint32_t triangles, uint32_t meshlets) { return {vertices, triangles * 3, meshlets}; }

int main() {
    {
        BoundedExtractionPublisher publisher(hvx_extraction_limits{2, 1, 20, 60, 4});
        const hvx_reservation first = publisher.reserve(key(0), 1, counts(8, 8, 1)).reservation;
        publisher.publish(first);
        hvx_published_surface before{};
        CHECK(publisher.current(key(0), &before) && before.generation == 1);
        const hvx_reservation_outcome replacement = publisher.reserve(key(0), 2, counts(10, 10, 1));
        CHECK(replacement.kind == HVX_RESERVED);
        hvx_published_surface still{};
        CHECK(publisher.current(key(0), &still) && still.generation == 1 && still.allocation.vertices.first == before.allocation.vertices.first);
        CHECK(publisher.counters().used_vertices == 18);
        const hvx_publication_outcome outcome = publisher.publish(replacement.reservation);
        CHECK(outcome.kind == 0 && outcome.has_replaced == 1 && outcome.replaced.generation == 1);
        CHECK(publisher.counters().used_vertices == 10);
    }
    {
        BoundedExtractionPublisher publisher(hvx_extraction_limits{2, 2, 10, 12, 1});
        bool threw = false;
        try {
            publisher.reserve(key(0), 1, counts(8, 5, 1));
        } catch (const Error& e) {
            threw = e.status() == HVX_E_ARENA_CAPACITY && publisher.detail() == 1;  // ExtractionCapacity::Indices
        }
        CHECK(threw && publisher.counters().used_vertices == 0);
        publisher.publish(publisher.reserve(key(0), 1, counts(8, 4, 1)).reservation);
        CHECK(publisher.evict(key(0), 1).kind == 0);
        CHECK(publisher.counters().used_vertices == 0 && publisher.counters().used_indices == 0);
        CHECK(publisher.reserve(key(1), 1, counts(10, 4, 1)).reservation.allocation.vertices.first == 0);
        bool invalid = false;
        try {
            BoundedExtractionPublisher bad(hvx_extraction_limits{2, 3, 1, 1, 1});
        } catch (const Error& e) {
            invalid = e.status() == HVX_E_INVALID_LIMITS;
        }
        CHECK(invalid);
    }
    std::printf("OK cpp publisher\n");
    return 0;
}
