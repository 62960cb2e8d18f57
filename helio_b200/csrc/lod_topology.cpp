// lod_topology.cpp -- host-side LOD scheduler input: which faces of which chunks own a
// transition mesh, the bounded horizon page plan, and the static multi-GPU partition.
//
// Follows PV/src/lod_topology.rs:
//   TerrainLodTopology::new                      :27-100
//   validate_tangent_root                        :123-146
//   HorizonLodFixturePlan::build_with_minimum_lod :169-217
//   PageBounds / tangent_children / balance      :232-346
// The page edge is a parameter (the reference hard-codes 32 cells).  Pure host code, no CUDA.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <set>
#include <tuple>
#include <vector>

#include "../../include/hvx.h"

namespace {

typedef __int128 i128;
typedef unsigned __int128 u128;

constexpr uint32_t MAX_ADDRESSABLE_LOD = 57;  // helio-planet-voxel-core/src/types.rs:18

struct Key {
    uint8_t lod;
    int64_t x, y, z;
    bool operator<(const Key& o) const { return std::tie(lod, x, y, z) < std::tie(o.lod, o.x, o.y, o.z); }
    bool operator==(const Key& o) const { return lod == o.lod && x == o.x && y == o.y && z == o.z; }
    int64_t axis(int a) const { return a == 0 ? x : a == 1 ? y : z; }
};

struct Bounds {
    i128 min[3], max[3];
};

// PageKey::lod0_cell_min with checked arithmetic (types.rs:258-280)
bool page_min(const Key& k, uint32_t edge, int64_t out[3]) {
    if (k.lod > MAX_ADDRESSABLE_LOD) return false;
    int64_t span;
    if (__builtin_mul_overflow(static_cast<int64_t>(edge), static_cast<int64_t>(1) << k.lod, &span)) return false;
    for (int a = 0; a < 3; ++a)
        if (__builtin_mul_overflow(k.axis(a), span, &out[a])) return false;
    return true;
}

bool bounds_of(const Key& k, uint32_t edge, Bounds* b) {
    int64_t mn[3];
    if (!page_min(k, edge, mn)) return false;
    const i128 span = static_cast<i128>(edge) << k.lod;
    for (int a = 0; a < 3; ++a) {
        b->min[a] = mn[a];
        b->max[a] = static_cast<i128>(mn[a]) + span;
    }
    return true;
}

bool overlaps_volume(const Bounds& l, const Bounds& r) {
    for (int a = 0; a < 3; ++a)
        if (!(l.min[a] < r.max[a] && r.min[a] < l.max[a])) return false;
    return true;
}

// returns axis and whether `l` touches with its positive face, or -1
int shared_face(const Bounds& l, const Bounds& r, bool* l_positive) {
    for (int a = 0; a < 3; ++a) {
        const bool pos = l.max[a] == r.min[a], neg = r.max[a] == l.min[a];
        if (!pos && !neg) continue;
        bool overlap = true;
        for (int o = 0; o < 3; ++o)
            if (o != a && !(l.min[o] < r.max[o] && r.min[o] < l.max[o])) overlap = false;
        if (overlap) {
            *l_positive = pos;
            return a;
        }
    }
    return -1;
}

constexpr int TANGENT_AXES[2] = {0, 2};

bool tangent_shared_edge(const Bounds& l, const Bounds& r) {
    for (int i = 0; i < 2; ++i) {
        const int a = TANGENT_AXES[i], o = TANGENT_AXES[1 - i];
        if ((l.max[a] == r.min[a] || r.max[a] == l.min[a]) && l.min[o] < r.max[o] && r.min[o] < l.max[o]) return true;
    }
    return false;
}

int64_t div_euclid(int64_t a, int64_t b) {  // b > 0
    int64_t q = a / b;
    if (a % b < 0) --q;
    return q;
}

// PageKey::address_lod0_cell (types.rs:300-309), page part only
bool address_lod0_cell(uint32_t lod, const int64_t cell[3], uint32_t edge, Key* out) {
    if (lod > MAX_ADDRESSABLE_LOD) return false;
    const int64_t span = static_cast<int64_t>(edge) << lod;
    out->lod = static_cast<uint8_t>(lod);
    out->x = div_euclid(cell[0], span);
    out->y = div_euclid(cell[1], span);
    out->z = div_euclid(cell[2], span);
    return true;
}

int topology(const std::set<Key>& unique, uint32_t edge, std::map<Key, uint8_t>* masks, hvx_lod_stats* stats) {
    if (unique.empty()) return HVX_E_TOPOLOGY_EMPTY;
    std::vector<Key> pages(unique.begin(), unique.end());
    std::vector<Bounds> bounds(pages.size());
    for (size_t i = 0; i < pages.size(); ++i)
        if (!bounds_of(pages[i], edge, &bounds[i])) return HVX_E_ADDRESS;
    masks->clear();
    for (const Key& k : pages) (*masks)[k] = 0;
    for (size_t l = 0; l < pages.size(); ++l)
        for (size_t r = l + 1; r < pages.size(); ++r) {
            if (overlaps_volume(bounds[l], bounds[r])) return HVX_E_TOPOLOGY_OVERLAP;
            bool l_pos = false;
            const int axis = shared_face(bounds[l], bounds[r], &l_pos);
            if (axis < 0) continue;
            const int diff = pages[l].lod > pages[r].lod ? pages[l].lod - pages[r].lod : pages[r].lod - pages[l].lod;
            if (diff > 1) return HVX_E_TOPOLOGY_UNBALANCED;
            if (diff == 1) {
                const bool l_coarse = pages[l].lod > pages[r].lod;
                const Key& coarse = l_coarse ? pages[l] : pages[r];
                const bool positive = l_coarse ? l_pos : !l_pos;
                (*masks)[coarse] |= static_cast<uint8_t>(1u << (2 * axis + (positive ? 1 : 0)));
            }
        }
    if (stats) {
        stats->pages = static_cast<uint32_t>(pages.size());
        stats->minimum_lod = pages.front().lod;
        stats->maximum_lod = pages.back().lod;
        stats->transition_faces = 0;
        for (auto& kv : *masks) stats->transition_faces += static_cast<uint32_t>(__builtin_popcount(kv.second));
    }
    return HVX_OK;
}

bool tangent_children(const Key& parent, Key out[4]) {
    if (parent.lod == 0) return false;
    int64_t bx, bz;
    if (__builtin_mul_overflow(parent.x, static_cast<int64_t>(2), &bx) ||
        __builtin_mul_overflow(parent.z, static_cast<int64_t>(2), &bz))
        return false;
    const uint8_t lod = parent.lod - 1;
    out[0] = Key{lod, bx, -1, bz};
    out[1] = Key{lod, bx + 1, -1, bz};
    out[2] = Key{lod, bx, -1, bz + 1};
    out[3] = Key{lod, bx + 1, -1, bz + 1};
    return true;
}

int balance(std::set<Key>* leaves, uint32_t edge, size_t max_pages) {
    for (;;) {
        std::vector<Key> pages(leaves->begin(), leaves->end());
        std::vector<Bounds> bounds(pages.size());
        for (size_t i = 0; i < pages.size(); ++i)
            if (!bounds_of(pages[i], edge, &bounds[i])) return HVX_E_ADDRESS;
        bool found = false;
        Key coarse{};
        for (size_t l = 0; l < pages.size() && !found; ++l)
            for (size_t r = l + 1; r < pages.size(); ++r) {
                const int diff = pages[l].lod > pages[r].lod ? pages[l].lod - pages[r].lod : pages[r].lod - pages[l].lod;
                if (diff > 1 && tangent_shared_edge(bounds[l], bounds[r])) {
                    coarse = pages[l].lod > pages[r].lod ? pages[l] : pages[r];
                    found = true;
                    break;
                }
            }
        if (!found) return HVX_OK;
        if (leaves->size() + 3 > max_pages) return HVX_E_TOPOLOGY_PAGE_BUDGET;
        Key kids[4];
        if (!tangent_children(coarse, kids)) return HVX_E_ADDRESS;
        leaves->erase(coarse);
        leaves->insert(kids, kids + 4);
    }
}

int validate_tangent_root(const std::set<Key>& pages, const Key& root, uint32_t edge) {
    Bounds rb;
    if (!bounds_of(root, edge, &rb)) return HVX_E_ADDRESS;
    const u128 root_area = static_cast<u128>(rb.max[0] - rb.min[0]) * static_cast<u128>(rb.max[2] - rb.min[2]);
    u128 covered = 0;
    for (const Key& k : pages) {
        Bounds b;
        if (!bounds_of(k, edge, &b)) return HVX_E_ADDRESS;
        for (int i = 0; i < 2; ++i) {
            const int a = TANGENT_AXES[i];
            if (!(rb.min[a] <= b.min[a] && b.max[a] <= rb.max[a])) return HVX_E_TOPOLOGY_COVERAGE;
        }
        covered += static_cast<u128>(b.max[0] - b.min[0]) * static_cast<u128>(b.max[2] - b.min[2]);
    }
    return covered == root_area ? HVX_OK : HVX_E_TOPOLOGY_COVERAGE;
}

void write_pages(const std::map<Key, uint8_t>& masks, hvx_page* out) {
    size_t i = 0;
    for (auto& kv : masks) {
        std::memset(&out[i], 0, sizeof(hvx_page));
        out[i].page_xyz[0] = kv.first.x;
        out[i].page_xyz[1] = kv.first.y;
        out[i].page_xyz[2] = kv.first.z;
        out[i].lod = kv.first.lod;
        out[i].transition_mask = kv.second;
        ++i;
    }
}

}  // namespace

extern "C" {

int hvx_lod_topology(hvx_page* pages, uint32_t n, uint32_t edge, hvx_lod_stats* stats) {
    if ((n && !pages) || (edge != 32 && edge != 64)) return HVX_E_INVALID_ARGUMENT;
    std::set<Key> unique;
    for (uint32_t i = 0; i < n; ++i) {
        Key k{pages[i].lod, pages[i].page_xyz[0], pages[i].page_xyz[1], pages[i].page_xyz[2]};
        int64_t mn[3];
        if (!page_min(k, edge, mn)) return HVX_E_ADDRESS;  // page.validate()
        if (!unique.insert(k).second) return HVX_E_TOPOLOGY_DUPLICATE;
    }
    std::map<Key, uint8_t> masks;
    const int rc = topology(unique, edge, &masks, stats);
    if (rc) return rc;
    write_pages(masks, pages);
    return HVX_OK;
}

int hvx_horizon_plan(const int64_t focus[3], uint32_t root_lod, uint32_t minimum_lod, uint32_t max_pages,
                     uint32_t edge, hvx_page* out, uint32_t* n_out, hvx_page* root_out, hvx_lod_stats* stats) {
    if (!focus || !out || !n_out || (edge != 32 && edge != 64)) return HVX_E_INVALID_ARGUMENT;
    if (root_lod == 0 || root_lod > MAX_ADDRESSABLE_LOD) return HVX_E_TOPOLOGY_ROOT_LOD;
    if (minimum_lod >= root_lod) return HVX_E_TOPOLOGY_MINIMUM_LOD;
    if (max_pages < 4) return HVX_E_TOPOLOGY_PAGE_BUDGET;
    Key root;
    if (!address_lod0_cell(root_lod, focus, edge, &root)) return HVX_E_ADDRESS;
    root.y = -1;
    std::set<Key> leaves{root};
    for (int target_lod = static_cast<int>(root_lod) - 1; target_lod >= static_cast<int>(minimum_lod); --target_lod) {
        Key target;
        if (!address_lod0_cell(static_cast<uint32_t>(target_lod) + 1, focus, edge, &target)) return HVX_E_ADDRESS;
        target.y = -1;
        if (leaves.erase(target) == 0) return HVX_E_TOPOLOGY_MISSING_PARENT;
        Key kids[4];
        if (!tangent_children(target, kids)) return HVX_E_ADDRESS;
        leaves.insert(kids, kids + 4);
        const int rc = balance(&leaves, edge, max_pages);
        if (rc) return rc;
        if (leaves.size() > max_pages) return HVX_E_TOPOLOGY_PAGE_BUDGET;
    }
    std::map<Key, uint8_t> masks;
    int rc = topology(leaves, edge, &masks, stats);
    if (rc) return rc;
    if ((rc = validate_tangent_root(leaves, root, edge))) return rc;
    write_pages(masks, out);
    *n_out = static_cast<uint32_t>(masks.size());
    if (root_out) {
        std::memset(root_out, 0, sizeof(hvx_page));
        root_out->page_xyz[0] = root.x;
        root_out->page_xyz[1] = root.y;
        root_out->page_xyz[2] = root.z;
        root_out->lod = root.lod;
    }
    return HVX_OK;
}

uint64_t hvx_chunk_cost(uint32_t edge, uint32_t transition_mask) {
    const uint64_t s = edge + 2, w = 2ull * edge + 3;
    return 4 * s * s * s + 12 * w * w * static_cast<uint64_t>(__builtin_popcount(transition_mask & 0x3fu));
}

int hvx_partition_chunks(const uint64_t* cost, uint32_t n, uint32_t ranks, uint32_t* owner) {
    if (ranks == 0 || (n && (!cost || !owner))) return HVX_E_INVALID_ARGUMENT;
    // longest-processing-time first: heaviest chunk to the currently lightest rank
    std::vector<uint32_t> order(n);
    for (uint32_t i = 0; i < n; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return cost[a] > cost[b]; });
    std::vector<uint64_t> load(ranks, 0);
    for (uint32_t i : order) {
        uint32_t best = 0;
        for (uint32_t r = 1; r < ranks; ++r)
            if (load[r] < load[best]) best = r;
        owner[i] = best;
        load[best] += cost[i];
    }
    return HVX_OK;
}

}  // extern "C"
