// helio_voxel_cuda.hpp -- header-only C++ mirror of the reference's extractor objects, over the
// C ABI in hvx.h.  Same names, argument order and error behaviour as
//   TransvoxelGpuExtractorConfig / TransvoxelGpuExtractor      PV/src/transvoxel_emit.rs:57-396
//   TransvoxelGpuClassifier                                    PV/src/transvoxel_gpu.rs:148-356
//   TransvoxelGpuTransitionExtractor(+Config)                  PV/src/transvoxel_transition_gpu.rs:148-520
// (PV = crates/passes/3d/helio-pass-planetary-voxel in the reference tree).  Where the Rust takes
// (&wgpu::Device, &wgpu::Queue) these take a CUDA device ordinal at construction; Result<_, enum>
// becomes an exception carrying the hvx_status.  No CUDA headers are needed to use this file.
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "hvx.h"

namespace helio_voxel_cuda {

constexpr uint32_t PAGE_EDGE = 32;  // helio-planet-voxel-core/src/types.rs:6

/// TransvoxelGpuError / TransvoxelTransitionGpuError (PV/src/transvoxel_gpu.rs:445-459,
/// PV/src/transvoxel_transition_gpu.rs:712-732) plus the ABI's own failures.
class Error : public std::runtime_error {
   public:
    Error(int status, const std::string& what) : std::runtime_error(what), status_(status) {}
    int status() const { return status_; }
    bool is_sample_count() const { return status_ == HVX_E_SAMPLE_COUNT; }
    bool is_invalid_extraction_capacity() const { return status_ == HVX_E_INVALID_CAPACITY; }
    bool is_device_limit() const { return status_ == HVX_E_DEVICE_LIMIT; }
    bool is_transition_mask() const { return status_ == HVX_E_TRANSITION_MASK; }

   private:
    int status_;
};

/// PV/src/transvoxel_emit.rs:57-85
struct TransvoxelGpuExtractorConfig {
    uint32_t max_vertices = 393216;
    uint32_t max_indices = 491520;
    static TransvoxelGpuExtractorConfig create(uint32_t max_vertices, uint32_t max_indices) {
        if (max_vertices == 0 || max_indices == 0)
            throw Error(HVX_E_INVALID_CAPACITY, "Transvoxel extraction capacities must be nonzero");
        return {max_vertices, max_indices};
    }
};

/// PV/src/transvoxel_transition_gpu.rs:148-181
struct TransvoxelGpuTransitionExtractorConfig {
    uint32_t max_vertices = 73728;
    uint32_t max_indices = 221184;
    static TransvoxelGpuTransitionExtractorConfig create(uint32_t max_vertices, uint32_t max_indices) {
        if (max_vertices == 0 || max_indices == 0)
            throw Error(HVX_E_INVALID_CAPACITY, "Transvoxel transition capacities must be nonzero");
        return {max_vertices, max_indices};
    }
};

struct ResourceStats {
    uint32_t buffers;
    uint64_t allocated_bytes;
};

namespace detail {
class Ctx {
   public:
    Ctx(int device, const hvx_config& cfg) {
        const int rc = hvx_create(&ctx_, device, &cfg);
        if (rc != HVX_OK) throw Error(rc, hvx_last_error(nullptr));
    }
    ~Ctx() { hvx_destroy(ctx_); }
    Ctx(const Ctx&) = delete;
    Ctx& operator=(const Ctx&) = delete;
    hvx_ctx* get() const { return ctx_; }
    void check(int rc) const {
        if (rc != HVX_OK) throw Error(rc, hvx_last_error(ctx_));
    }
    template <typename T>
    std::vector<T> read(int buffer, uint64_t first, uint64_t count) const {
        std::vector<T> out(count);
        check(hvx_read(ctx_, buffer, first * sizeof(T), count * sizeof(T), out.data()));
        return out;
    }
    ResourceStats stats() const {
        uint32_t n = 0;
        for (int b = 0; b < HVX_BUF_COUNT; ++b) n += hvx_buffer_bytes(ctx_, b) != 0;
        return {n, hvx_allocated_bytes(ctx_)};
    }

   private:
    hvx_ctx* ctx_ = nullptr;
};
inline hvx_chunk_desc desc(uint64_t generation, uint64_t dirty, uint32_t mask) { return {generation, dirty, mask, 0u, 0u, 0u}; }
}  // namespace detail

/// One page per dispatch, like the reference (PV/src/transvoxel_emit.rs:92-396).
class TransvoxelGpuExtractor {
   public:
    explicit TransvoxelGpuExtractor(int device, TransvoxelGpuExtractorConfig config = {}, uint32_t edge = PAGE_EDGE,
                                    bool debug_records = true)
        : config_(TransvoxelGpuExtractorConfig::create(config.max_vertices, config.max_indices)),
          ctx_(device, hvx_config{edge, 1, config.max_vertices, config.max_indices, 0, 0,
                                  debug_records ? HVX_CFG_DEBUG_RECORDS : 0u, 0}),
          cells_(uint64_t(edge) * edge * edge) {}

    /// dispatch(samples, generation, dirty_microbricks, transition_mask)  -- transvoxel_emit.rs:233-254
    void dispatch(const uint32_t* samples, size_t sample_count, uint64_t generation, uint64_t dirty_microbricks,
                  uint8_t transition_mask) {
        const hvx_chunk_desc d = detail::desc(generation, dirty_microbricks, transition_mask);
        ctx_.check(hvx_extract_regular(ctx_.get(), samples, sample_count, &d, 1));
    }
    hvx_emission_counters counters_buffer() const { return ctx_.read<hvx_emission_counters>(HVX_BUF_REGULAR_COUNTERS, 0, 1)[0]; }
    std::vector<hvx_vertex> vertices_buffer(uint64_t count) const { return ctx_.read<hvx_vertex>(HVX_BUF_REGULAR_VERTICES, 0, count); }
    std::vector<uint32_t> indices_buffer(uint64_t count) const { return ctx_.read<uint32_t>(HVX_BUF_REGULAR_INDICES, 0, count); }
    std::vector<hvx_cell_offset> offsets_buffer() const { return ctx_.read<hvx_cell_offset>(HVX_BUF_REGULAR_OFFSETS, 0, cells_); }
    std::vector<hvx_scan_block> blocks_buffer() const { return ctx_.read<hvx_scan_block>(HVX_BUF_REGULAR_BLOCKS, 0, cells_ / 256); }
    /// device pointers for zero-copy consumers (the arenas are overwritten by the next dispatch)
    const hvx_vertex* device_vertices() const { return static_cast<const hvx_vertex*>(hvx_buffer(ctx_.get(), HVX_BUF_REGULAR_VERTICES)); }
    const uint32_t* device_indices() const { return static_cast<const uint32_t*>(hvx_buffer(ctx_.get(), HVX_BUF_REGULAR_INDICES)); }
    TransvoxelGpuExtractorConfig config() const { return config_; }
    ResourceStats resource_stats() const { return ctx_.stats(); }
    void resize(uint32_t, uint32_t) {}  // extraction owns no surface-size-dependent resources

   private:
    TransvoxelGpuExtractorConfig config_;
    detail::Ctx ctx_;
    uint64_t cells_;
};

/// PV/src/transvoxel_gpu.rs:148-356
class TransvoxelGpuClassifier {
   public:
    explicit TransvoxelGpuClassifier(int device, uint32_t edge = PAGE_EDGE)
        : ctx_(device, hvx_config{edge, 1, 393216, 491520, 0, 0, HVX_CFG_DEBUG_RECORDS, 0}), cells_(uint64_t(edge) * edge * edge) {}
    void dispatch(const uint32_t* samples, size_t sample_count, uint64_t generation, uint64_t dirty_microbricks) {
        const hvx_chunk_desc d = detail::desc(generation, dirty_microbricks, 0);
        ctx_.check(hvx_classify_regular(ctx_.get(), samples, sample_count, &d, 1));
    }
    std::vector<hvx_cell_record> output_buffer() const { return ctx_.read<hvx_cell_record>(HVX_BUF_REGULAR_CELLS, 0, cells_); }
    hvx_classify_counters counters_buffer() const { return ctx_.read<hvx_classify_counters>(HVX_BUF_REGULAR_CLASSIFY, 0, 1)[0]; }
    ResourceStats resource_stats() const { return ctx_.stats(); }
    void resize(uint32_t, uint32_t) {}

   private:
    detail::Ctx ctx_;
    uint64_t cells_;
};

/// PV/src/transvoxel_transition_gpu.rs:190-520
class TransvoxelGpuTransitionExtractor {
   public:
    explicit TransvoxelGpuTransitionExtractor(int device, TransvoxelGpuTransitionExtractorConfig config = {},
                                              uint32_t edge = PAGE_EDGE, bool debug_records = true)
        : config_(TransvoxelGpuTransitionExtractorConfig::create(config.max_vertices, config.max_indices)),
          ctx_(device, hvx_config{edge, 1, 1, 1, config.max_vertices, config.max_indices,
                                  debug_records ? HVX_CFG_DEBUG_RECORDS : 0u, 0}),
          cells_(6ull * edge * edge) {}
    /// dispatch(face_slabs, transition_mask, generation)  -- transvoxel_transition_gpu.rs:366-380
    void dispatch(const uint32_t* face_slabs, size_t sample_count, uint8_t transition_mask, uint64_t generation) {
        const hvx_chunk_desc d = detail::desc(generation, ~0ull, transition_mask);
        ctx_.check(hvx_extract_transition(ctx_.get(), face_slabs, sample_count, &d, 1));
    }
    hvx_transition_counters counters_buffer() const { return ctx_.read<hvx_transition_counters>(HVX_BUF_TRANSITION_COUNTERS, 0, 1)[0]; }
    std::vector<hvx_vertex> vertices_buffer(uint64_t count) const { return ctx_.read<hvx_vertex>(HVX_BUF_TRANSITION_VERTICES, 0, count); }
    std::vector<uint32_t> indices_buffer(uint64_t count) const { return ctx_.read<uint32_t>(HVX_BUF_TRANSITION_INDICES, 0, count); }
    std::vector<hvx_cell_record> cells_buffer() const { return ctx_.read<hvx_cell_record>(HVX_BUF_TRANSITION_CELLS, 0, cells_); }
    TransvoxelGpuTransitionExtractorConfig config() const { return config_; }
    ResourceStats resource_stats() const { return ctx_.stats(); }
    void resize(uint32_t, uint32_t) {}

   private:
    TransvoxelGpuTransitionExtractorConfig config_;
    detail::Ctx ctx_;
    uint64_t cells_;
};

/// The batch form the page queue feeds once it stops submitting one page per frame: N chunks per dispatch into
/// fixed-stride per-chunk slots (chunk k at k * max_vertices / k * max_indices).  `prepare` validates and builds the
/// descriptors (no device work), `encode` queues the extraction, `extract_to_host` adds the pipelined packed read-back
/// -- the prepare / encode split of PV/src/transvoxel_emit.rs:256-330.
class ChunkBatchExtractor {
   public:
    struct Request {
        uint64_t generation = 1;
        uint64_t dirty_microbricks = ~0ull;
        uint8_t transition_mask = 0;
        uint32_t cost_hint = 0;   // e.g. the chunk's vertex count last time: heaviest chunks start first
        bool uniform = false;     // the producer knows the chunk holds no surface: not uploaded, not read
    };
    struct Meshes {
        std::vector<hvx_vertex> vertices;  // back to back in chunk order
        std::vector<uint32_t> indices;     // chunk-local values
        std::vector<hvx_range> ranges;     // packed placement per chunk
        std::vector<hvx_emission_counters> counters;
    };

    ChunkBatchExtractor(int device, uint32_t edge, uint32_t max_chunks, TransvoxelGpuExtractorConfig regular = {49152, 73728})
        : edge_(edge), ctx_(device, hvx_config{edge, max_chunks, regular.max_vertices, regular.max_indices, 0, 0, 0u, 0}) {}

    std::vector<hvx_chunk_desc> prepare(const std::vector<Request>& requests) const {
        std::vector<hvx_chunk_desc> descs(requests.size());
        for (size_t i = 0; i < requests.size(); ++i) {
            if (requests[i].transition_mask & ~0x3fu) throw Error(HVX_E_TRANSITION_MASK, "transition mask uses bits outside the six page faces");
            descs[i] = hvx_chunk_desc{requests[i].generation, requests[i].dirty_microbricks, requests[i].transition_mask,
                                      requests[i].cost_hint, requests[i].uniform ? HVX_CHUNK_UNIFORM : 0u, 0u};
        }
        return descs;
    }
    uint64_t sample_words() const { return uint64_t(edge_ + 2) * (edge_ + 2) * (edge_ + 2); }
    /// ExtractionFixture::new on the device for `n` pages (xyz triples); the samples stay in the ctx arena.
    void fill(uint32_t kind, const int64_t* page_xyz, const uint8_t* lod, uint32_t n) { ctx_.check(hvx_fill_density(ctx_.get(), kind, page_xyz, lod, n, nullptr)); }
    /// samples: host or device CellWords, or nullptr for the ctx arena.
    void encode(const std::vector<hvx_chunk_desc>& descs, const uint32_t* samples = nullptr) {
        const uint32_t n = static_cast<uint32_t>(descs.size());
        ctx_.check(hvx_extract_regular(ctx_.get(), samples, n * sample_words(), descs.data(), n));
    }
    Meshes extract_to_host(const std::vector<hvx_chunk_desc>& descs, const uint32_t* samples, uint64_t vertex_capacity, uint64_t index_capacity) {
        const uint32_t n = static_cast<uint32_t>(descs.size());
        Meshes m{std::vector<hvx_vertex>(vertex_capacity), std::vector<uint32_t>(index_capacity), std::vector<hvx_range>(n),
                 std::vector<hvx_emission_counters>(n)};
        uint64_t tv = 0, ti = 0;
        ctx_.check(hvx_extract_regular_to_host(ctx_.get(), samples, n * sample_words(), descs.data(), n, m.vertices.data(), vertex_capacity,
                                               m.indices.data(), index_capacity, m.ranges.data(), m.counters.data(), &tv, &ti));
        m.vertices.resize(tv);
        m.indices.resize(ti);
        return m;
    }
    /// One sphere edit (GpuVoxelEdit) on the resident samples of the `n` pages; returns the dirty microbricks per chunk.
    std::vector<uint64_t> apply_edit(const hvx_voxel_edit& edit, const int64_t* page_xyz, const uint8_t* lod, uint32_t n, uint32_t* touched = nullptr) {
        std::vector<uint64_t> dirty(n);
        ctx_.check(hvx_apply_edit(ctx_.get(), &edit, page_xyz, lod, n, nullptr, dirty.data(), touched));
        return dirty;
    }
    /// Optional vertex-reuse output (off by default: the reference shares no vertices between cells): merges the
    /// bit-identical vertex records of chunks [0, n) of the last extraction in place (hvx_weld_meshes).
    void weld_meshes(uint32_t n, bool transition = false) { ctx_.check(hvx_weld_meshes(ctx_.get(), transition ? 1 : 0, n)); }
    std::vector<hvx_emission_counters> counters_buffer(uint32_t n) const { return ctx_.read<hvx_emission_counters>(HVX_BUF_REGULAR_COUNTERS, 0, n); }
    std::vector<hvx_range> ranges_buffer(uint32_t n) const { return ctx_.read<hvx_range>(HVX_BUF_REGULAR_RANGES, 0, n); }
    void synchronize() { ctx_.check(hvx_synchronize(ctx_.get())); }
    ResourceStats resource_stats() const { return ctx_.stats(); }

   private:
    uint32_t edge_;
    detail::Ctx ctx_;
};

/// Batched GpuSurfaceSampler (PV/src/surface_sampling.rs:184-350): halo blocks + transition slabs of n jobs
/// from the resident page atlas into the context's sample arenas, then the extractors run on them.
class GpuSurfaceSampler {
   public:
    GpuSurfaceSampler(int device, uint32_t max_jobs, TransvoxelGpuExtractorConfig regular = {},
                      TransvoxelGpuTransitionExtractorConfig transition = {})
        : ctx_(device, hvx_config{PAGE_EDGE, max_jobs, regular.max_vertices, regular.max_indices, transition.max_vertices,
                                  transition.max_indices, 0u, 0}) {}
    /// Upload the residency layer's page table once per publication (PV/src/table.rs:62-72); dispatch(table = nullptr)
    /// then reads the bound table.  entries = 0 unbinds.
    void bind_page_table(const hvx_page_table_entry* table, uint32_t entries) { ctx_.check(hvx_gather_bind_table(ctx_.get(), table, entries)); }
    /// prepare + encode for n jobs.  table: table_mask + 1 entries, or nullptr for the bound table; atlas: linear R32Uint
    /// texels (host or device).
    void dispatch(const hvx_residency& residency, const hvx_page_table_entry* table, const uint32_t* atlas,
                  uint64_t atlas_words, const hvx_gather_job* jobs, uint32_t n) {
        ctx_.check(hvx_gather_surface(ctx_.get(), &residency, table, atlas, atlas_words, jobs, n));
        jobs_.assign(jobs, jobs + n);
    }
    /// regular + transition extraction of the gathered jobs, straight from the arenas
    void extract(uint64_t dirty_microbricks = ~0ull) {
        std::vector<hvx_chunk_desc> d(jobs_.size());
        bool any_faces = false;
        for (size_t i = 0; i < jobs_.size(); ++i) {
            d[i] = detail::desc(jobs_[i].generation_low | (static_cast<uint64_t>(jobs_[i].generation_high) << 32),
                                dirty_microbricks, static_cast<uint8_t>(jobs_[i].transition_mask));
            any_faces |= jobs_[i].transition_mask != 0;
        }
        const uint32_t n = static_cast<uint32_t>(d.size());
        ctx_.check(hvx_extract_regular(ctx_.get(), nullptr, n * 39304ull, d.data(), n));
        if (any_faces) ctx_.check(hvx_extract_transition(ctx_.get(), nullptr, n * 80802ull, d.data(), n));
    }
    std::vector<hvx_gather_counters> counters_buffer() const { return ctx_.read<hvx_gather_counters>(HVX_BUF_GATHER_COUNTERS, 0, jobs_.size()); }
    std::vector<uint32_t> indirect_buffer() const { return ctx_.read<uint32_t>(HVX_BUF_GATHER_INDIRECT, 0, 24 * jobs_.size()); }
    std::vector<hvx_emission_counters> regular_counters() const { return ctx_.read<hvx_emission_counters>(HVX_BUF_REGULAR_COUNTERS, 0, jobs_.size()); }
    ResourceStats resource_stats() const { return ctx_.stats(); }

   private:
    detail::Ctx ctx_;
    std::vector<hvx_gather_job> jobs_;
};

/// BoundedExtractionPublisher (PV/src/extraction.rs:342-603) over hvx_extraction_*: host-side, generation-safe
/// arena placement.  Errors (ExtractionError, :672-698) surface as Error with the HVX_E_* status; `detail()` of
/// a failed reserve is the exhausted arena (0 vertices, 1 indices, 2 meshlets) or the pending-page maximum.
class BoundedExtractionPublisher {
   public:
    explicit BoundedExtractionPublisher(const hvx_extraction_limits& limits) : limits_(limits) {
        const int status = hvx_extraction_publisher_create(&limits, &pub_);
        if (status != HVX_OK) throw Error(status, hvx_status_name(status));
    }
    ~BoundedExtractionPublisher() { hvx_extraction_publisher_destroy(pub_); }
    BoundedExtractionPublisher(const BoundedExtractionPublisher&) = delete;
    BoundedExtractionPublisher& operator=(const BoundedExtractionPublisher&) = delete;

    const hvx_extraction_limits& limits() const { return limits_; }
    hvx_reservation_outcome reserve(const hvx_planet_page_key& key, uint64_t generation, const hvx_surface_counts& counts) {
        hvx_reservation_outcome out{};
        const int status = hvx_extraction_reserve(pub_, &key, generation, &counts, &out);
        detail_ = out.detail;
        if (status != HVX_OK) throw Error(status, hvx_status_name(status));
        return out;
    }
    hvx_publication_outcome publish(const hvx_reservation& reservation) {
        hvx_publication_outcome out{};
        const int status = hvx_extraction_publish(pub_, &reservation, &out);
        if (status != HVX_OK) throw Error(status, hvx_status_name(status));
        return out;
    }
    bool cancel_pending(const hvx_planet_page_key& key, uint64_t generation) {
        int cancelled = 0;
        const int status = hvx_extraction_cancel_pending(pub_, &key, generation, &cancelled);
        if (status != HVX_OK) throw Error(status, hvx_status_name(status));
        return cancelled != 0;
    }
    hvx_evict_outcome evict(const hvx_planet_page_key& key, uint64_t generation) {
        hvx_evict_outcome out{};
        hvx_extraction_evict(pub_, &key, generation, &out);
        return out;
    }
    bool current(const hvx_planet_page_key& key, hvx_published_surface* out) const { return hvx_extraction_current(pub_, &key, out) == 1; }
    bool pending(const hvx_planet_page_key& key, hvx_reservation* out) const { return hvx_extraction_pending(pub_, &key, out) == 1; }
    hvx_extraction_publisher_counters counters() const {
        hvx_extraction_publisher_counters out{};
        hvx_extraction_publisher_get_counters(pub_, &out);
        return out;
    }
    uint32_t detail() const { return detail_; }
    /// device side: bounded arenas on the context's GPU + one copy launch per batch of reservations
    hvx_extraction_publisher* handle() { return pub_; }

   private:
    hvx_extraction_limits limits_;
    hvx_extraction_publisher* pub_ = nullptr;
    uint32_t detail_ = 0;
};

}  // namespace helio_voxel_cuda
