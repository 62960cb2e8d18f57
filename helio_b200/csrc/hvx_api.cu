// hvx_api.cu -- the extern "C" boundary (include/hvx.h): context, arenas, validation, launches.
//
// Mirrors the reference's extractor objects:
//   TransvoxelGpuExtractor            PV/src/transvoxel_emit.rs:92-396
//   TransvoxelGpuClassifier           PV/src/transvoxel_gpu.rs:148-356
//   TransvoxelGpuTransitionExtractor  PV/src/transvoxel_transition_gpu.rs:190-520
// Ownership follows the reference: the ctx owns every device buffer, the caller borrows inputs
// for the duration of the call, outputs are overwritten by the next dispatch.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "extraction_publisher.h"
#include "hvx_kernels.h"

using namespace hvx;

struct hvx_ctx {
    hvx_config cfg{};
    int device = 0;
    DeviceInfo dev{};
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t handover = nullptr;   // orders a new stream after the old one (hvx_set_stream)
    cudaStream_t copy_stream = nullptr;  // host-sample dispatches: uploads of sub-batch k+1 beside the kernel of sub-batch k
    cudaStream_t d2h_stream = nullptr;   // hvx_extract_regular_to_host: pack + read-back behind the kernels
    std::vector<cudaEvent_t> events;     // [2 * sub-batches]: upload done, kernel done
    hvx_range* h_ranges = nullptr;       // pinned, [max_chunks]: ranges on their way to the host
    uint32_t* d_uniform = nullptr;       // [max_chunks] chunks flagged HVX_CHUNK_UNIFORM
    SplitItem* d_items = nullptr;        // [MAX_SPLIT_ITEMS] z-ranges of chunks, when a dispatch is too small to fill the machine
    uint4* d_item_totals = nullptr;      // [MAX_SPLIT_ITEMS] per-part totals (the look-back state of the split walk), zeroed once
    uint32_t split_generation = 0;       // tag of the newest split dispatch
    void* buf[HVX_BUF_COUNT] = {};
    uint64_t buf_bytes[HVX_BUF_COUNT] = {};
    // The per-dispatch host inputs of the regular path (descriptors, start order, skipped chunks, split items) travel
    // in ONE copy: they are packed back to back in a pinned staging buffer (two of them, alternating, each guarded by
    // an event) and land in one device block; the four pointers below point into it.
    unsigned char* d_batch = nullptr;
    uint64_t batch_bytes = 0;
    unsigned char* h_batch[2] = {nullptr, nullptr};
    cudaEvent_t batch_done[2] = {nullptr, nullptr};
    uint32_t batch_flip = 0;
    unsigned char* h_tbatch[2] = {nullptr, nullptr};  // the same for the transition dispatch's descriptors
    cudaEvent_t tbatch_done[2] = {nullptr, nullptr};
    uint32_t tbatch_flip = 0;
    uint32_t* d_touched = nullptr;    // [max_chunks] hvx_apply_edit: chunks whose samples the edit changes
    uint32_t* h_touched = nullptr;    // pinned twin (the list goes up with a truly asynchronous copy)
    cudaEvent_t touched_done = nullptr;
    std::vector<int64_t> edit_pages;  // page list / LODs hvx_apply_edit uploaded last (an edit frame repeats them)
    std::vector<uint8_t> edit_lods;
    ChunkDesc* d_descs = nullptr;     // [max_chunks] descriptors of the last REGULAR dispatch
    ChunkDesc* d_tdescs = nullptr;    // [max_chunks] descriptors of the last TRANSITION dispatch
    uint32_t n_regular = 0, n_transition = 0;  // sizes of those dispatches (hvx_build_meshlets reads the generations)
    uint32_t debug_mode = 0;          // hvx_debug_set_mode
    uint32_t spread_pct = 75;         // heavy chunks of a hinted batch are spread over this share of the start order (0: plain descending order; hvx_debug_set_mode bits 12..19)
    bool split_last_wave = false;     // hvx_debug_set_mode bit 9: split the chunks of a thin last wave (measured slower; A/B and tests)
    bool no_split = false;            // hvx_debug_set_mode bit 8: never split chunks across CTAs (A/B measurements, tests)
    uint32_t* d_order = nullptr;      // [max_chunks] start order of a batch with cost hints
    int64_t* d_pages = nullptr;
    uint8_t* d_lod = nullptr;
    // fBm terrain fill: distinct (x, z, lod) columns of the batch and their shared height maps
    uint32_t* d_col_index = nullptr;  // [max_chunks] chunk -> column
    long long* d_col_xz = nullptr;    // [max_chunks][2]
    uint8_t* d_col_lod = nullptr;     // [max_chunks]
    float* d_heights = nullptr;       // [heights_cols][(E+2)^2], grown on demand
    uint64_t heights_cols = 0;
    void* stage[3] = {nullptr, nullptr, nullptr};  // gather inputs staged from the host: table, atlas, jobs
    uint64_t stage_bytes[3] = {0, 0, 0};
    void* bound_table = nullptr;      // hvx_gather_bind_table: the page table, resident until the next bind
    uint64_t bound_table_bytes = 0;
    uint32_t* d_work = nullptr;       // [8] self-rearming work counters: regular [0],[2], transition [1],[3], weld [4],[6]
    uint32_t* d_weld = nullptr;       // hvx_weld_meshes scratch (grow-only)
    uint64_t weld_bytes = 0;
    hvx_range* d_packed = nullptr;    // [max_chunks] packed placement for hvx_read_meshes
    void* pack_v = nullptr;           // staging for hvx_read_meshes
    void* pack_i = nullptr;
    uint64_t pack_v_bytes = 0, pack_i_bytes = 0;
    uint64_t allocated = 0;
    uint64_t launches = 0;
    std::string error;
};

namespace {

constexpr uint32_t MAX_SPLIT_ITEMS = 8192;  // a dispatch is only split while it has fewer than 16 waves of chunks (16 x 444 resident CTAs)
constexpr uint32_t MAX_PARTS = 16;  // the look-back reads all earlier parts at once (one lane each), so its cost does not grow with the parts

thread_local std::string g_create_error;

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != device) ok = cudaSetDevice(device) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int fail(hvx_ctx* ctx, int status, const char* fmt, ...) {
    char msg[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(msg, sizeof msg, fmt, ap);
    va_end(ap);
    if (ctx) ctx->error = msg; else g_create_error = msg;
    return status;
}

int cuda_fail(hvx_ctx* ctx, cudaError_t e, const char* what) {
    return fail(ctx, HVX_E_CUDA, "CUDA error in %s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}

#define HVX_CUDA(ctx, call)                                   \
    do {                                                      \
        cudaError_t e_ = (call);                              \
        if (e_ != cudaSuccess) return cuda_fail(ctx, e_, #call); \
    } while (0)

uint64_t sample_words(uint32_t edge) { return static_cast<uint64_t>(edge + 2) * (edge + 2) * (edge + 2); }
uint64_t slab_words(uint32_t edge) { return 18ull * (2 * edge + 3) * (2 * edge + 3); }
uint64_t cells(uint32_t edge) { return static_cast<uint64_t>(edge) * edge * edge; }
uint64_t tcells(uint32_t edge) { return 6ull * edge * edge; }

uint64_t arena_bytes(const hvx_config& c, int id) {
    const uint64_t n = c.max_chunks;
    const bool dbg = (c.flags & HVX_CFG_DEBUG_RECORDS) != 0;
    const bool tr = c.max_transition_vertices != 0;
    switch (id) {
        case HVX_BUF_SAMPLES: return n * sample_words(c.edge) * 4;
        case HVX_BUF_SLABS: return n * slab_words(c.edge) * 4;
        case HVX_BUF_REGULAR_VERTICES: return n * c.max_vertices * sizeof(hvx_vertex);
        case HVX_BUF_REGULAR_INDICES: return n * c.max_indices * 4ull;
        case HVX_BUF_REGULAR_COUNTERS: return n * sizeof(hvx_emission_counters);
        case HVX_BUF_REGULAR_CLASSIFY: return n * sizeof(hvx_classify_counters);
        case HVX_BUF_REGULAR_RANGES: return n * sizeof(hvx_range);
        case HVX_BUF_REGULAR_CELLS: return dbg ? n * cells(c.edge) * sizeof(hvx_cell_record) : 0;
        case HVX_BUF_REGULAR_OFFSETS: return dbg ? n * cells(c.edge) * sizeof(hvx_cell_offset) : 0;
        case HVX_BUF_REGULAR_BLOCKS: return dbg ? n * (cells(c.edge) / 256) * sizeof(hvx_scan_block) : 0;
        case HVX_BUF_TRANSITION_VERTICES: return tr ? n * c.max_transition_vertices * sizeof(hvx_vertex) : 0;
        case HVX_BUF_TRANSITION_INDICES: return tr ? n * c.max_transition_indices * 4ull : 0;
        case HVX_BUF_TRANSITION_COUNTERS: return tr ? n * sizeof(hvx_transition_counters) : 0;
        case HVX_BUF_TRANSITION_RANGES: return tr ? n * sizeof(hvx_range) : 0;
        case HVX_BUF_TRANSITION_CELLS: return tr && dbg ? n * tcells(c.edge) * sizeof(hvx_cell_record) : 0;
        case HVX_BUF_TRANSITION_OFFSETS: return tr && dbg ? n * tcells(c.edge) * sizeof(hvx_cell_offset) : 0;
        case HVX_BUF_TRANSITION_BLOCKS: return tr && dbg ? n * (tcells(c.edge) / 256) * sizeof(hvx_scan_block) : 0;
        case HVX_BUF_REGULAR_MESHLETS: return n * ((c.max_indices + 62ull) / 63) * sizeof(hvx_meshlet);
        case HVX_BUF_REGULAR_MESHLET_BOUNDS: return n * ((c.max_indices + 62ull) / 63) * sizeof(hvx_meshlet_bounds);
        case HVX_BUF_REGULAR_MESHLET_COUNTS: return n * 4;
        case HVX_BUF_TRANSITION_MESHLETS: return tr ? n * ((c.max_transition_indices + 62ull) / 63) * sizeof(hvx_meshlet) : 0;
        case HVX_BUF_TRANSITION_MESHLET_BOUNDS: return tr ? n * ((c.max_transition_indices + 62ull) / 63) * sizeof(hvx_meshlet_bounds) : 0;
        case HVX_BUF_TRANSITION_MESHLET_COUNTS: return tr ? n * 4 : 0;
        case HVX_BUF_GATHER_COUNTERS: return c.edge == 32 ? n * sizeof(hvx_gather_counters) : 0;
        case HVX_BUF_GATHER_INDIRECT: return c.edge == 32 ? n * 24 * 4ull : 0;
        default: return 0;
    }
}

int ensure_buffer(hvx_ctx* ctx, int id) {
    if (id < 0 || id >= HVX_BUF_COUNT) return fail(ctx, HVX_E_INVALID_ARGUMENT, "unknown buffer id %d", id);
    if (ctx->buf[id]) return HVX_OK;
    const uint64_t bytes = arena_bytes(ctx->cfg, id);
    if (bytes == 0) return fail(ctx, HVX_E_INVALID_ARGUMENT, "buffer %d is not enabled by this configuration", id);
    void* ptr = nullptr;
    cudaError_t e = cudaMalloc(&ptr, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, HVX_E_DEVICE_LIMIT, "DeviceLimit: buffer %d requires %llu bytes: %s", id,
                    static_cast<unsigned long long>(bytes), cudaGetErrorString(e));
    }
    // debug records start "stale" (generation 0, valid bit clear) like a freshly created wgpu buffer
    if (id >= HVX_BUF_REGULAR_COUNTERS && id != HVX_BUF_TRANSITION_VERTICES && id != HVX_BUF_TRANSITION_INDICES) {
        e = cudaMemsetAsync(ptr, 0, bytes, ctx->stream);
        if (e != cudaSuccess) {
            cudaFree(ptr);
            return cuda_fail(ctx, e, "cudaMemsetAsync");
        }
    }
    ctx->buf[id] = ptr;
    ctx->buf_bytes[id] = bytes;
    ctx->allocated += bytes;
    return HVX_OK;
}

template <typename T>
int small_alloc(hvx_ctx* ctx, T** out, uint64_t count) {
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(out), count * sizeof(T));
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc");
    ctx->allocated += count * sizeof(T);
    return HVX_OK;
}

// Resolve an input pointer: NULL -> ctx arena; device pointer -> as is; host pointer -> staged
// into the ctx arena with an async H2D copy on the ctx stream.
int resolve_input(hvx_ctx* ctx, const uint32_t* ptr, uint64_t words, int arena_id, const uint32_t** out) {
    if (ptr == nullptr) {
        int rc = ensure_buffer(ctx, arena_id);
        if (rc) return rc;
        *out = static_cast<const uint32_t*>(ctx->buf[arena_id]);
        return HVX_OK;
    }
    cudaPointerAttributes attr{};
    cudaError_t e = cudaPointerGetAttributes(&attr, ptr);
    if (e != cudaSuccess) {
        cudaGetLastError();
        attr.type = cudaMemoryTypeUnregistered;
    }
    if (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged) {
        if (reinterpret_cast<uintptr_t>(ptr) % 16 != 0)
            return fail(ctx, HVX_E_INVALID_ARGUMENT, "device sample pointer must be 16-byte aligned");
        if (attr.type == cudaMemoryTypeDevice && attr.device != ctx->device)
            return fail(ctx, HVX_E_INVALID_ARGUMENT, "sample pointer lives on device %d, ctx is on device %d",
                        attr.device, ctx->device);
        *out = ptr;
        return HVX_OK;
    }
    int rc = ensure_buffer(ctx, arena_id);
    if (rc) return rc;
    HVX_CUDA(ctx, cudaMemcpyAsync(ctx->buf[arena_id], ptr, words * 4, cudaMemcpyHostToDevice, ctx->stream));
    *out = static_cast<const uint32_t*>(ctx->buf[arena_id]);
    return HVX_OK;
}

int upload_descs(hvx_ctx* ctx, ChunkDesc* dst, const hvx_chunk_desc* descs, uint32_t n) {
    static_assert(sizeof(hvx_chunk_desc) == sizeof(ChunkDesc), "desc layout");
    // through pinned staging (two buffers, reused once the copy that read them is done): a copy from the caller's
    // pageable array is staged by the driver anyway, synchronously, and costs the stream a second operation
    const uint32_t flip = ctx->tbatch_flip ^= 1u;
    HVX_CUDA(ctx, cudaEventSynchronize(ctx->tbatch_done[flip]));
    memcpy(ctx->h_tbatch[flip], descs, static_cast<size_t>(n) * sizeof(ChunkDesc));
    HVX_CUDA(ctx, cudaMemcpyAsync(dst, ctx->h_tbatch[flip], static_cast<size_t>(n) * sizeof(ChunkDesc), cudaMemcpyHostToDevice, ctx->stream));
    HVX_CUDA(ctx, cudaEventRecord(ctx->tbatch_done[flip], ctx->stream));
    return HVX_OK;
}

int check_batch(hvx_ctx* ctx, const void* descs, uint32_t n) {
    if (!ctx) return HVX_E_INVALID_ARGUMENT;
    if (n > ctx->cfg.max_chunks)
        return fail(ctx, HVX_E_BATCH_CAPACITY, "batch of %u chunks exceeds max_chunks %u", n, ctx->cfg.max_chunks);
    if (n != 0 && descs == nullptr) return fail(ctx, HVX_E_INVALID_ARGUMENT, "descs is NULL");
    return HVX_OK;
}

bool is_device_pointer(hvx_ctx* ctx, const void* ptr, int* status) {
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, ptr) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    if (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeManaged) return false;
    if (reinterpret_cast<uintptr_t>(ptr) % 16 != 0)
        *status = fail(ctx, HVX_E_INVALID_ARGUMENT, "device sample pointer must be 16-byte aligned");
    else if (attr.type == cudaMemoryTypeDevice && attr.device != ctx->device)
        *status = fail(ctx, HVX_E_INVALID_ARGUMENT, "sample pointer lives on device %d, ctx is on device %d", attr.device, ctx->device);
    return true;
}

int ensure_pipeline(hvx_ctx* ctx, size_t events) {
    if (!ctx->copy_stream) HVX_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    if (!ctx->d2h_stream) HVX_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
    while (ctx->events.size() < events) {
        cudaEvent_t e = nullptr;
        HVX_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->events.push_back(e);
    }
    return HVX_OK;
}

// Packed host destination of a pipelined dispatch (hvx_extract_regular_to_host).
struct HostMeshOut {
    hvx_vertex* vertices;
    uint64_t vertex_cap;
    uint32_t* indices;
    uint64_t index_cap;
    hvx_range* ranges;
    hvx_emission_counters* counters;
    uint64_t total_vertices = 0, total_indices = 0;
};

int grow_device(hvx_ctx* ctx, void** ptr, uint64_t* have, uint64_t need, cudaStream_t user) {
    if (*have >= need) return HVX_OK;
    if (*ptr) {
        HVX_CUDA(ctx, cudaStreamSynchronize(user));
        cudaFree(*ptr);
        ctx->allocated -= *have;
        *ptr = nullptr;
        *have = 0;
    }
    const uint64_t bytes = need + need / 4 + 4096;
    cudaError_t e = cudaMalloc(ptr, bytes);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc(pack staging)");
    *have = bytes;
    ctx->allocated += bytes;
    return HVX_OK;
}

// The order in which the chunks of a hinted batch are started (hvx_start_order exposes it for tests).
// Descending hint, ties in chunk order: the scheduler's LPT rule, so the launch does not end on a lone heavy chunk.
// A batch of a few heavy chunks among many light ones (the headline: 298 of 4096 cross the surface, the others only
// stream) is bound by HBM as a whole, and a heavy chunk leaves its SM's share of the bandwidth unused while its
// emission runs.  Started all at once, the heavy chunks leave most of the machine's bandwidth idle for their duration;
// spread over the start order, the SMs that stream take up what the emitting ones leave.  So: heavy = more than four
// times the median hint (no chunk of an all-surface batch is), placed evenly over the first `pct` per cent of the
// start order, still heaviest first; the rest of the order is light chunks.  pct = 0: plain descending order.
template <class HintOf>
void start_order(uint32_t* order, uint32_t m, HintOf hint_of, uint32_t pct) {
    std::stable_sort(order, order + m, [&](uint32_t a, uint32_t b) { return hint_of(a) > hint_of(b); });
    if (m == 0 || pct == 0) return;
    const uint32_t median = hint_of(order[m / 2]);
    uint32_t heavy = 0;
    while (heavy < m && static_cast<uint64_t>(hint_of(order[heavy])) > 4ull * median) ++heavy;
    if (heavy == 0 || heavy > m / 4) return;
    const uint32_t span = std::max<uint32_t>(heavy, static_cast<uint32_t>(static_cast<uint64_t>(m) * std::min(pct, 100u) / 100));
    std::vector<uint32_t> mixed(m);
    uint32_t h = 0, l = heavy;
    for (uint32_t pos = 0; pos < m; ++pos) {
        const bool slot_of_heavy = h < heavy && pos == static_cast<uint32_t>(static_cast<uint64_t>(h) * span / heavy);
        mixed[pos] = slot_of_heavy ? order[h++] : (l < m ? order[l++] : order[h++]);
    }
    std::copy(mixed.begin(), mixed.end(), order);
}

// One regular dispatch over n chunks.
//   * device-resident samples (or the ctx arena): one launch.
//   * HOST samples: the batch is cut into sub-batches of ~256 MiB.  Sub-batch k is copied on the copy stream
//     while sub-batch k-1 is extracted on the ctx stream (one event per hand-off), so the call costs the upload
//     and one sub-batch of kernel time instead of upload + kernel.  Chunks flagged HVX_CHUNK_UNIFORM are neither
//     uploaded nor walked; one small launch writes their (empty) records.
//   * `out` (hvx_extract_regular_to_host): each sub-batch's ranges come back through pinned staging as soon as
//     its kernel is done; this thread places the sub-batch in the packed output and queues pack + D2H on a third
//     stream, so the read-back of sub-batch k overlaps the upload of k+2 and the kernel of k+1.
int run_regular(hvx_ctx* ctx, const uint32_t* samples, uint64_t words, const hvx_chunk_desc* descs, uint32_t n,
                uint32_t mode, HostMeshOut* out = nullptr) {
    int rc = check_batch(ctx, descs, n);
    if (rc) return rc;
    const uint64_t chunk_words = sample_words(ctx->cfg.edge);
    const uint64_t expected = static_cast<uint64_t>(n) * chunk_words;
    if (words != expected)
        return fail(ctx, HVX_E_SAMPLE_COUNT, "Transvoxel classification received %llu samples; expected %llu",
                    static_cast<unsigned long long>(words), static_cast<unsigned long long>(expected));
    if (n == 0) return HVX_OK;
    for (uint32_t i = 0; i < n; ++i)
        if (descs[i].flags & ~HVX_CHUNK_UNIFORM) return fail(ctx, HVX_E_INVALID_ARGUMENT, "chunk %u: unknown descriptor flags %#x", i, descs[i].flags);
    DeviceGuard guard(ctx->device);

    // ---- where the samples are -----------------------------------------------------------------------------
    const uint32_t* d_samples = nullptr;
    bool from_host = false;
    if (samples == nullptr) {
        if ((rc = ensure_buffer(ctx, HVX_BUF_SAMPLES))) return rc;
        d_samples = static_cast<const uint32_t*>(ctx->buf[HVX_BUF_SAMPLES]);
    } else {
        int status = HVX_OK;
        if (is_device_pointer(ctx, samples, &status)) {
            if (status) return status;
            d_samples = samples;
        } else {
            if ((rc = ensure_buffer(ctx, HVX_BUF_SAMPLES))) return rc;
            d_samples = static_cast<const uint32_t*>(ctx->buf[HVX_BUF_SAMPLES]);
            from_host = true;
        }
    }

    // ---- sub-batches, work lists (heaviest first, uniform chunks left out) -----------------------------------
    const uint32_t sub = from_host ? static_cast<uint32_t>(std::max<uint64_t>(1, (256ull << 20) / (chunk_words * 4))) : n;
    const uint32_t n_sub = (n + sub - 1) / sub;
    if ((from_host || out) && (rc = ensure_pipeline(ctx, 2ull * n_sub))) return rc;
    bool hinted = false, any_uniform = false;
    for (uint32_t i = 0; i < n; ++i) {
        hinted |= descs[i].cost_hint != 0u;
        any_uniform |= (descs[i].flags & HVX_CHUNK_UNIFORM) != 0u || descs[i].dirty_microbricks == 0;
    }
    std::vector<uint32_t> order, uniform, n_work(n_sub), skipped_begin(n_sub + 1, 0);  // uniform: ids relative to their sub-batch
    if (hinted || any_uniform) order.resize(n);
    for (uint32_t k = 0; k < n_sub; ++k) {
        const uint32_t first = k * sub, count = std::min(sub, n - first);
        if (order.empty()) {
            n_work[k] = count;
            continue;
        }
        uint32_t m = 0;
        for (uint32_t i = 0; i < count; ++i) {
            // nothing to walk: flagged uniform, or no dirty microbrick (an edit frame re-submits every resident chunk)
            if ((descs[first + i].flags & HVX_CHUNK_UNIFORM) || descs[first + i].dirty_microbricks == 0) uniform.push_back(i);
            else order[first + m++] = i;  // ids are relative to the sub-batch
        }
        n_work[k] = m;
        skipped_begin[k + 1] = static_cast<uint32_t>(uniform.size());
        // descending hint, ties in chunk order (the scheduler's LPT rule)
        if (hinted) start_order(order.data() + first, m, [&](uint32_t i) { return descs[first + i].cost_hint; }, ctx->spread_pct);
    }
    // ---- too few chunks to fill the machine (one page, an edit frame): walk z-ranges of chunks instead -----------
    // Each chunk of the work list becomes up to MAX_PARTS consecutive items over the steps that can hold a dirty cell;
    // a counting launch leaves every part's totals, the extraction launch looks back over them (regular_extract.cu).
    std::vector<SplitItem> items;
    const uint32_t resident = static_cast<uint32_t>(ctx->dev.sm_count) * (ctx->cfg.edge == 32 ? 3u : 1u);
    if (n_sub == 1 && n_work[0] != 0 && n_work[0] < 16u * resident && mode == MODE_EXTRACT && ctx->debug_mode == 0 &&
        !(ctx->cfg.flags & HVX_CFG_FIRST_GENERATION) && !ctx->no_split) {
        const uint32_t steps_per_chunk = (ctx->cfg.edge + 2) / 2 - 1, steps_per_brick = ctx->cfg.edge / 8;
        // Fewer chunks than resident CTAs: every chunk is split.  A few waves of chunks whose last wave is less than
        // half full (3140 pages on 444 CTAs = 7.07 waves run as long as 8) can split the chunks of that last wave -- the
        // lightest, the list is heaviest first -- over the idle CTAs while the others stay whole (one walk).  Measured
        // on the planet set's per-rank shards (tools/probe_planet_shard.py) that LOSES: 0.364 instead of 0.323 ms for
        // 3141 pages, 0.695 instead of 0.615 ms for 6281 -- every walk of the item-list kernel starts with a dependent
        // global read of its item in three warp roles, which costs the whole chunks more than the last wave gains.
        // It stays reachable for A/B runs and tests (hvx_debug_set_mode bit 9) and is off otherwise.
        uint32_t parts_wanted = 1, first_split = 0;
        if (n_work[0] < resident) {
            parts_wanted = std::min(MAX_PARTS, std::max(1u, 2u * resident / n_work[0]));
        } else if (ctx->split_last_wave) {
            const uint32_t rem = n_work[0] % resident;
            if (rem != 0 && 2u * rem <= resident) {
                parts_wanted = std::min(MAX_PARTS, resident / rem);
                first_split = n_work[0] - rem;
            }
        }
        if (parts_wanted >= 2) {
            for (uint32_t w = 0; w < n_work[0]; ++w) {
                const uint32_t chunk = order.empty() ? w : order[w];
                const uint64_t dirty = descs[chunk].dirty_microbricks;
                uint32_t lo = steps_per_chunk, hi = 1;   // first / last step that can hold a dirty cell
                for (uint32_t mz = 0; mz < 4; ++mz)
                    if ((dirty >> (16 * mz)) & 0xffffull) {
                        lo = std::min(lo, mz * steps_per_brick + 1);
                        hi = std::max(hi, (mz + 1) * steps_per_brick);
                    }
                const uint32_t span = hi - lo + 1, parts = w < first_split ? 1u : std::min(parts_wanted, span);
                for (uint32_t q = 0; q < parts; ++q)
                    items.push_back({chunk, static_cast<uint8_t>(lo + span * q / parts), static_cast<uint8_t>(lo + span * (q + 1) / parts - 1),
                                     static_cast<uint8_t>(q), static_cast<uint8_t>(parts)});
            }
            if (items.size() > MAX_SPLIT_ITEMS) items.clear();
        }
    }
    // ---- one upload for everything the host contributes to the dispatch -----------------------------------------
    {
        auto align16 = [](uint64_t v) { return (v + 15ull) & ~15ull; };
        const uint64_t off_descs = 0, off_order = align16(off_descs + static_cast<uint64_t>(n) * sizeof(ChunkDesc));
        const uint64_t off_uniform = align16(off_order + order.size() * sizeof(uint32_t));
        const uint64_t off_items = align16(off_uniform + uniform.size() * sizeof(uint32_t));
        const uint64_t bytes = align16(off_items + items.size() * sizeof(SplitItem));
        const uint32_t flip = ctx->batch_flip ^= 1u;
        HVX_CUDA(ctx, cudaEventSynchronize(ctx->batch_done[flip]));  // the copy that last read this staging buffer (two dispatches ago)
        unsigned char* h = ctx->h_batch[flip];
        static_assert(sizeof(hvx_chunk_desc) == sizeof(ChunkDesc), "desc layout");
        memcpy(h + off_descs, descs, static_cast<size_t>(n) * sizeof(ChunkDesc));
        if (!order.empty()) memcpy(h + off_order, order.data(), order.size() * sizeof(uint32_t));
        if (!uniform.empty()) memcpy(h + off_uniform, uniform.data(), uniform.size() * sizeof(uint32_t));
        if (!items.empty()) memcpy(h + off_items, items.data(), items.size() * sizeof(SplitItem));
        HVX_CUDA(ctx, cudaMemcpyAsync(ctx->d_batch, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
        HVX_CUDA(ctx, cudaEventRecord(ctx->batch_done[flip], ctx->stream));
        ctx->d_descs = reinterpret_cast<ChunkDesc*>(ctx->d_batch + off_descs);
        ctx->d_order = reinterpret_cast<uint32_t*>(ctx->d_batch + off_order);
        ctx->d_uniform = reinterpret_cast<uint32_t*>(ctx->d_batch + off_uniform);
        ctx->d_items = reinterpret_cast<SplitItem*>(ctx->d_batch + off_items);
    }
    ctx->n_regular = n;

    RegularParams base{};
    base.mode = mode;
    if (ctx->debug_mode == 1) base.mode = MODE_STREAM_ONLY;  // hvx_debug_set_mode: roofline probes, no compute
    else if (ctx->debug_mode == 2) base.mode = MODE_BITS_ONLY;
    base.first_generation = (ctx->cfg.flags & HVX_CFG_FIRST_GENERATION) ? 1u : 0u;
    base.max_vertices = ctx->cfg.max_vertices;
    base.max_indices = ctx->cfg.max_indices;
    base.work_counter = ctx->d_work;
    const uint64_t cell_count = cells(ctx->cfg.edge);
    auto* const all_vertices = static_cast<hvx_vertex*>(ctx->buf[HVX_BUF_REGULAR_VERTICES]);
    auto* const all_indices = static_cast<uint32_t*>(ctx->buf[HVX_BUF_REGULAR_INDICES]);
    auto* const all_counters = static_cast<hvx_emission_counters*>(ctx->buf[HVX_BUF_REGULAR_COUNTERS]);
    auto* const all_classify = static_cast<hvx_classify_counters*>(ctx->buf[HVX_BUF_REGULAR_CLASSIFY]);
    auto* const all_ranges = static_cast<hvx_range*>(ctx->buf[HVX_BUF_REGULAR_RANGES]);
    auto* const all_cells = static_cast<hvx_cell_record*>(ctx->buf[HVX_BUF_REGULAR_CELLS]);
    auto* const all_offsets = static_cast<hvx_cell_offset*>(ctx->buf[HVX_BUF_REGULAR_OFFSETS]);
    auto* const all_blocks = static_cast<hvx_scan_block*>(ctx->buf[HVX_BUF_REGULAR_BLOCKS]);

    if (out && !ctx->h_ranges) {
        HVX_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_ranges), static_cast<size_t>(ctx->cfg.max_chunks) * sizeof(hvx_range), cudaHostAllocDefault));
    }
    if (from_host) {
        // the arena may still be read by work queued earlier on the ctx stream
        HVX_CUDA(ctx, cudaEventRecord(ctx->handover, ctx->stream));
        HVX_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->handover, 0));
    }
    for (uint32_t k = 0; k < n_sub; ++k) {
        const uint32_t first = k * sub, count = std::min(sub, n - first);
        if (from_host) {
            // runs of consecutive chunks that are not flagged uniform, one copy each
            uint32_t i = 0;
            while (i < count) {
                while (i < count && (descs[first + i].flags & HVX_CHUNK_UNIFORM)) ++i;
                uint32_t j = i;
                while (j < count && !(descs[first + j].flags & HVX_CHUNK_UNIFORM)) ++j;
                if (j > i) {
                    const uint64_t off = (static_cast<uint64_t>(first) + i) * chunk_words;
                    HVX_CUDA(ctx, cudaMemcpyAsync(static_cast<uint32_t*>(ctx->buf[HVX_BUF_SAMPLES]) + off, samples + off,
                                                  static_cast<uint64_t>(j - i) * chunk_words * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
                }
                i = j;
            }
            HVX_CUDA(ctx, cudaEventRecord(ctx->events[2 * k], ctx->copy_stream));
            HVX_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->events[2 * k], 0));
        }
        RegularParams p = base;
        p.samples = d_samples + static_cast<uint64_t>(first) * chunk_words;
        p.descs = ctx->d_descs + first;
        p.order = order.empty() ? nullptr : ctx->d_order + first;
        p.n_chunks = count;
        p.n_work = n_work[k];
        p.skipped = ctx->d_uniform + skipped_begin[k];
        p.n_skipped = skipped_begin[k + 1] - skipped_begin[k];
        if (!items.empty()) {   // the items are already in start order
            p.order = nullptr;
            p.items = ctx->d_items;
            p.item_totals = ctx->d_item_totals;
            p.split_generation = ++ctx->split_generation;
            p.n_work = static_cast<uint32_t>(items.size());
        }
        p.chunk_base = first;
        for (uint32_t i = 0; i < count && !p.any_partial; ++i)
            p.any_partial = descs[first + i].dirty_microbricks != ~0ull && descs[first + i].dirty_microbricks != 0 &&
                            !(descs[first + i].flags & HVX_CHUNK_UNIFORM);
        p.vertices = all_vertices + static_cast<uint64_t>(first) * ctx->cfg.max_vertices;
        p.indices = all_indices + static_cast<uint64_t>(first) * ctx->cfg.max_indices;
        p.counters = all_counters + first;
        p.classify = all_classify + first;
        p.ranges = all_ranges + first;
        p.cells = all_cells ? all_cells + static_cast<uint64_t>(first) * cell_count : nullptr;
        p.offsets = all_offsets ? all_offsets + static_cast<uint64_t>(first) * cell_count : nullptr;
        p.blocks = all_blocks ? all_blocks + static_cast<uint64_t>(first) * (cell_count / 256) : nullptr;
        if (p.n_work != 0) {
            cudaError_t e = launch_regular(static_cast<int>(ctx->cfg.edge), p, ctx->dev, ctx->stream);
            if (e != cudaSuccess) return cuda_fail(ctx, e, "launch_regular");
            ctx->launches += ((p.cells != nullptr && !p.first_generation) ? 3 : 1) ;  // + the two record kernels
        } else if (p.cells != nullptr) {  // nothing to extract, but the scan blocks of the sub-batch still report it
            cudaError_t e = launch_regular_records(static_cast<int>(ctx->cfg.edge), p, ctx->stream);
            if (e != cudaSuccess) return cuda_fail(ctx, e, "launch_regular_records");
            ctx->launches += 2;
        }
        if (p.n_skipped != 0 && (p.n_work == 0 || p.first_generation)) {
            // no decoupled launch to write the skipped chunks' records on the side: a small launch of its own
            cudaError_t e = launch_uniform_records(static_cast<int>(ctx->cfg.edge), p, p.skipped, p.n_skipped, ctx->stream);
            if (e != cudaSuccess) return cuda_fail(ctx, e, "launch_uniform_records");
            ctx->launches += 1;
        }
        if (out) {
            HVX_CUDA(ctx, cudaMemcpyAsync(ctx->h_ranges + first, all_ranges + first, static_cast<size_t>(count) * sizeof(hvx_range),
                                          cudaMemcpyDeviceToHost, ctx->stream));
            HVX_CUDA(ctx, cudaEventRecord(ctx->events[2 * k + 1], ctx->stream));
        }
    }
    if (!out) return HVX_OK;

    // ---- packed read-back, sub-batch by sub-batch, behind the kernels ------------------------------------------
    uint64_t tv = 0, ti = 0;
    bool fits = true;
    std::vector<hvx_range> local(std::min(sub, n));
    for (uint32_t k = 0; k < n_sub; ++k) {
        const uint32_t first = k * sub, count = std::min(sub, n - first);
        HVX_CUDA(ctx, cudaEventSynchronize(ctx->events[2 * k + 1]));
        uint64_t sv = 0, si = 0;
        for (uint32_t i = 0; i < count; ++i) {
            const hvx_range r = ctx->h_ranges[first + i];
            out->ranges[first + i] = {static_cast<uint32_t>(tv + sv), r.vertex_count, static_cast<uint32_t>(ti + si), r.index_count};
            local[i] = {static_cast<uint32_t>(sv), r.vertex_count, static_cast<uint32_t>(si), r.index_count};
            sv += r.vertex_count;
            si += r.index_count;
        }
        fits = fits && tv + sv <= out->vertex_cap && ti + si <= out->index_cap && tv + sv <= 0xffffffffull && ti + si <= 0xffffffffull &&
               (sv == 0 || out->vertices) && (si == 0 || out->indices);
        if (fits && (sv || si)) {
            if ((rc = grow_device(ctx, &ctx->pack_v, &ctx->pack_v_bytes, sv * sizeof(hvx_vertex), ctx->d2h_stream))) return rc;
            if ((rc = grow_device(ctx, &ctx->pack_i, &ctx->pack_i_bytes, si * 4, ctx->d2h_stream))) return rc;
            HVX_CUDA(ctx, cudaMemcpyAsync(ctx->d_packed + first, local.data(), static_cast<size_t>(count) * sizeof(hvx_range), cudaMemcpyHostToDevice, ctx->d2h_stream));
            cudaError_t e = launch_pack(all_vertices, all_indices, all_ranges + first, ctx->d_packed + first, count,
                                        static_cast<hvx_vertex*>(ctx->pack_v), static_cast<uint32_t*>(ctx->pack_i), ctx->dev, ctx->d2h_stream);
            if (e != cudaSuccess) return cuda_fail(ctx, e, "launch_pack");
            ctx->launches += 1;
            if (sv) HVX_CUDA(ctx, cudaMemcpyAsync(out->vertices + tv, ctx->pack_v, sv * sizeof(hvx_vertex), cudaMemcpyDeviceToHost, ctx->d2h_stream));
            if (si) HVX_CUDA(ctx, cudaMemcpyAsync(out->indices + ti, ctx->pack_i, si * 4, cudaMemcpyDeviceToHost, ctx->d2h_stream));
        }
        tv += sv;
        ti += si;
    }
    // every kernel of the dispatch has finished (the last event was waited for)
    if (out->counters)
        HVX_CUDA(ctx, cudaMemcpyAsync(out->counters, all_counters, static_cast<size_t>(n) * sizeof(hvx_emission_counters), cudaMemcpyDeviceToHost, ctx->d2h_stream));
    HVX_CUDA(ctx, cudaStreamSynchronize(ctx->d2h_stream));
    out->total_vertices = tv;
    out->total_indices = ti;
    if (tv > 0xffffffffull || ti > 0xffffffffull) return fail(ctx, HVX_E_DEVICE_LIMIT, "DeviceLimit: packed mesh exceeds 32-bit ranges");
    if (!fits)
        return fail(ctx, HVX_E_INVALID_CAPACITY, "output capacity too small: need %llu vertices / %llu indices",
                    static_cast<unsigned long long>(tv), static_cast<unsigned long long>(ti));
    return HVX_OK;
}

int validate_pages(hvx_ctx* ctx, const int64_t* page_xyz, const uint8_t* lod, uint32_t n, bool slabs) {
    const int64_t edge = ctx->cfg.edge;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t l = lod ? lod[i] : 0;
        if (l > 57) return fail(ctx, HVX_E_ADDRESS, "planetary page LOD %u exceeds the addressable maximum", l);
        if (slabs && l == 0)
            return fail(ctx, HVX_E_FINEST_LOD, "LOD0 has no finer neighbor and cannot own a transition mesh");
        const int64_t span = edge << l, scale = 1ll << l;
        for (int a = 0; a < 3; ++a) {
            int64_t lo, hi, tmp;
            if (__builtin_mul_overflow(page_xyz[3 * i + a], span, &lo) ||
                __builtin_add_overflow(lo, span, &hi) || __builtin_add_overflow(hi, 2 * scale, &tmp) ||
                __builtin_sub_overflow(lo, 2 * scale, &tmp))
                return fail(ctx, HVX_E_ADDRESS, "planetary page coordinate arithmetic overflowed");
        }
    }
    return HVX_OK;
}

int run_fill(hvx_ctx* ctx, uint32_t kind, const int64_t* page_xyz, const uint8_t* lod, uint32_t n, uint32_t* out,
             bool slabs) {
    if (!ctx) return HVX_E_INVALID_ARGUMENT;
    if (n > ctx->cfg.max_chunks)
        return fail(ctx, HVX_E_BATCH_CAPACITY, "batch of %u chunks exceeds max_chunks %u", n, ctx->cfg.max_chunks);
    if (!(kind <= 5 || kind == 16 || kind == 17)) return fail(ctx, HVX_E_INVALID_ARGUMENT, "unknown field kind %u", kind);
    if (n == 0) return HVX_OK;
    if (!page_xyz) return fail(ctx, HVX_E_INVALID_ARGUMENT, "page_xyz is NULL");
    int rc = validate_pages(ctx, page_xyz, lod, n, slabs);
    if (rc) return rc;
    DeviceGuard guard(ctx->device);
    const int arena = slabs ? HVX_BUF_SLABS : HVX_BUF_SAMPLES;
    if (!out) {
        if ((rc = ensure_buffer(ctx, arena))) return rc;
        out = static_cast<uint32_t*>(ctx->buf[arena]);
    }
    ctx->edit_pages.clear();  // d_pages / d_lod are about to hold this fill's list
    ctx->edit_lods.clear();
    HVX_CUDA(ctx, cudaMemcpyAsync(ctx->d_pages, page_xyz, static_cast<size_t>(n) * 3 * sizeof(int64_t),
                                  cudaMemcpyHostToDevice, ctx->stream));
    if (lod) HVX_CUDA(ctx, cudaMemcpyAsync(ctx->d_lod, lod, n, cudaMemcpyHostToDevice, ctx->stream));
    else HVX_CUDA(ctx, cudaMemsetAsync(ctx->d_lod, 0, n, ctx->stream));
    FillParams p{kind, n, ctx->d_pages, ctx->d_lod, out, nullptr, nullptr};
    if (kind == 16 && !slabs) {
        // the fBm height depends only on (x, z): evaluate it once per distinct column of the batch
        // (open-addressing hash on (x, z, lod); the batch is at most max_chunks long)
        uint32_t table_size = 16;
        while (table_size < 2u * n) table_size <<= 1;
        std::vector<uint32_t> table(table_size, UINT32_MAX);
        std::vector<uint32_t> col_index(n);
        std::vector<long long> col_xz;
        std::vector<uint8_t> col_lod;
        col_xz.reserve(2ull * n);
        col_lod.reserve(n);
        for (uint32_t i = 0; i < n; ++i) {
            const uint8_t l = lod ? lod[i] : 0;
            const long long x = page_xyz[3 * i], z = page_xyz[3 * i + 2];
            uint64_t h = static_cast<uint64_t>(x) * 0x9E3779B97F4A7C15ull ^ static_cast<uint64_t>(z) * 0xC2B2AE3D27D4EB4Full ^ l;
            h ^= h >> 29;
            uint32_t slot = static_cast<uint32_t>(h) & (table_size - 1);
            for (;; slot = (slot + 1) & (table_size - 1)) {
                const uint32_t c = table[slot];
                if (c == UINT32_MAX) {
                    table[slot] = static_cast<uint32_t>(col_lod.size());
                    col_index[i] = table[slot];
                    col_xz.push_back(x);
                    col_xz.push_back(z);
                    col_lod.push_back(l);
                    break;
                }
                if (col_xz[2 * c] == x && col_xz[2 * c + 1] == z && col_lod[c] == l) {
                    col_index[i] = c;
                    break;
                }
            }
        }
        const uint64_t n_cols = col_lod.size(), s = ctx->cfg.edge + 2;
        if (n_cols > ctx->heights_cols) {
            if (ctx->d_heights) {
                HVX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                cudaFree(ctx->d_heights);
                ctx->allocated -= ctx->heights_cols * s * s * sizeof(float);
                ctx->d_heights = nullptr;
                ctx->heights_cols = 0;
            }
            if ((rc = small_alloc(ctx, &ctx->d_heights, n_cols * s * s))) return rc;
            ctx->heights_cols = n_cols;
        }
        HVX_CUDA(ctx, cudaMemcpyAsync(ctx->d_col_index, col_index.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        HVX_CUDA(ctx, cudaMemcpyAsync(ctx->d_col_xz, col_xz.data(), col_xz.size() * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
        HVX_CUDA(ctx, cudaMemcpyAsync(ctx->d_col_lod, col_lod.data(), n_cols, cudaMemcpyHostToDevice, ctx->stream));
        cudaError_t eh = launch_terrain_heights(static_cast<int>(ctx->cfg.edge), ctx->d_col_xz, ctx->d_col_lod,
                                                static_cast<uint32_t>(n_cols), ctx->d_heights, ctx->stream);
        if (eh != cudaSuccess) return cuda_fail(ctx, eh, "launch_terrain_heights");
        ctx->launches += 1;
        p.heights = ctx->d_heights;
        p.col_index = ctx->d_col_index;
    }
    cudaError_t e = slabs ? launch_fill_slabs(static_cast<int>(ctx->cfg.edge), p, ctx->dev, ctx->stream)
                          : launch_fill_samples(static_cast<int>(ctx->cfg.edge), p, ctx->dev, ctx->stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "launch_fill");
    ctx->launches += 1;
    return HVX_OK;
}

}  // namespace

extern "C" {

uint32_t hvx_abi_version(void) { return HVX_ABI_VERSION; }

const char* hvx_status_name(int status) {
    switch (status) {
        case HVX_OK: return "HVX_OK";
        case HVX_E_SAMPLE_COUNT: return "HVX_E_SAMPLE_COUNT";
        case HVX_E_INVALID_CAPACITY: return "HVX_E_INVALID_CAPACITY";
        case HVX_E_DEVICE_LIMIT: return "HVX_E_DEVICE_LIMIT";
        case HVX_E_TRANSITION_MASK: return "HVX_E_TRANSITION_MASK";
        case HVX_E_INVALID_ARGUMENT: return "HVX_E_INVALID_ARGUMENT";
        case HVX_E_CUDA: return "HVX_E_CUDA";
        case HVX_E_BATCH_CAPACITY: return "HVX_E_BATCH_CAPACITY";
        case HVX_E_FINEST_LOD: return "HVX_E_FINEST_LOD";
        case HVX_E_ADDRESS: return "HVX_E_ADDRESS";
        case HVX_E_TOPOLOGY_EMPTY: return "HVX_E_TOPOLOGY_EMPTY";
        case HVX_E_TOPOLOGY_DUPLICATE: return "HVX_E_TOPOLOGY_DUPLICATE";
        case HVX_E_TOPOLOGY_OVERLAP: return "HVX_E_TOPOLOGY_OVERLAP";
        case HVX_E_TOPOLOGY_UNBALANCED: return "HVX_E_TOPOLOGY_UNBALANCED";
        case HVX_E_TOPOLOGY_ROOT_LOD: return "HVX_E_TOPOLOGY_ROOT_LOD";
        case HVX_E_TOPOLOGY_MINIMUM_LOD: return "HVX_E_TOPOLOGY_MINIMUM_LOD";
        case HVX_E_TOPOLOGY_PAGE_BUDGET: return "HVX_E_TOPOLOGY_PAGE_BUDGET";
        case HVX_E_TOPOLOGY_MISSING_PARENT: return "HVX_E_TOPOLOGY_MISSING_PARENT";
        case HVX_E_TOPOLOGY_COVERAGE: return "HVX_E_TOPOLOGY_COVERAGE";
        case HVX_E_INVALID_LIMITS: return "HVX_E_INVALID_LIMITS";
        case HVX_E_ARITHMETIC_OVERFLOW: return "HVX_E_ARITHMETIC_OVERFLOW";
        case HVX_E_NON_TRIANGLE_INDEX_COUNT: return "HVX_E_NON_TRIANGLE_INDEX_COUNT";
        case HVX_E_INCOMPLETE_SURFACE_COUNTS: return "HVX_E_INCOMPLETE_SURFACE_COUNTS";
        case HVX_E_PENDING_CAPACITY: return "HVX_E_PENDING_CAPACITY";
        case HVX_E_ARENA_CAPACITY: return "HVX_E_ARENA_CAPACITY";
        case HVX_E_GENERATION_CONFLICT: return "HVX_E_GENERATION_CONFLICT";
        case HVX_E_RESERVATION_MISSING: return "HVX_E_RESERVATION_MISSING";
        case HVX_E_RESERVATION_MISMATCH: return "HVX_E_RESERVATION_MISMATCH";
        case HVX_E_DEVICE_BUFFER_LIMIT: return "HVX_E_DEVICE_BUFFER_LIMIT";
        default: return "HVX_E_UNKNOWN";
    }
}

const char* hvx_last_error(const hvx_ctx* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

int hvx_create(hvx_ctx** out, int device, const hvx_config* config) {
    if (!out || !config) return fail(nullptr, HVX_E_INVALID_ARGUMENT, "hvx_create: NULL argument");
    *out = nullptr;
    const hvx_config c = *config;
    if (c.edge != 32 && c.edge != 64) return fail(nullptr, HVX_E_INVALID_ARGUMENT, "edge must be 32 or 64, got %u", c.edge);
    if (c.max_chunks == 0) return fail(nullptr, HVX_E_INVALID_ARGUMENT, "max_chunks must be nonzero");
    if (c.flags & ~(HVX_CFG_DEBUG_RECORDS | HVX_CFG_FIRST_GENERATION))
        return fail(nullptr, HVX_E_INVALID_ARGUMENT, "unknown configuration flags %#x", c.flags);
    // TransvoxelGpuExtractorConfig::new, PV/src/transvoxel_emit.rs:63-75
    if (c.max_vertices == 0 || c.max_indices == 0)
        return fail(nullptr, HVX_E_INVALID_CAPACITY,
                    "Transvoxel extraction capacities must be nonzero (vertices=%u, indices=%u)", c.max_vertices,
                    c.max_indices);
    // TransvoxelGpuTransitionExtractorConfig::new, PV/src/transvoxel_transition_gpu.rs:160-171
    if ((c.max_transition_vertices == 0) != (c.max_transition_indices == 0))
        return fail(nullptr, HVX_E_INVALID_CAPACITY,
                    "Transvoxel transition capacities must be nonzero (vertices=%u, indices=%u)",
                    c.max_transition_vertices, c.max_transition_indices);
    if (static_cast<uint64_t>(c.max_chunks) * c.max_vertices > 0xffffffffull ||
        static_cast<uint64_t>(c.max_chunks) * c.max_indices > 0xffffffffull ||
        static_cast<uint64_t>(c.max_chunks) * c.max_transition_vertices > 0xffffffffull ||
        static_cast<uint64_t>(c.max_chunks) * c.max_transition_indices > 0xffffffffull)
        return fail(nullptr, HVX_E_DEVICE_LIMIT, "DeviceLimit: max_chunks * per-chunk capacity must fit 32-bit ranges");

    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(nullptr, HVX_E_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= count) return fail(nullptr, HVX_E_INVALID_ARGUMENT, "device %d out of range (%d devices)", device, count);
    hvx_ctx* ctx = new (std::nothrow) hvx_ctx();
    if (!ctx) return fail(nullptr, HVX_E_INVALID_ARGUMENT, "out of host memory");
    ctx->cfg = c;
    ctx->device = device;
    DeviceGuard guard(device);
    auto bail = [&](int rc) {
        g_create_error = ctx->error;
        hvx_destroy(ctx);
        return rc;
    };
    cudaDeviceProp prop{};
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail(cuda_fail(ctx, e, "cudaGetDeviceProperties"));
    if (prop.major < 10)
        return bail(fail(ctx, HVX_E_DEVICE_LIMIT, "DeviceLimit: requires compute capability 10.0 (sm_100a), device is %d.%d",
                         prop.major, prop.minor));
    ctx->dev.ordinal = device;
    ctx->dev.sm_count = prop.multiProcessorCount;
    ctx->dev.max_smem_optin = static_cast<int>(prop.sharedMemPerBlockOptin);
    if (regular_smem_bytes(static_cast<int>(c.edge)) > prop.sharedMemPerBlockOptin)
        return bail(fail(ctx, HVX_E_DEVICE_LIMIT, "DeviceLimit: shared memory per block: required %zu, available %zu",
                         regular_smem_bytes(static_cast<int>(c.edge)), prop.sharedMemPerBlockOptin));
    if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess)
        return bail(cuda_fail(ctx, e, "cudaStreamCreate"));
    ctx->stream = ctx->own_stream;
    if ((e = cudaEventCreateWithFlags(&ctx->handover, cudaEventDisableTiming)) != cudaSuccess)
        return bail(cuda_fail(ctx, e, "cudaEventCreate"));
    int rc;
    ctx->batch_bytes = static_cast<uint64_t>(c.max_chunks) * (sizeof(ChunkDesc) + 8) + MAX_SPLIT_ITEMS * sizeof(SplitItem) + 64;
    if ((rc = small_alloc(ctx, &ctx->d_batch, ctx->batch_bytes))) return bail(rc);
    for (int i = 0; i < 2; ++i) {
        if ((e = cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_batch[i]), ctx->batch_bytes, cudaHostAllocDefault)) != cudaSuccess)
            return bail(cuda_fail(ctx, e, "cudaHostAlloc"));
        if ((e = cudaEventCreateWithFlags(&ctx->batch_done[i], cudaEventDisableTiming)) != cudaSuccess)
            return bail(cuda_fail(ctx, e, "cudaEventCreate"));
    }
    if ((rc = small_alloc(ctx, &ctx->d_touched, c.max_chunks))) return bail(rc);
    if ((e = cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_touched), static_cast<size_t>(c.max_chunks) * sizeof(uint32_t), cudaHostAllocDefault)) != cudaSuccess)
        return bail(cuda_fail(ctx, e, "cudaHostAlloc"));
    if ((e = cudaEventCreateWithFlags(&ctx->touched_done, cudaEventDisableTiming)) != cudaSuccess) return bail(cuda_fail(ctx, e, "cudaEventCreate"));
    if (c.max_transition_vertices != 0) {
        if ((rc = small_alloc(ctx, &ctx->d_tdescs, c.max_chunks))) return bail(rc);
        for (int i = 0; i < 2; ++i) {
            if ((e = cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_tbatch[i]), static_cast<size_t>(c.max_chunks) * sizeof(ChunkDesc), cudaHostAllocDefault)) != cudaSuccess)
                return bail(cuda_fail(ctx, e, "cudaHostAlloc"));
            if ((e = cudaEventCreateWithFlags(&ctx->tbatch_done[i], cudaEventDisableTiming)) != cudaSuccess)
                return bail(cuda_fail(ctx, e, "cudaEventCreate"));
        }
    }

    if ((rc = small_alloc(ctx, &ctx->d_item_totals, MAX_SPLIT_ITEMS))) return bail(rc);
    if ((e = cudaMemsetAsync(ctx->d_item_totals, 0, MAX_SPLIT_ITEMS * sizeof(uint4), ctx->stream)) != cudaSuccess)
        return bail(cuda_fail(ctx, e, "cudaMemsetAsync"));
    if ((rc = small_alloc(ctx, &ctx->d_pages, 3ull * c.max_chunks))) return bail(rc);
    if ((rc = small_alloc(ctx, &ctx->d_lod, c.max_chunks))) return bail(rc);
    if ((rc = small_alloc(ctx, &ctx->d_col_index, c.max_chunks))) return bail(rc);
    if ((rc = small_alloc(ctx, &ctx->d_col_xz, 2ull * c.max_chunks))) return bail(rc);
    if ((rc = small_alloc(ctx, &ctx->d_col_lod, c.max_chunks))) return bail(rc);
    if ((rc = small_alloc(ctx, &ctx->d_work, 8))) return bail(rc);
    if ((e = cudaMemsetAsync(ctx->d_work, 0, 8 * sizeof(uint32_t), ctx->stream)) != cudaSuccess) return bail(cuda_fail(ctx, e, "cudaMemsetAsync"));
    if ((rc = small_alloc(ctx, &ctx->d_packed, c.max_chunks))) return bail(rc);
    for (int id = HVX_BUF_REGULAR_VERTICES; id <= HVX_BUF_TRANSITION_BLOCKS; ++id)  // meshlet arenas stay lazy
        if (arena_bytes(c, id) != 0 && (rc = ensure_buffer(ctx, id))) return bail(rc);
    if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return bail(cuda_fail(ctx, e, "cudaStreamSynchronize"));
    *out = ctx;
    return HVX_OK;
}

void hvx_destroy(hvx_ctx* ctx) {
    if (!ctx) return;
    DeviceGuard guard(ctx->device);
    if (ctx->own_stream) cudaStreamSynchronize(ctx->own_stream);
    for (void*& b : ctx->buf)
        if (b) cudaFree(b);
    cudaFree(ctx->d_batch);
    cudaFree(ctx->d_touched);
    if (ctx->h_touched) cudaFreeHost(ctx->h_touched);
    if (ctx->touched_done) cudaEventDestroy(ctx->touched_done);
    for (int i = 0; i < 2; ++i) {
        if (ctx->h_batch[i]) cudaFreeHost(ctx->h_batch[i]);
        if (ctx->batch_done[i]) cudaEventDestroy(ctx->batch_done[i]);
        if (ctx->h_tbatch[i]) cudaFreeHost(ctx->h_tbatch[i]);
        if (ctx->tbatch_done[i]) cudaEventDestroy(ctx->tbatch_done[i]);
    }
    cudaFree(ctx->d_tdescs);
    cudaFree(ctx->d_pages);
    cudaFree(ctx->d_lod);
    cudaFree(ctx->d_col_index);
    cudaFree(ctx->d_col_xz);
    cudaFree(ctx->d_col_lod);
    cudaFree(ctx->d_heights);
    for (void* b : ctx->stage) cudaFree(b);
    cudaFree(ctx->bound_table);
    cudaFree(ctx->d_work);
    cudaFree(ctx->d_weld);
    cudaFree(ctx->d_packed);
    cudaFree(ctx->pack_v);
    cudaFree(ctx->pack_i);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    if (ctx->d2h_stream) cudaStreamSynchronize(ctx->d2h_stream);
    cudaFree(ctx->d_item_totals);
    if (ctx->h_ranges) cudaFreeHost(ctx->h_ranges);
    for (cudaEvent_t e : ctx->events) cudaEventDestroy(e);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
    if (ctx->handover) cudaEventDestroy(ctx->handover);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int hvx_get_config(const hvx_ctx* ctx, hvx_config* out) {
    if (!ctx || !out) return HVX_E_INVALID_ARGUMENT;
    *out = ctx->cfg;
    return HVX_OK;
}

uint64_t hvx_allocated_bytes(const hvx_ctx* ctx) { return ctx ? ctx->allocated : 0; }
uint64_t hvx_launch_count(const hvx_ctx* ctx) { return ctx ? ctx->launches : 0; }

const char* hvx_regular_kernel_name(const hvx_ctx* ctx, int partial) {
    if (!ctx) return "";
    return regular_kernel_name(static_cast<int>(ctx->cfg.edge), (ctx->cfg.flags & HVX_CFG_FIRST_GENERATION) != 0, partial != 0);
}

int hvx_set_stream(hvx_ctx* ctx, void* cuda_stream) {
    if (!ctx) return HVX_E_INVALID_ARGUMENT;
    cudaStream_t next = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    if (next == ctx->stream) return HVX_OK;
    // the ctx scratch (descriptors, work counters, staging) may still be in use by work queued on the old stream:
    // the new stream starts after it
    DeviceGuard guard(ctx->device);
    HVX_CUDA(ctx, cudaEventRecord(ctx->handover, ctx->stream));
    HVX_CUDA(ctx, cudaStreamWaitEvent(next, ctx->handover, 0));
    ctx->stream = next;
    return HVX_OK;
}

int hvx_start_order(const uint32_t* cost_hints, uint32_t n, uint32_t spread_pct, uint32_t* order_out) {
    if (n != 0 && (!cost_hints || !order_out)) return fail(nullptr, HVX_E_INVALID_ARGUMENT, "hvx_start_order: NULL argument");
    if (spread_pct > 100u) return fail(nullptr, HVX_E_INVALID_ARGUMENT, "spread_pct %u is not a percentage", spread_pct);
    for (uint32_t i = 0; i < n; ++i) order_out[i] = i;
    start_order(order_out, n, [&](uint32_t i) { return cost_hints[i]; }, spread_pct);
    return HVX_OK;
}

int hvx_debug_set_mode(hvx_ctx* ctx, uint32_t mode) {
    if (!ctx) return HVX_E_INVALID_ARGUMENT;
    if ((mode & 0xffu) > 2 || (mode & ~0xff3ffu))
        return fail(ctx, HVX_E_INVALID_ARGUMENT, "debug mode must be 0 (off), 1 (stream only) or 2 (stream + sign bits), optionally | 0x100 (no split walk) | 0x200 (split a thin last wave)");
    ctx->debug_mode = mode & 0xffu;
    ctx->no_split = (mode & 0x100u) != 0;
    ctx->split_last_wave = (mode & 0x200u) != 0;
    // bits 12..19: where the heavy chunks of a hinted batch go: 0 = default (75 per cent), 1..100 = that share of the
    // start order, 255 = plain descending order
    const uint32_t spread = (mode >> 12) & 0xffu;
    ctx->spread_pct = spread == 0 ? 75u : spread == 255u ? 0u : std::min(spread, 100u);
    return HVX_OK;
}

void* hvx_get_stream(const hvx_ctx* ctx) { return ctx ? ctx->stream : nullptr; }

int hvx_synchronize(hvx_ctx* ctx) {
    if (!ctx) return HVX_E_INVALID_ARGUMENT;
    DeviceGuard guard(ctx->device);
    HVX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return HVX_OK;
}

int hvx_fill_density(hvx_ctx* ctx, uint32_t kind, const int64_t* page_xyz, const uint8_t* lod, uint32_t n,
                     uint32_t* d_samples) {
    return run_fill(ctx, kind, page_xyz, lod, n, d_samples, false);
}

int hvx_fill_slabs(hvx_ctx* ctx, uint32_t kind, const int64_t* page_xyz, const uint8_t* lod, uint32_t n,
                   uint32_t* d_slabs) {
    return run_fill(ctx, kind, page_xyz, lod, n, d_slabs, true);
}

int hvx_apply_edit(hvx_ctx* ctx, const hvx_voxel_edit* edit, const int64_t* page_xyz, const uint8_t* lod, uint32_t n,
                   uint32_t* d_samples, uint64_t* dirty_out, uint32_t* touched_out) {
    static_assert(sizeof(hvx_voxel_edit) == 32, "GpuVoxelEdit layout");
    if (!ctx) return HVX_E_INVALID_ARGUMENT;
    if (touched_out) *touched_out = 0;
    if (!edit || (n != 0 && (!page_xyz || !dirty_out))) return fail(ctx, HVX_E_INVALID_ARGUMENT, "hvx_apply_edit: NULL argument");
    if (n > ctx->cfg.max_chunks) return fail(ctx, HVX_E_BATCH_CAPACITY, "batch of %u chunks exceeds max_chunks %u", n, ctx->cfg.max_chunks);
    if (edit->op_type != 1u && edit->op_type != 2u)
        return fail(ctx, HVX_E_INVALID_ARGUMENT, "edit op %u is not a sphere edit (1 AddSphere, 2 SubtractSphere)", edit->op_type);
    if (!(edit->radius > 0.0f) || !std::isfinite(edit->radius) || !std::isfinite(edit->center[0]) ||
        !std::isfinite(edit->center[1]) || !std::isfinite(edit->center[2]))
        return fail(ctx, HVX_E_INVALID_ARGUMENT, "edit centre and radius must be finite, radius positive");
    if (n == 0) return HVX_OK;
    int rc = validate_pages(ctx, page_xyz, lod, n, false);
    if (rc) return rc;
    // the octree rule (octree.rs:139-173), in float like the reference: a box is untouched when the centre is further
    // from the box centre than half the box plus the radius on any axis
    const int E = static_cast<int>(ctx->cfg.edge), Q = E / 4;
    auto overlaps = [&](const float lo[3], const float hi[3], float r) {
        for (int a = 0; a < 3; ++a) {
            const float c = (lo[a] + hi[a]) * 0.5f, h = (hi[a] - lo[a]) * 0.5f;
            if (std::fabs(edit->center[a] - c) > h + r) return false;
        }
        return true;
    };
    std::vector<uint32_t> touched;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t l = lod ? lod[i] : 0;
        const float cell_m = 0.1f * static_cast<float>(1ull << (l > 30 ? 30 : l));
        float lo[3], hi[3];
        for (int a = 0; a < 3; ++a) {  // the sample block, halo included
            lo[a] = static_cast<float>(page_xyz[3 * i + a] * E - 1) * cell_m;
            hi[a] = static_cast<float>(page_xyz[3 * i + a] * E + E) * cell_m;
        }
        dirty_out[i] = 0;
        if (!overlaps(lo, hi, edit->radius)) continue;
        touched.push_back(i);
        const float reach = edit->radius + 2.0f * cell_m;
        uint64_t bits = 0;
        for (int mz = 0; mz < 4; ++mz)
            for (int my = 0; my < 4; ++my)
                for (int mx = 0; mx < 4; ++mx) {
                    const int m[3] = {mx, my, mz};
                    for (int a = 0; a < 3; ++a) {
                        lo[a] = static_cast<float>(page_xyz[3 * i + a] * E + m[a] * Q) * cell_m;
                        hi[a] = static_cast<float>(page_xyz[3 * i + a] * E + (m[a] + 1) * Q) * cell_m;
                    }
                    if (overlaps(lo, hi, reach)) bits |= 1ull << (mx + 4 * my + 16 * mz);
                }
        dirty_out[i] = bits;
    }
    if (touched_out) *touched_out = static_cast<uint32_t>(touched.size());
    if (touched.empty()) return HVX_OK;
    DeviceGuard guard(ctx->device);
    if (!d_samples) {
        if ((rc = ensure_buffer(ctx, HVX_BUF_SAMPLES))) return rc;
        d_samples = static_cast<uint32_t*>(ctx->buf[HVX_BUF_SAMPLES]);
    }
    // an edit frame names the same resident pages as the frame before: they are uploaded again only when they changed
    // (compared by content; hvx_fill_* use the same device arrays, so they drop the cache)
    const bool same_pages = ctx->edit_pages.size() == 3ull * n && memcmp(ctx->edit_pages.data(), page_xyz, 3ull * n * sizeof(int64_t)) == 0 &&
                            ctx->edit_lods.size() == n && (lod ? memcmp(ctx->edit_lods.data(), lod, n) == 0
                                                               : std::all_of(ctx->edit_lods.begin(), ctx->edit_lods.end(), [](uint8_t l) { return l == 0; }));
    if (!same_pages) {
        HVX_CUDA(ctx, cudaMemcpyAsync(ctx->d_pages, page_xyz, static_cast<size_t>(n) * 3 * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
        if (lod) HVX_CUDA(ctx, cudaMemcpyAsync(ctx->d_lod, lod, n, cudaMemcpyHostToDevice, ctx->stream));
        else HVX_CUDA(ctx, cudaMemsetAsync(ctx->d_lod, 0, n, ctx->stream));
        ctx->edit_pages.assign(page_xyz, page_xyz + 3ull * n);
        if (lod) ctx->edit_lods.assign(lod, lod + n); else ctx->edit_lods.assign(n, 0);
    }
    HVX_CUDA(ctx, cudaEventSynchronize(ctx->touched_done));  // the previous edit's copy has left the pinned list
    memcpy(ctx->h_touched, touched.data(), touched.size() * sizeof(uint32_t));
    HVX_CUDA(ctx, cudaMemcpyAsync(ctx->d_touched, ctx->h_touched, touched.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    HVX_CUDA(ctx, cudaEventRecord(ctx->touched_done, ctx->stream));
    EditParams p{};
    p.op = edit->op_type;
    p.material = edit->material;
    for (int a = 0; a < 3; ++a) p.center[a] = edit->center[a];
    p.radius = edit->radius;
    p.n_touched = static_cast<uint32_t>(touched.size());
    p.ids = ctx->d_touched;
    p.page_xyz = ctx->d_pages;
    p.lod = ctx->d_lod;
    p.samples = d_samples;
    cudaError_t e = launch_edit_sphere(E, p, ctx->stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "launch_edit_sphere");
    ctx->launches += 1;
    return HVX_OK;
}

int hvx_extract_regular(hvx_ctx* ctx, const uint32_t* samples, uint64_t words, const hvx_chunk_desc* descs, uint32_t n) {
    return run_regular(ctx, samples, words, descs, n, MODE_EXTRACT);
}

int hvx_classify_regular(hvx_ctx* ctx, const uint32_t* samples, uint64_t words, const hvx_chunk_desc* descs, uint32_t n) {
    return run_regular(ctx, samples, words, descs, n, MODE_CLASSIFY);
}

int hvx_extract_regular_to_host(hvx_ctx* ctx, const uint32_t* samples, uint64_t words, const hvx_chunk_desc* descs, uint32_t n,
                                hvx_vertex* vertices_out, uint64_t vertex_cap, uint32_t* indices_out, uint64_t index_cap,
                                hvx_range* ranges_out, hvx_emission_counters* counters_out, uint64_t* total_vertices,
                                uint64_t* total_indices) {
    if (total_vertices) *total_vertices = 0;
    if (total_indices) *total_indices = 0;
    if (ctx && n != 0 && !ranges_out) return fail(ctx, HVX_E_INVALID_ARGUMENT, "ranges_out is NULL");
    HostMeshOut out{vertices_out, vertex_cap, indices_out, index_cap, ranges_out, counters_out};
    const int rc = run_regular(ctx, samples, words, descs, n, MODE_EXTRACT, &out);
    if (total_vertices) *total_vertices = out.total_vertices;
    if (total_indices) *total_indices = out.total_indices;
    return rc;
}

int hvx_extract_transition(hvx_ctx* ctx, const uint32_t* slabs, uint64_t words, const hvx_chunk_desc* descs, uint32_t n) {
    int rc = check_batch(ctx, descs, n);
    if (rc) return rc;
    if (ctx->cfg.max_transition_vertices == 0)
        return fail(ctx, HVX_E_INVALID_CAPACITY, "Transvoxel transition capacities must be nonzero (vertices=0, indices=0)");
    const uint64_t expected = static_cast<uint64_t>(n) * slab_words(ctx->cfg.edge);
    if (words != expected)
        return fail(ctx, HVX_E_SAMPLE_COUNT, "Transvoxel transition extraction received %llu scalar samples; expected %llu",
                    static_cast<unsigned long long>(words), static_cast<unsigned long long>(expected));
    for (uint32_t i = 0; i < n; ++i)
        if (descs[i].transition_mask & ~0x3fu)
            return fail(ctx, HVX_E_TRANSITION_MASK, "transition mask %#x uses bits outside the six page faces",
                        descs[i].transition_mask);
    if (n == 0) return HVX_OK;
    DeviceGuard guard(ctx->device);
    const uint32_t* d_slabs = nullptr;
    if ((rc = resolve_input(ctx, slabs, words, HVX_BUF_SLABS, &d_slabs))) return rc;
    if ((rc = upload_descs(ctx, ctx->d_tdescs, descs, n))) return rc;
    ctx->n_transition = n;
    TransitionParams p{};
    p.slabs = d_slabs;
    p.descs = ctx->d_tdescs;
    p.n_chunks = n;
    p.max_vertices = ctx->cfg.max_transition_vertices;
    p.max_indices = ctx->cfg.max_transition_indices;
    p.vertices = static_cast<hvx_vertex*>(ctx->buf[HVX_BUF_TRANSITION_VERTICES]);
    p.indices = static_cast<uint32_t*>(ctx->buf[HVX_BUF_TRANSITION_INDICES]);
    p.counters = static_cast<hvx_transition_counters*>(ctx->buf[HVX_BUF_TRANSITION_COUNTERS]);
    p.ranges = static_cast<hvx_range*>(ctx->buf[HVX_BUF_TRANSITION_RANGES]);
    p.cells = static_cast<hvx_cell_record*>(ctx->buf[HVX_BUF_TRANSITION_CELLS]);
    p.offsets = static_cast<hvx_cell_offset*>(ctx->buf[HVX_BUF_TRANSITION_OFFSETS]);
    p.blocks = static_cast<hvx_scan_block*>(ctx->buf[HVX_BUF_TRANSITION_BLOCKS]);
    p.work_counter = ctx->d_work + 1;
    cudaError_t e = launch_transition(static_cast<int>(ctx->cfg.edge), p, ctx->dev, ctx->stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "launch_transition");
    ctx->launches += 1;
    return HVX_OK;
}

// Batched inputs: a device pointer is used in place, a host pointer is staged (grow-only scratch).
static int stage_to(hvx_ctx* ctx, void** slot, uint64_t* cap, const void* ptr, uint64_t bytes, const void** out) {
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, ptr) != cudaSuccess) {
        cudaGetLastError();
        attr.type = cudaMemoryTypeUnregistered;
    }
    if (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged) {
        if (attr.type == cudaMemoryTypeDevice && attr.device != ctx->device)
            return fail(ctx, HVX_E_INVALID_ARGUMENT, "input lives on device %d, ctx is on device %d", attr.device, ctx->device);
        *out = ptr;
        return HVX_OK;
    }
    if (bytes > *cap) {
        if (*slot) {
            HVX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            cudaFree(*slot);
            ctx->allocated -= *cap;
            *slot = nullptr;
            *cap = 0;
        }
        cudaError_t e = cudaMalloc(slot, bytes);
        if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc");
        *cap = bytes;
        ctx->allocated += bytes;
    }
    HVX_CUDA(ctx, cudaMemcpyAsync(*slot, ptr, bytes, cudaMemcpyHostToDevice, ctx->stream));
    *out = *slot;
    return HVX_OK;
}
static int stage_input(hvx_ctx* ctx, int which, const void* ptr, uint64_t bytes, const void** out) {
    return stage_to(ctx, &ctx->stage[which], &ctx->stage_bytes[which], ptr, bytes, out);
}

int hvx_gather_surface(hvx_ctx* ctx, const hvx_residency* residency, const hvx_page_table_entry* table,
                       const uint32_t* atlas, uint64_t atlas_words, const hvx_gather_job* jobs, uint32_t n) {
    static_assert(sizeof(hvx_page_table_entry) == 48 && sizeof(hvx_residency) == 32 && sizeof(hvx_gather_job) == 64 &&
                      sizeof(hvx_gather_counters) == 32, "gather POD layouts");
    if (!ctx) return HVX_E_INVALID_ARGUMENT;
    if (ctx->cfg.edge != 32) return fail(ctx, HVX_E_INVALID_ARGUMENT, "surface gather needs edge 32 (the residency page edge), ctx has %u", ctx->cfg.edge);
    if (n > ctx->cfg.max_chunks) return fail(ctx, HVX_E_BATCH_CAPACITY, "batch of %u chunks exceeds max_chunks %u", n, ctx->cfg.max_chunks);
    if (n == 0) return HVX_OK;
    if (!residency || !atlas || !jobs) return fail(ctx, HVX_E_INVALID_ARGUMENT, "residency, atlas and jobs must be non-NULL");
    if (!table && !ctx->bound_table) return fail(ctx, HVX_E_INVALID_ARGUMENT, "table is NULL and no table is bound (hvx_gather_bind_table)");
    const hvx_residency& r = *residency;
    if ((r.table_mask & (r.table_mask + 1u)) != 0u) return fail(ctx, HVX_E_INVALID_ARGUMENT, "page table capacity must be a power of two (mask %#x)", r.table_mask);
    if (r.max_probe == 0u || r.max_probe > r.table_mask + 1u) return fail(ctx, HVX_E_INVALID_ARGUMENT, "max_probe %u must be in [1, capacity %u]", r.max_probe, r.table_mask + 1u);
    const uint64_t tiles = static_cast<uint64_t>(r.atlas_tiles_x) * r.atlas_tiles_y * r.atlas_tiles_z;
    if (tiles == 0 || atlas_words != tiles * 32768ull || atlas_words > 0xffffffffull)
        return fail(ctx, HVX_E_SAMPLE_COUNT, "atlas holds %llu words; %u x %u x %u tiles of 32^3 need %llu",
                    static_cast<unsigned long long>(atlas_words), r.atlas_tiles_x, r.atlas_tiles_y, r.atlas_tiles_z,
                    static_cast<unsigned long long>(tiles * 32768ull));
    bool any_transition = false;
    for (uint32_t i = 0; i < n; ++i) {
        if (jobs[i].transition_mask & ~0x3fu)
            return fail(ctx, HVX_E_TRANSITION_MASK, "transition mask %#x uses bits outside the six page faces", jobs[i].transition_mask);
        if (jobs[i].lod == 0 && jobs[i].transition_mask != 0)
            return fail(ctx, HVX_E_FINEST_LOD, "LOD0 pages cannot own coarse-side transition faces");
        if (jobs[i].lod > 24) return fail(ctx, HVX_E_ADDRESS, "lod %u is outside the addressable range", jobs[i].lod);
        any_transition |= jobs[i].transition_mask != 0;
    }
    DeviceGuard guard(ctx->device);
    int rc;
    if ((rc = ensure_buffer(ctx, HVX_BUF_SAMPLES))) return rc;
    if (any_transition && (rc = ensure_buffer(ctx, HVX_BUF_SLABS))) return rc;
    if ((rc = ensure_buffer(ctx, HVX_BUF_GATHER_COUNTERS))) return rc;
    if ((rc = ensure_buffer(ctx, HVX_BUF_GATHER_INDIRECT))) return rc;
    const void *d_table = nullptr, *d_atlas = nullptr, *d_jobs = nullptr;
    const uint64_t table_bytes = (static_cast<uint64_t>(r.table_mask) + 1) * sizeof(hvx_page_table_entry);
    if (table == nullptr) {
        if (ctx->bound_table_bytes != table_bytes)
            return fail(ctx, HVX_E_INVALID_ARGUMENT, "table is NULL and the bound table holds %llu bytes, the residency uniform needs %llu",
                        static_cast<unsigned long long>(ctx->bound_table_bytes), static_cast<unsigned long long>(table_bytes));
        d_table = ctx->bound_table;
    } else if ((rc = stage_input(ctx, 0, table, table_bytes, &d_table))) {
        return rc;
    }
    if ((rc = stage_input(ctx, 1, atlas, atlas_words * 4, &d_atlas))) return rc;
    if ((rc = stage_input(ctx, 2, jobs, static_cast<uint64_t>(n) * sizeof(hvx_gather_job), &d_jobs))) return rc;
    HVX_CUDA(ctx, cudaMemsetAsync(ctx->buf[HVX_BUF_GATHER_COUNTERS], 0, static_cast<size_t>(n) * sizeof(hvx_gather_counters), ctx->stream));
    HVX_CUDA(ctx, cudaMemsetAsync(ctx->buf[HVX_BUF_GATHER_INDIRECT], 0, static_cast<size_t>(n) * 24 * 4, ctx->stream));
    GatherParams p{};
    p.n_jobs = n;
    p.residency = r;
    p.table = static_cast<const hvx_page_table_entry*>(d_table);
    p.atlas = static_cast<const uint32_t*>(d_atlas);
    p.jobs = static_cast<const hvx_gather_job*>(d_jobs);
    p.samples = static_cast<uint32_t*>(ctx->buf[HVX_BUF_SAMPLES]);
    p.slabs = static_cast<uint32_t*>(ctx->buf[HVX_BUF_SLABS]);
    p.counters = static_cast<hvx_gather_counters*>(ctx->buf[HVX_BUF_GATHER_COUNTERS]);
    p.indirect = static_cast<uint32_t*>(ctx->buf[HVX_BUF_GATHER_INDIRECT]);
    cudaError_t e = launch_gather(p, ctx->dev, ctx->stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "launch_gather");
    ctx->launches += 2;
    return HVX_OK;
}

int hvx_gather_bind_table(hvx_ctx* ctx, const hvx_page_table_entry* table, uint32_t entries) {
    if (!ctx) return HVX_E_INVALID_ARGUMENT;
    DeviceGuard guard(ctx->device);
    if (ctx->bound_table) {
        HVX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->bound_table);
        ctx->allocated -= ctx->bound_table_bytes;
        ctx->bound_table = nullptr;
        ctx->bound_table_bytes = 0;
    }
    if (!table || entries == 0) return HVX_OK;  // unbind
    if ((entries & (entries - 1)) != 0) return fail(ctx, HVX_E_INVALID_ARGUMENT, "page table capacity must be a power of two, got %u", entries);
    const uint64_t bytes = static_cast<uint64_t>(entries) * sizeof(hvx_page_table_entry);
    cudaError_t e = cudaMalloc(&ctx->bound_table, bytes);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc(page table)");
    ctx->bound_table_bytes = bytes;
    ctx->allocated += bytes;
    // host or device source; the caller's array may change as soon as this returns
    HVX_CUDA(ctx, cudaMemcpyAsync(ctx->bound_table, table, bytes, cudaMemcpyDefault, ctx->stream));
    HVX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return HVX_OK;
}

// ---- surface publication ---------------------------------------------------------------
struct hvx_publisher {
    hvx_ctx* ctx = nullptr;
    uint32_t slots = 0;
    void* buf[HVX_PUB_COUNT] = {};
    uint64_t bytes[HVX_PUB_COUNT] = {};
    void* stage[4] = {nullptr, nullptr, nullptr, nullptr};  // jobs, job_chunk, page metadata, draw pages
    uint64_t stage_bytes[4] = {0, 0, 0, 0};
};

int hvx_publisher_create(hvx_ctx* ctx, uint32_t slots, hvx_publisher** out) {
    static_assert(sizeof(hvx_surface_job) == 48 && sizeof(hvx_page_meta) == 32 && sizeof(hvx_surface_state) == 48 &&
                      sizeof(hvx_draw_page) == 48 && sizeof(hvx_surface_feedback) == 32 && sizeof(hvx_draw_indexed_indirect) == 20,
                  "publication POD layouts");
    if (!ctx || !out) return HVX_E_INVALID_ARGUMENT;
    *out = nullptr;
    if (slots == 0) return fail(ctx, HVX_E_INVALID_CAPACITY, "a publisher needs at least one slot");
    const hvx_config& c = ctx->cfg;
    const uint64_t banks = 2ull * slots;
    if (banks * c.max_vertices > 0x7fffffffull || banks * c.max_indices > 0xffffffffull ||
        banks * c.max_transition_vertices > 0x7fffffffull || banks * c.max_transition_indices > 0xffffffffull)
        return fail(ctx, HVX_E_DEVICE_LIMIT, "DeviceLimit: %u slots x 2 banks exceed 32-bit draw arguments", slots);
    DeviceGuard guard(ctx->device);
    hvx_publisher* pub = new (std::nothrow) hvx_publisher;
    if (!pub) return fail(ctx, HVX_E_DEVICE_LIMIT, "out of host memory");
    pub->ctx = ctx;
    pub->slots = slots;
    pub->bytes[HVX_PUB_REGULAR_VERTICES] = banks * c.max_vertices * sizeof(hvx_vertex);
    pub->bytes[HVX_PUB_REGULAR_INDICES] = banks * c.max_indices * 4ull;
    pub->bytes[HVX_PUB_TRANSITION_VERTICES] = banks * c.max_transition_vertices * sizeof(hvx_vertex);
    pub->bytes[HVX_PUB_TRANSITION_INDICES] = banks * c.max_transition_indices * 4ull;
    pub->bytes[HVX_PUB_STATES] = slots * sizeof(hvx_surface_state);
    pub->bytes[HVX_PUB_REGULAR_DRAWS] = slots * sizeof(hvx_draw_indexed_indirect);
    pub->bytes[HVX_PUB_TRANSITION_DRAWS] = slots * sizeof(hvx_draw_indexed_indirect);
    pub->bytes[HVX_PUB_FEEDBACK] = sizeof(hvx_surface_feedback);
    for (int i = 0; i < HVX_PUB_COUNT; ++i) {
        if (pub->bytes[i] == 0) continue;
        cudaError_t e = cudaMalloc(&pub->buf[i], pub->bytes[i]);
        if (e == cudaSuccess && i >= HVX_PUB_STATES) e = cudaMemsetAsync(pub->buf[i], 0, pub->bytes[i], ctx->stream);
        if (e != cudaSuccess) {
            cudaGetLastError();
            const int rc = fail(ctx, HVX_E_DEVICE_LIMIT, "DeviceLimit: publisher buffer %d requires %llu bytes: %s", i,
                                static_cast<unsigned long long>(pub->bytes[i]), cudaGetErrorString(e));
            hvx_publisher_destroy(pub);
            return rc;
        }
        ctx->allocated += pub->bytes[i];
    }
    *out = pub;
    return HVX_OK;
}

void hvx_publisher_destroy(hvx_publisher* pub) {
    if (!pub) return;
    DeviceGuard guard(pub->ctx->device);
    cudaStreamSynchronize(pub->ctx->stream);
    for (int i = 0; i < HVX_PUB_COUNT; ++i)
        if (pub->buf[i]) {
            cudaFree(pub->buf[i]);
            pub->ctx->allocated -= pub->bytes[i];
        }
    for (int i = 0; i < 4; ++i)
        if (pub->stage[i]) {
            cudaFree(pub->stage[i]);
            pub->ctx->allocated -= pub->stage_bytes[i];
        }
    delete pub;
}

static void fill_publish_params(hvx_publisher* pub, PublishParams& p) {
    hvx_ctx* ctx = pub->ctx;
    p.slots = pub->slots;
    p.has_transition = ctx->cfg.max_transition_vertices != 0 ? 1u : 0u;
    p.src_max_vertices = ctx->cfg.max_vertices;
    p.src_max_indices = ctx->cfg.max_indices;
    p.src_max_tvertices = ctx->cfg.max_transition_vertices;
    p.src_max_tindices = ctx->cfg.max_transition_indices;
    p.regular_counters = static_cast<const hvx_emission_counters*>(ctx->buf[HVX_BUF_REGULAR_COUNTERS]);
    p.transition_counters = static_cast<const hvx_transition_counters*>(ctx->buf[HVX_BUF_TRANSITION_COUNTERS]);
    p.src_vertices = static_cast<const hvx_vertex*>(ctx->buf[HVX_BUF_REGULAR_VERTICES]);
    p.src_indices = static_cast<const uint32_t*>(ctx->buf[HVX_BUF_REGULAR_INDICES]);
    p.src_tvertices = static_cast<const hvx_vertex*>(ctx->buf[HVX_BUF_TRANSITION_VERTICES]);
    p.src_tindices = static_cast<const uint32_t*>(ctx->buf[HVX_BUF_TRANSITION_INDICES]);
    p.states = static_cast<hvx_surface_state*>(pub->buf[HVX_PUB_STATES]);
    p.vertices = static_cast<hvx_vertex*>(pub->buf[HVX_PUB_REGULAR_VERTICES]);
    p.indices = static_cast<uint32_t*>(pub->buf[HVX_PUB_REGULAR_INDICES]);
    p.tvertices = static_cast<hvx_vertex*>(pub->buf[HVX_PUB_TRANSITION_VERTICES]);
    p.tindices = static_cast<uint32_t*>(pub->buf[HVX_PUB_TRANSITION_INDICES]);
    p.regular_draws = static_cast<hvx_draw_indexed_indirect*>(pub->buf[HVX_PUB_REGULAR_DRAWS]);
    p.transition_draws = static_cast<hvx_draw_indexed_indirect*>(pub->buf[HVX_PUB_TRANSITION_DRAWS]);
    p.feedback = static_cast<hvx_surface_feedback*>(pub->buf[HVX_PUB_FEEDBACK]);
}

int hvx_publish_surfaces(hvx_publisher* pub, const hvx_surface_job* jobs, const uint32_t* job_chunk,
                         const hvx_page_meta* page_metadata, uint32_t n) {
    if (!pub) return HVX_E_INVALID_ARGUMENT;
    hvx_ctx* ctx = pub->ctx;
    if (n == 0) return HVX_OK;
    if (!jobs || !job_chunk || !page_metadata) return fail(ctx, HVX_E_INVALID_ARGUMENT, "jobs, job_chunk and page_metadata must be non-NULL");
    if (n > pub->slots) return fail(ctx, HVX_E_BATCH_CAPACITY, "batch of %u jobs exceeds the publisher's %u slots", n, pub->slots);
    const hvx_config& c = ctx->cfg;
    std::vector<uint8_t> seen(pub->slots, 0);
    for (uint32_t i = 0; i < n; ++i) {
        const hvx_surface_job& j = jobs[i];
        if (j.slot >= pub->slots) return fail(ctx, HVX_E_INVALID_ARGUMENT, "job %u names slot %u of %u", i, j.slot, pub->slots);
        if (seen[j.slot]) return fail(ctx, HVX_E_INVALID_ARGUMENT, "slot %u appears twice in one publication batch", j.slot);
        seen[j.slot] = 1;
        if (job_chunk[i] >= c.max_chunks) return fail(ctx, HVX_E_BATCH_CAPACITY, "job %u reads chunk %u of %u", i, job_chunk[i], c.max_chunks);
        if (j.regular_max_vertices != c.max_vertices || j.regular_max_indices != c.max_indices ||
            j.transition_max_vertices != c.max_transition_vertices || j.transition_max_indices != c.max_transition_indices)
            return fail(ctx, HVX_E_INVALID_CAPACITY, "job %u was built for bank capacities (%u, %u, %u, %u); the publisher's are (%u, %u, %u, %u)", i,
                        j.regular_max_vertices, j.regular_max_indices, j.transition_max_vertices, j.transition_max_indices,
                        c.max_vertices, c.max_indices, c.max_transition_vertices, c.max_transition_indices);
    }
    DeviceGuard guard(ctx->device);
    PublishParams p{};
    fill_publish_params(pub, p);
    p.n_jobs = n;
    const void *dj = nullptr, *dc = nullptr, *dm = nullptr;
    int rc;
    if ((rc = stage_to(ctx, &pub->stage[0], &pub->stage_bytes[0], jobs, static_cast<uint64_t>(n) * sizeof(hvx_surface_job), &dj))) return rc;
    if ((rc = stage_to(ctx, &pub->stage[1], &pub->stage_bytes[1], job_chunk, static_cast<uint64_t>(n) * 4, &dc))) return rc;
    if ((rc = stage_to(ctx, &pub->stage[2], &pub->stage_bytes[2], page_metadata, static_cast<uint64_t>(pub->slots) * sizeof(hvx_page_meta), &dm))) return rc;
    p.jobs = static_cast<const hvx_surface_job*>(dj);
    p.job_chunk = static_cast<const uint32_t*>(dc);
    p.meta = static_cast<const hvx_page_meta*>(dm);
    cudaError_t e = launch_publish(p, ctx->dev, ctx->stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "launch_publish");
    ctx->launches += 2;
    return HVX_OK;
}

int hvx_refresh_visibility(hvx_publisher* pub, const hvx_draw_page* draw_pages) {
    if (!pub) return HVX_E_INVALID_ARGUMENT;
    hvx_ctx* ctx = pub->ctx;
    if (!draw_pages) return fail(ctx, HVX_E_INVALID_ARGUMENT, "draw_pages is NULL");
    DeviceGuard guard(ctx->device);
    PublishParams p{};
    fill_publish_params(pub, p);
    const void* dp = nullptr;
    int rc;
    if ((rc = stage_to(ctx, &pub->stage[3], &pub->stage_bytes[3], draw_pages, static_cast<uint64_t>(pub->slots) * sizeof(hvx_draw_page), &dp))) return rc;
    p.draw_pages = static_cast<const hvx_draw_page*>(dp);
    cudaError_t e = launch_visibility(p, ctx->dev, ctx->stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "launch_visibility");
    ctx->launches += 1;
    return HVX_OK;
}

void* hvx_publisher_buffer(hvx_publisher* pub, int id) { return (pub && id >= 0 && id < HVX_PUB_COUNT) ? pub->buf[id] : nullptr; }
uint64_t hvx_publisher_buffer_bytes(hvx_publisher* pub, int id) { return (pub && id >= 0 && id < HVX_PUB_COUNT) ? pub->bytes[id] : 0; }

int hvx_publisher_read(hvx_publisher* pub, int id, uint64_t offset, uint64_t bytes, void* dst) {
    if (!pub || id < 0 || id >= HVX_PUB_COUNT || !dst) return HVX_E_INVALID_ARGUMENT;
    hvx_ctx* ctx = pub->ctx;
    if (offset > pub->bytes[id] || bytes > pub->bytes[id] - offset) return fail(ctx, HVX_E_INVALID_ARGUMENT, "read past the end of publisher buffer %d", id);
    DeviceGuard guard(ctx->device);
    HVX_CUDA(ctx, cudaMemcpyAsync(dst, static_cast<const char*>(pub->buf[id]) + offset, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    HVX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return HVX_OK;
}

int hvx_publisher_write(hvx_publisher* pub, int id, uint64_t offset, uint64_t bytes, const void* src) {
    if (!pub || id < 0 || id >= HVX_PUB_COUNT || !src) return HVX_E_INVALID_ARGUMENT;
    hvx_ctx* ctx = pub->ctx;
    if (offset > pub->bytes[id] || bytes > pub->bytes[id] - offset) return fail(ctx, HVX_E_INVALID_ARGUMENT, "write past the end of publisher buffer %d", id);
    DeviceGuard guard(ctx->device);
    HVX_CUDA(ctx, cudaMemcpyAsync(static_cast<char*>(pub->buf[id]) + offset, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    HVX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return HVX_OK;
}

// ---- bounded extraction publisher, device side (PV/src/extraction.rs) --------------------
static void extraction_publisher_release_device(hvx_extraction_publisher* pub) {
    if (!pub->ctx) return;
    hvx_ctx* ctx = pub->ctx;
    DeviceGuard guard(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < HVX_XPUB_COUNT; ++i)
        if (pub->buf[i]) {
            cudaFree(pub->buf[i]);
            ctx->allocated -= pub->bytes[i];
            pub->buf[i] = nullptr;
        }
    if (pub->d_jobs) {
        cudaFree(pub->d_jobs);
        ctx->allocated -= pub->d_jobs_bytes;
        pub->d_jobs = nullptr;
    }
    pub->ctx = nullptr;
}

int hvx_extraction_publisher_attach(hvx_extraction_publisher* pub, hvx_ctx* ctx) {
    static_assert(sizeof(hvx_extraction_request) == 32 && sizeof(hvx_extraction_range) == 32 &&
                      sizeof(hvx_extraction_counters) == 48 && sizeof(CommitJob) == 40,
                  "extraction POD layouts (PV/src/extraction.rs:8-109)");
    if (!pub || !ctx) return HVX_E_INVALID_ARGUMENT;
    if (pub->ctx) return fail(ctx, HVX_E_INVALID_ARGUMENT, "the extraction publisher is already attached");
    hvx_extraction_plan plan;
    const hvx_extraction_limits limits = pub->host.limits();
    if (int rc = hvx_extraction_limits_plan(&limits, &plan)) return fail(ctx, rc, "%s", hvx_status_name(rc));
    DeviceGuard guard(ctx->device);
    pub->ctx = ctx;
    pub->release_device = extraction_publisher_release_device;
    pub->bytes[HVX_XPUB_VERTICES] = plan.vertex_bytes;
    pub->bytes[HVX_XPUB_INDICES] = plan.index_bytes;
    pub->bytes[HVX_XPUB_PAGE_RANGES] = plan.page_range_bytes;
    pub->bytes[HVX_XPUB_COUNTERS] = plan.counter_bytes;
    for (int i = 0; i < HVX_XPUB_COUNT; ++i) {
        cudaError_t e = cudaMalloc(&pub->buf[i], pub->bytes[i]);
        if (e == cudaSuccess && i >= HVX_XPUB_PAGE_RANGES) e = cudaMemsetAsync(pub->buf[i], 0, pub->bytes[i], ctx->stream);
        if (e != cudaSuccess) {
            cudaGetLastError();
            if (pub->buf[i]) cudaFree(pub->buf[i]);  // allocated, but the clear failed
            pub->buf[i] = nullptr;
            const int rc = fail(ctx, HVX_E_DEVICE_LIMIT, "DeviceLimit: extraction arena %d requires %llu bytes: %s", i,
                                static_cast<unsigned long long>(pub->bytes[i]), cudaGetErrorString(e));
            extraction_publisher_release_device(pub);
            return rc;
        }
        ctx->allocated += pub->bytes[i];
    }
    return HVX_OK;
}

int hvx_extraction_commit(hvx_extraction_publisher* pub, const uint32_t* chunk, const uint32_t* page_slot,
                          const hvx_reservation* reservations, uint32_t n) {
    if (!pub || !pub->ctx) return HVX_E_INVALID_ARGUMENT;
    hvx_ctx* ctx = pub->ctx;
    if (n == 0) return HVX_OK;
    if (!chunk || !page_slot || !reservations) return fail(ctx, HVX_E_INVALID_ARGUMENT, "hvx_extraction_commit: NULL argument");
    const hvx_extraction_limits limits = pub->host.limits();
    std::vector<CommitJob> jobs(n);
    for (uint32_t i = 0; i < n; ++i) {
        const hvx_surface_allocation& a = reservations[i].allocation;
        if (chunk[i] >= ctx->cfg.max_chunks) return fail(ctx, HVX_E_BATCH_CAPACITY, "job %u: chunk %u exceeds max_chunks %u", i, chunk[i], ctx->cfg.max_chunks);
        if (page_slot[i] >= limits.max_page_slots)
            return fail(ctx, HVX_E_INVALID_ARGUMENT, "job %u: page slot %u exceeds max_page_slots %u", i, page_slot[i], limits.max_page_slots);
        if (static_cast<uint64_t>(a.vertices.first) + a.vertices.count > limits.max_vertices ||
            static_cast<uint64_t>(a.indices.first) + a.indices.count > limits.max_indices)
            return fail(ctx, HVX_E_INVALID_ARGUMENT, "job %u: reservation lies outside the bounded arenas", i);
        jobs[i].chunk = chunk[i];
        jobs[i].page_slot = page_slot[i];
        jobs[i].range = gpu_range(a, reservations[i].generation);
    }
    DeviceGuard guard(ctx->device);
    const void* dj = nullptr;
    int rc;
    if ((rc = stage_to(ctx, &pub->d_jobs, &pub->d_jobs_bytes, jobs.data(), static_cast<uint64_t>(n) * sizeof(CommitJob), &dj))) return rc;
    HVX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `jobs` is pageable host memory that dies with this call
    CommitParams p{};
    p.n_jobs = n;
    p.src_max_vertices = ctx->cfg.max_vertices;
    p.src_max_indices = ctx->cfg.max_indices;
    p.jobs = static_cast<const CommitJob*>(dj);
    p.regular_counters = static_cast<const hvx_emission_counters*>(ctx->buf[HVX_BUF_REGULAR_COUNTERS]);
    p.src_vertices = static_cast<const hvx_vertex*>(ctx->buf[HVX_BUF_REGULAR_VERTICES]);
    p.src_indices = static_cast<const uint32_t*>(ctx->buf[HVX_BUF_REGULAR_INDICES]);
    p.vertices = static_cast<hvx_vertex*>(pub->buf[HVX_XPUB_VERTICES]);
    p.indices = static_cast<uint32_t*>(pub->buf[HVX_XPUB_INDICES]);
    p.page_ranges = static_cast<hvx_extraction_range*>(pub->buf[HVX_XPUB_PAGE_RANGES]);
    p.counters = static_cast<hvx_extraction_counters*>(pub->buf[HVX_XPUB_COUNTERS]);
    cudaError_t e = launch_commit(p, ctx->dev, ctx->stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "launch_commit");
    ctx->launches += 1;
    return HVX_OK;
}

void* hvx_extraction_publisher_buffer(hvx_extraction_publisher* pub, int id) {
    return (pub && id >= 0 && id < HVX_XPUB_COUNT) ? pub->buf[id] : nullptr;
}

int hvx_extraction_publisher_read(hvx_extraction_publisher* pub, int id, uint64_t offset, uint64_t bytes, void* dst) {
    if (!pub || !pub->ctx || id < 0 || id >= HVX_XPUB_COUNT || !dst) return HVX_E_INVALID_ARGUMENT;
    hvx_ctx* ctx = pub->ctx;
    if (offset > pub->bytes[id] || bytes > pub->bytes[id] - offset) return fail(ctx, HVX_E_INVALID_ARGUMENT, "read past the end of extraction arena %d", id);
    DeviceGuard guard(ctx->device);
    HVX_CUDA(ctx, cudaMemcpyAsync(dst, static_cast<const char*>(pub->buf[id]) + offset, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    HVX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return HVX_OK;
}

// ---- legacy 8^3-brick marching cubes (helio-pass-voxel-mesh) -----------------------------
struct hvx_brick_mesher {
    hvx_ctx* ctx = nullptr;
    uint32_t max_bricks = 0;
    void* buf[HVX_BRICK_BUF_COUNT] = {};
    uint64_t bytes[HVX_BRICK_BUF_COUNT] = {};
    void* stage[3] = {nullptr, nullptr, nullptr};  // meta, voxels, dirty list
    uint64_t stage_bytes[3] = {0, 0, 0};
};

int hvx_brick_mesher_create(hvx_ctx* ctx, uint32_t max_bricks, hvx_brick_mesher** out) {
    static_assert(sizeof(hvx_brick_meta) == 8 && sizeof(hvx_dirty_brick) == 32 && sizeof(hvx_brick_meshlet) == 32,
                  "legacy brick POD layouts");
    if (!ctx || !out) return HVX_E_INVALID_ARGUMENT;
    *out = nullptr;
    if (max_bricks == 0) return fail(ctx, HVX_E_INVALID_CAPACITY, "a brick mesher needs at least one brick slot");
    if (static_cast<uint64_t>(max_bricks) * HVX_BRICK_MAX_ENTRIES > 0x7fffffffull)
        return fail(ctx, HVX_E_DEVICE_LIMIT, "DeviceLimit: %u brick slots exceed 32-bit draw arguments", max_bricks);
    DeviceGuard guard(ctx->device);
    hvx_brick_mesher* m = new (std::nothrow) hvx_brick_mesher;
    if (!m) return fail(ctx, HVX_E_DEVICE_LIMIT, "out of host memory");
    m->ctx = ctx;
    m->max_bricks = max_bricks;
    const uint64_t entries = static_cast<uint64_t>(max_bricks) * HVX_BRICK_MAX_ENTRIES;
    m->bytes[HVX_BRICK_VERTICES] = entries * 16ull;
    m->bytes[HVX_BRICK_NORMALS] = entries * 16ull;
    m->bytes[HVX_BRICK_INDICES] = entries * 4ull;
    m->bytes[HVX_BRICK_DESCRIPTORS] = max_bricks * sizeof(hvx_brick_meshlet);
    m->bytes[HVX_BRICK_DRAWS] = max_bricks * sizeof(hvx_draw_indexed_indirect);
    m->bytes[HVX_BRICK_REJECTED] = sizeof(uint32_t);
    for (int i = 0; i < HVX_BRICK_BUF_COUNT; ++i) {
        cudaError_t e = cudaMalloc(&m->buf[i], m->bytes[i]);
        if (e == cudaSuccess && i >= HVX_BRICK_DESCRIPTORS) e = cudaMemsetAsync(m->buf[i], 0, m->bytes[i], ctx->stream);
        if (e != cudaSuccess) {
            cudaGetLastError();
            if (m->buf[i]) cudaFree(m->buf[i]);  // allocated, but the clear failed
            m->buf[i] = nullptr;
            const int rc = fail(ctx, HVX_E_DEVICE_LIMIT, "DeviceLimit: brick buffer %d requires %llu bytes: %s", i,
                                static_cast<unsigned long long>(m->bytes[i]), cudaGetErrorString(e));
            hvx_brick_mesher_destroy(m);
            return rc;
        }
        ctx->allocated += m->bytes[i];
    }
    *out = m;
    return HVX_OK;
}

void hvx_brick_mesher_destroy(hvx_brick_mesher* m) {
    if (!m) return;
    DeviceGuard guard(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    for (int i = 0; i < HVX_BRICK_BUF_COUNT; ++i)
        if (m->buf[i]) {
            cudaFree(m->buf[i]);
            m->ctx->allocated -= m->bytes[i];
        }
    for (int i = 0; i < 3; ++i)
        if (m->stage[i]) {
            cudaFree(m->stage[i]);
            m->ctx->allocated -= m->stage_bytes[i];
        }
    delete m;
}

int hvx_brick_extract(hvx_brick_mesher* m, const hvx_brick_meta* meta, uint32_t n_meta, const uint32_t* voxels,
                      uint64_t n_words, const hvx_dirty_brick* dirty, uint32_t n_dirty) {
    if (!m) return HVX_E_INVALID_ARGUMENT;
    hvx_ctx* ctx = m->ctx;
    if (n_dirty == 0) return HVX_OK;
    if (!meta || !voxels || !dirty) return fail(ctx, HVX_E_INVALID_ARGUMENT, "hvx_brick_extract: NULL argument");
    // Host-resident lists are validated here, before anything touches the GPU.  Device-resident lists (the
    // steady-state path: metadata and dirty list produced on the device) are bounds-checked by the kernel, which
    // skips an offending entry and counts it in HVX_BRICK_REJECTED.
    auto on_device = [](const void* ptr) {
        cudaPointerAttributes attr{};
        if (cudaPointerGetAttributes(&attr, ptr) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
    };
    const bool host_lists = !on_device(meta) && !on_device(dirty);
    if (!host_lists && (!on_device(meta) || !on_device(dirty)))
        return fail(ctx, HVX_E_INVALID_ARGUMENT, "meta and dirty must both be host arrays or both be device arrays");
    for (uint32_t i = 0; host_lists && i < n_dirty; ++i) {
        const uint32_t slot = dirty[i].brick_slot;
        if (slot >= m->max_bricks || slot >= n_meta)
            return fail(ctx, HVX_E_BATCH_CAPACITY, "dirty brick %u: slot %u exceeds capacity %u", i, slot,
                        m->max_bricks < n_meta ? m->max_bricks : n_meta);
        if (static_cast<uint64_t>(meta[slot].data_offset) + HVX_BRICK_VOXEL_WORDS > n_words)
            return fail(ctx, HVX_E_SAMPLE_COUNT, "SampleCount: brick slot %u needs words [%u, %u) of %llu", slot,
                        meta[slot].data_offset, meta[slot].data_offset + HVX_BRICK_VOXEL_WORDS,
                        static_cast<unsigned long long>(n_words));
    }
    DeviceGuard guard(ctx->device);
    const void *dm = nullptr, *dv = nullptr, *dd = nullptr;
    int rc;
    if ((rc = stage_to(ctx, &m->stage[0], &m->stage_bytes[0], meta, static_cast<uint64_t>(n_meta) * sizeof(hvx_brick_meta), &dm))) return rc;
    if ((rc = stage_to(ctx, &m->stage[1], &m->stage_bytes[1], voxels, n_words * 4ull, &dv))) return rc;
    if ((rc = stage_to(ctx, &m->stage[2], &m->stage_bytes[2], dirty, static_cast<uint64_t>(n_dirty) * sizeof(hvx_dirty_brick), &dd))) return rc;
    BrickParams p{};
    p.n_dirty = n_dirty;
    p.slot_limit = m->max_bricks < n_meta ? m->max_bricks : n_meta;
    p.n_words = n_words;
    p.rejected = static_cast<uint32_t*>(m->buf[HVX_BRICK_REJECTED]);
    p.meta = static_cast<const hvx_brick_meta*>(dm);
    p.voxels = static_cast<const uint32_t*>(dv);
    p.dirty = static_cast<const hvx_dirty_brick*>(dd);
    p.vertices = static_cast<float4*>(m->buf[HVX_BRICK_VERTICES]);
    p.normals = static_cast<float4*>(m->buf[HVX_BRICK_NORMALS]);
    p.indices = static_cast<uint32_t*>(m->buf[HVX_BRICK_INDICES]);
    p.descriptors = static_cast<hvx_brick_meshlet*>(m->buf[HVX_BRICK_DESCRIPTORS]);
    p.draws = static_cast<hvx_draw_indexed_indirect*>(m->buf[HVX_BRICK_DRAWS]);
    cudaError_t e = launch_bricks(p, ctx->dev, ctx->stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "launch_bricks");
    ctx->launches += 1;
    // caller-owned pageable arrays: the staged copies must have left them before we return
    if (host_lists) HVX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return HVX_OK;
}

int hvx_brick_clear_slot(hvx_brick_mesher* m, uint32_t brick_slot) {
    if (!m) return HVX_E_INVALID_ARGUMENT;
    hvx_ctx* ctx = m->ctx;
    if (brick_slot >= m->max_bricks) return fail(ctx, HVX_E_BATCH_CAPACITY, "brick slot %u exceeds capacity %u", brick_slot, m->max_bricks);
    DeviceGuard guard(ctx->device);
    HVX_CUDA(ctx, cudaMemsetAsync(static_cast<char*>(m->buf[HVX_BRICK_DRAWS]) + static_cast<size_t>(brick_slot) * sizeof(hvx_draw_indexed_indirect),
                                  0, sizeof(hvx_draw_indexed_indirect), ctx->stream));
    return HVX_OK;
}

void* hvx_brick_buffer(hvx_brick_mesher* m, int id) { return (m && id >= 0 && id < HVX_BRICK_BUF_COUNT) ? m->buf[id] : nullptr; }
uint64_t hvx_brick_buffer_bytes(hvx_brick_mesher* m, int id) { return (m && id >= 0 && id < HVX_BRICK_BUF_COUNT) ? m->bytes[id] : 0; }

int hvx_brick_read(hvx_brick_mesher* m, int id, uint64_t offset, uint64_t bytes, void* dst) {
    if (!m || id < 0 || id >= HVX_BRICK_BUF_COUNT || !dst) return HVX_E_INVALID_ARGUMENT;
    hvx_ctx* ctx = m->ctx;
    if (offset > m->bytes[id] || bytes > m->bytes[id] - offset) return fail(ctx, HVX_E_INVALID_ARGUMENT, "read past the end of brick buffer %d", id);
    DeviceGuard guard(ctx->device);
    HVX_CUDA(ctx, cudaMemcpyAsync(dst, static_cast<const char*>(m->buf[id]) + offset, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    HVX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return HVX_OK;
}

int hvx_build_meshlets(hvx_ctx* ctx, int kind, uint32_t n) {
    if (!ctx) return HVX_E_INVALID_ARGUMENT;
    if (kind != 0 && kind != 1) return fail(ctx, HVX_E_INVALID_ARGUMENT, "kind must be 0 (regular) or 1 (transition)");
    if (n > ctx->cfg.max_chunks) return fail(ctx, HVX_E_BATCH_CAPACITY, "batch of %u chunks exceeds max_chunks %u", n, ctx->cfg.max_chunks);
    if (kind == 1 && ctx->cfg.max_transition_vertices == 0)
        return fail(ctx, HVX_E_INVALID_CAPACITY, "Transvoxel transition capacities must be nonzero (vertices=0, indices=0)");
    if (n == 0) return HVX_OK;
    const uint32_t last = kind ? ctx->n_transition : ctx->n_regular;
    if (n > last)
        return fail(ctx, HVX_E_BATCH_CAPACITY, "meshlets requested for %u chunks, the last %s extraction held %u", n,
                    kind ? "transition" : "regular", last);
    DeviceGuard guard(ctx->device);
    const int mid = kind ? HVX_BUF_TRANSITION_MESHLETS : HVX_BUF_REGULAR_MESHLETS;
    int rc;
    for (int id = mid; id < mid + 3; ++id)
        if ((rc = ensure_buffer(ctx, id))) return rc;
    MeshletParams p{};
    p.n_chunks = n;
    p.transition = static_cast<uint32_t>(kind);
    p.max_vertices = kind ? ctx->cfg.max_transition_vertices : ctx->cfg.max_vertices;
    p.max_indices = kind ? ctx->cfg.max_transition_indices : ctx->cfg.max_indices;
    p.max_meshlets = (p.max_indices + 62u) / 63u;
    p.vertices = static_cast<hvx_vertex*>(ctx->buf[kind ? HVX_BUF_TRANSITION_VERTICES : HVX_BUF_REGULAR_VERTICES]);
    p.indices = static_cast<uint32_t*>(ctx->buf[kind ? HVX_BUF_TRANSITION_INDICES : HVX_BUF_REGULAR_INDICES]);
    p.regular_counters = static_cast<hvx_emission_counters*>(ctx->buf[HVX_BUF_REGULAR_COUNTERS]);
    p.transition_counters = static_cast<hvx_transition_counters*>(ctx->buf[HVX_BUF_TRANSITION_COUNTERS]);
    p.descs = kind ? ctx->d_tdescs : ctx->d_descs;
    p.meshlets = static_cast<hvx_meshlet*>(ctx->buf[mid]);
    p.bounds = static_cast<hvx_meshlet_bounds*>(ctx->buf[mid + 1]);
    p.meshlet_counts = static_cast<uint32_t*>(ctx->buf[mid + 2]);
    cudaError_t e = launch_meshlets(p, ctx->dev, ctx->stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "launch_meshlets");
    ctx->launches += 1;
    return HVX_OK;
}

int hvx_weld_meshes(hvx_ctx* ctx, int kind, uint32_t n) {
    if (!ctx) return HVX_E_INVALID_ARGUMENT;
    if (kind != 0 && kind != 1) return fail(ctx, HVX_E_INVALID_ARGUMENT, "kind must be 0 (regular) or 1 (transition)");
    if (n > ctx->cfg.max_chunks) return fail(ctx, HVX_E_BATCH_CAPACITY, "batch of %u chunks exceeds max_chunks %u", n, ctx->cfg.max_chunks);
    if (kind == 1 && ctx->cfg.max_transition_vertices == 0)
        return fail(ctx, HVX_E_INVALID_CAPACITY, "Transvoxel transition capacities must be nonzero (vertices=0, indices=0)");
    if (n == 0) return HVX_OK;
    const uint32_t last = kind ? ctx->n_transition : ctx->n_regular;
    if (n > last)
        return fail(ctx, HVX_E_BATCH_CAPACITY, "weld requested for %u chunks, the last %s extraction held %u", n,
                    kind ? "transition" : "regular", last);
    DeviceGuard guard(ctx->device);
    WeldParams p{};
    p.n_chunks = n;
    p.max_vertices = kind ? ctx->cfg.max_transition_vertices : ctx->cfg.max_vertices;
    p.max_indices = kind ? ctx->cfg.max_transition_indices : ctx->cfg.max_indices;
    p.table_words = 64;
    while (p.table_words < 2ull * p.max_vertices) p.table_words <<= 1;
    p.scratch_words_per_cta = p.table_words + 2u * p.max_vertices;
    // four 512-thread CTAs per SM hide the phase barriers of small meshes behind each other; fewer when their scratch
    // (table + two maps per CTA) would pass 1 GiB
    uint32_t per_sm = 4;
    while (per_sm > 1 && static_cast<uint64_t>(ctx->dev.sm_count) * per_sm * p.scratch_words_per_cta * sizeof(uint32_t) > (1ull << 30)) per_sm /= 2;
    const uint32_t ctas = std::min<uint32_t>(n, static_cast<uint32_t>(ctx->dev.sm_count) * per_sm);
    const uint64_t need = static_cast<uint64_t>(ctx->dev.sm_count) * per_sm * p.scratch_words_per_cta * sizeof(uint32_t);
    if (need > ctx->weld_bytes) {  // grow-only scratch, sized for a full machine of CTAs at this kind's capacity
        if (ctx->d_weld) {
            HVX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            cudaFree(ctx->d_weld);
            ctx->allocated -= ctx->weld_bytes;
            ctx->d_weld = nullptr;
            ctx->weld_bytes = 0;
        }
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&ctx->d_weld), need);
        if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc");
        ctx->weld_bytes = need;
        ctx->allocated += need;
    }
    p.vertices = static_cast<hvx_vertex*>(ctx->buf[kind ? HVX_BUF_TRANSITION_VERTICES : HVX_BUF_REGULAR_VERTICES]);
    p.indices = static_cast<uint32_t*>(ctx->buf[kind ? HVX_BUF_TRANSITION_INDICES : HVX_BUF_REGULAR_INDICES]);
    p.ranges = static_cast<hvx_range*>(ctx->buf[kind ? HVX_BUF_TRANSITION_RANGES : HVX_BUF_REGULAR_RANGES]);
    p.counters = ctx->buf[kind ? HVX_BUF_TRANSITION_COUNTERS : HVX_BUF_REGULAR_COUNTERS];
    p.counter_stride = kind ? sizeof(hvx_transition_counters) : sizeof(hvx_emission_counters);
    p.emitted_vertices_offset = kind ? offsetof(hvx_transition_counters, emitted_vertices) : offsetof(hvx_emission_counters, emitted_vertices);
    p.scratch = ctx->d_weld;
    p.work_counter = ctx->d_work + 4;
    cudaError_t e = launch_weld(p, ctx->dev, ctas, ctx->stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "launch_weld");
    ctx->launches += 1;
    return HVX_OK;
}

int hvx_copy_segments(int device, void* cuda_stream, const uint32_t* d_src, uint32_t* d_dst, const uint64_t* segments, uint32_t n) {
    if (n == 0) return HVX_OK;
    if (!d_src || !d_dst || !segments) return fail(nullptr, HVX_E_INVALID_ARGUMENT, "hvx_copy_segments: NULL argument");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(nullptr, HVX_E_INVALID_ARGUMENT, "device %d cannot be selected", device);
    cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
    uint64_t* d_seg = nullptr;
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&d_seg), 24ull * n, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_seg, segments, 24ull * n, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = launch_copy_segments(d_src, d_dst, d_seg, n, stream);
    if (d_seg) cudaFreeAsync(d_seg, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);  // `segments` is the caller's host memory
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "hvx_copy_segments");
    return HVX_OK;
}

void* hvx_buffer(hvx_ctx* ctx, int id) {
    if (!ctx || id < 0 || id >= HVX_BUF_COUNT) return nullptr;
    DeviceGuard guard(ctx->device);
    if (!ctx->buf[id] && ensure_buffer(ctx, id) != HVX_OK) return nullptr;
    return ctx->buf[id];
}

uint64_t hvx_buffer_bytes(hvx_ctx* ctx, int id) {
    if (!ctx || id < 0 || id >= HVX_BUF_COUNT) return 0;
    return arena_bytes(ctx->cfg, id);
}

int hvx_read(hvx_ctx* ctx, int id, uint64_t offset, uint64_t bytes, void* dst) {
    if (!ctx) return HVX_E_INVALID_ARGUMENT;
    if (id < 0 || id >= HVX_BUF_COUNT || !ctx->buf[id]) return fail(ctx, HVX_E_INVALID_ARGUMENT, "buffer %d is not allocated", id);
    if (offset > ctx->buf_bytes[id] || bytes > ctx->buf_bytes[id] - offset) return fail(ctx, HVX_E_INVALID_ARGUMENT, "read past the end of buffer %d", id);
    if (bytes == 0) return HVX_OK;
    DeviceGuard guard(ctx->device);
    HVX_CUDA(ctx, cudaMemcpyAsync(dst, static_cast<const char*>(ctx->buf[id]) + offset, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    HVX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return HVX_OK;
}

int hvx_write(hvx_ctx* ctx, int id, uint64_t offset, uint64_t bytes, const void* src) {
    if (!ctx) return HVX_E_INVALID_ARGUMENT;
    DeviceGuard guard(ctx->device);
    int rc = ensure_buffer(ctx, id);
    if (rc) return rc;
    if (offset > ctx->buf_bytes[id] || bytes > ctx->buf_bytes[id] - offset) return fail(ctx, HVX_E_INVALID_ARGUMENT, "write past the end of buffer %d", id);
    if (bytes == 0) return HVX_OK;
    HVX_CUDA(ctx, cudaMemcpyAsync(static_cast<char*>(ctx->buf[id]) + offset, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    HVX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return HVX_OK;
}

int hvx_read_meshes(hvx_ctx* ctx, int kind, uint32_t first, uint32_t n, hvx_vertex* vertices_out, uint64_t vertex_cap,
                    uint32_t* indices_out, uint64_t index_cap, hvx_range* ranges_out, uint64_t* total_vertices,
                    uint64_t* total_indices) {
    if (!ctx || !ranges_out) return HVX_E_INVALID_ARGUMENT;
    if (kind != 0 && kind != 1) return fail(ctx, HVX_E_INVALID_ARGUMENT, "kind must be 0 (regular) or 1 (transition)");
    if (static_cast<uint64_t>(first) + n > ctx->cfg.max_chunks) return fail(ctx, HVX_E_BATCH_CAPACITY, "chunk range out of bounds");
    const int vid = kind ? HVX_BUF_TRANSITION_VERTICES : HVX_BUF_REGULAR_VERTICES;
    const int iid = kind ? HVX_BUF_TRANSITION_INDICES : HVX_BUF_REGULAR_INDICES;
    const int rid = kind ? HVX_BUF_TRANSITION_RANGES : HVX_BUF_REGULAR_RANGES;
    if (!ctx->buf[vid]) return fail(ctx, HVX_E_INVALID_ARGUMENT, "this configuration has no %s arenas", kind ? "transition" : "regular");
    if (total_vertices) *total_vertices = 0;
    if (total_indices) *total_indices = 0;
    if (n == 0) return HVX_OK;
    DeviceGuard guard(ctx->device);
    std::vector<hvx_range> slot(n);
    HVX_CUDA(ctx, cudaMemcpyAsync(slot.data(), static_cast<hvx_range*>(ctx->buf[rid]) + first, n * sizeof(hvx_range),
                                  cudaMemcpyDeviceToHost, ctx->stream));
    HVX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    uint64_t tv = 0, ti = 0;
    for (uint32_t i = 0; i < n; ++i) {
        ranges_out[i].first_vertex = static_cast<uint32_t>(tv);
        ranges_out[i].vertex_count = slot[i].vertex_count;
        ranges_out[i].first_index = static_cast<uint32_t>(ti);
        ranges_out[i].index_count = slot[i].index_count;
        tv += slot[i].vertex_count;
        ti += slot[i].index_count;
    }
    if (total_vertices) *total_vertices = tv;
    if (total_indices) *total_indices = ti;
    if (tv > 0xffffffffull || ti > 0xffffffffull) return fail(ctx, HVX_E_DEVICE_LIMIT, "DeviceLimit: packed mesh exceeds 32-bit ranges");
    if (tv > vertex_cap || ti > index_cap || (tv && !vertices_out) || (ti && !indices_out))
        return fail(ctx, HVX_E_INVALID_CAPACITY, "output capacity too small: need %llu vertices / %llu indices",
                    static_cast<unsigned long long>(tv), static_cast<unsigned long long>(ti));
    if (tv == 0 && ti == 0) return HVX_OK;
    auto grow = [&](void** ptr, uint64_t* have, uint64_t need) -> int {
        if (*have >= need) return HVX_OK;
        if (*ptr) {
            cudaFree(*ptr);
            ctx->allocated -= *have;
            *ptr = nullptr;
            *have = 0;
        }
        const uint64_t bytes = need + need / 4 + 4096;
        cudaError_t e = cudaMalloc(ptr, bytes);
        if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc(pack staging)");
        *have = bytes;
        ctx->allocated += bytes;
        return HVX_OK;
    };
    int rc;
    if ((rc = grow(&ctx->pack_v, &ctx->pack_v_bytes, tv * sizeof(hvx_vertex)))) return rc;
    if ((rc = grow(&ctx->pack_i, &ctx->pack_i_bytes, ti * 4))) return rc;
    HVX_CUDA(ctx, cudaMemcpyAsync(ctx->d_packed, ranges_out, n * sizeof(hvx_range), cudaMemcpyHostToDevice, ctx->stream));
    cudaError_t e = launch_pack(static_cast<hvx_vertex*>(ctx->buf[vid]), static_cast<uint32_t*>(ctx->buf[iid]),
                                static_cast<hvx_range*>(ctx->buf[rid]) + first, ctx->d_packed, n,
                                static_cast<hvx_vertex*>(ctx->pack_v), static_cast<uint32_t*>(ctx->pack_i), ctx->dev,
                                ctx->stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "launch_pack");
    ctx->launches += 1;
    if (tv) HVX_CUDA(ctx, cudaMemcpyAsync(vertices_out, ctx->pack_v, tv * sizeof(hvx_vertex), cudaMemcpyDeviceToHost, ctx->stream));
    if (ti) HVX_CUDA(ctx, cudaMemcpyAsync(indices_out, ctx->pack_i, ti * 4, cudaMemcpyDeviceToHost, ctx->stream));
    HVX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return HVX_OK;
}

}  // extern "C"
