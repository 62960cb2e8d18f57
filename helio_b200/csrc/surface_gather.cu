// surface_gather.cu -- halo block + transition slabs from the resident page atlas.
//
// Replaces gather_regular / gather_transition / finalize_gather
// (PV/src/surface_gather.wgsl:200-264; host side PV/src/surface_sampling.rs:198-337) for a batch of
// jobs.  The reference looks the page up in the open-addressed table once per gathered SAMPLE
// (120,106 hash walks per page with all six faces); a sample's page is a pure function of its
// coordinates, so here a CTA resolves the at most 4x4x4 candidate pages of its part once, into
// shared memory, with the reference's hash and probe sequence, and every sample then costs one
// atlas read and one store.  The counters keep the reference's per-sample meaning: table_probes
// adds the probe count of a sample's page for every sample, page_misses counts samples.
//
// grid = (7, jobs): part 0 is the (32+2)^3 block, parts 1..6 the six 67x67x3 fine-side slabs.
#include "hvx_device.cuh"
#include "hvx_kernels.h"

namespace hvx {

namespace {

constexpr int PAGE_EDGE = 32;
constexpr int REGULAR_EDGE = 34, REGULAR_COUNT = 34 * 34 * 34;
constexpr int SLAB_EDGE = 67, FACE_STRIDE = 67 * 67 * 3;
constexpr uint32_t AIR = 0x00007fffu;

// PV/src/surface_gather.wgsl:80-100 (the transition face bases, integer form)
__constant__ int c_origin[6][3] = {{0, 0, 1}, {1, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 1, 0}, {0, 0, 1}};
__constant__ int c_u[6][3] = {{0, 1, 0}, {0, 1, 0}, {0, 0, 1}, {0, 0, 1}, {1, 0, 0}, {1, 0, 0}};
__constant__ int c_v[6][3] = {{0, 0, -1}, {0, 0, 1}, {-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}};
__constant__ int c_out[6][3] = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}};

// PV/src/table.rs:163-167 / surface_gather.wgsl:102-105
__device__ __forceinline__ uint32_t mix_hash(uint32_t hash, uint32_t value) {
    const uint32_t mixed = (hash ^ value) * 0x045d9f3bu;
    return mixed ^ (mixed >> 16);
}

struct Lookup {
    uint32_t slot, generation_low, generation_high, probes, found;
};

// lookup_page (surface_gather.wgsl:125-149): linear probing from hash & mask, at most max_probe
// entries, an EMPTY entry ends the walk, tombstones are walked over.
__device__ Lookup lookup_page(const hvx_page_table_entry* __restrict__ table, const hvx_residency& res,
                              const uint32_t (&planet)[4], const int32_t (&rel)[3], uint32_t lod) {
    uint32_t hash = 0x811c9dc5u;
#pragma unroll
    for (int i = 0; i < 4; ++i) hash = mix_hash(hash, planet[i]);
#pragma unroll
    for (int i = 0; i < 3; ++i) hash = mix_hash(hash, static_cast<uint32_t>(rel[i]));
    hash = mix_hash(hash, lod);
    const uint32_t start = hash & res.table_mask;
    uint32_t probe = 0;
    for (; probe < res.max_probe; ++probe) {
        const hvx_page_table_entry e = table[(start + probe) & res.table_mask];
        if (e.state == 0u) return {0u, 0u, 0u, probe + 1u, 0u};
        if (e.state == 1u && e.planet_id[0] == planet[0] && e.planet_id[1] == planet[1] && e.planet_id[2] == planet[2] &&
            e.planet_id[3] == planet[3] && e.relative_lod0_cell_min[0] == rel[0] && e.relative_lod0_cell_min[1] == rel[1] &&
            e.relative_lod0_cell_min[2] == rel[2] && e.lod == lod)
            return {e.slot, e.generation_low, e.generation_high, probe + 1u, 1u};
    }
    return {0u, 0u, 0u, probe, 0u};
}

__device__ __forceinline__ bool epoch_matches(const hvx_residency& res, const hvx_gather_job& job) {
    return res.publication_epoch_low == job.residency_epoch_low && res.publication_epoch_high == job.residency_epoch_high;
}

// floor(c / 32) and c mod 32 for a cell coordinate in units of the gathered LOD
__device__ __forceinline__ int page_of(int c) { return c >> 5; }
__device__ __forceinline__ int local_of(int c) { return c & 31; }

__global__ void __launch_bounds__(256) gather_kernel(const GatherParams p) {
    __shared__ uint32_t pg_origin[64];  // atlas word index of the page's (0,0,0) texel, or ~0 when missing
    __shared__ uint32_t pg_probes[64];
    __shared__ uint32_t red[3][8];
    const uint32_t jobi = p.job_base + blockIdx.y;
    const int part = blockIdx.x;
    const hvx_gather_job job = p.jobs[jobi];
    const hvx_residency res = p.residency;
    if (!epoch_matches(res, job)) return;
    const bool regular = part == 0;
    const int face = part - 1;
    if (!regular && (job.lod == 0u || ((job.transition_mask >> face) & 1u) == 0u)) return;
    const uint32_t lod = regular ? job.lod : job.lod - 1u;
    const uint32_t scale = 1u << lod, span = PAGE_EDGE * scale;
    const uint32_t row_words = res.atlas_tiles_x * PAGE_EDGE, slice_words = row_words * res.atlas_tiles_y * PAGE_EDGE;

    // ---- the 4x4x4 candidate pages around the target page minimum (page indices -1 .. 2 per axis) ----
    if (threadIdx.x < 64) {
        const int ix = static_cast<int>(threadIdx.x & 3) - 1, iy = static_cast<int>((threadIdx.x >> 2) & 3) - 1,
                  iz = static_cast<int>(threadIdx.x >> 4) - 1;
        // i32 wrapping arithmetic like the shader's
        const int32_t rel[3] = {static_cast<int32_t>(static_cast<uint32_t>(job.relative_lod0_cell_min[0]) + static_cast<uint32_t>(ix) * span),
                                static_cast<int32_t>(static_cast<uint32_t>(job.relative_lod0_cell_min[1]) + static_cast<uint32_t>(iy) * span),
                                static_cast<int32_t>(static_cast<uint32_t>(job.relative_lod0_cell_min[2]) + static_cast<uint32_t>(iz) * span)};
        const Lookup l = lookup_page(p.table, res, job.planet_id, rel, lod);
        uint32_t origin = 0xffffffffu;
        if (l.found) {
            const uint32_t tx = l.slot % res.atlas_tiles_x, ty = (l.slot / res.atlas_tiles_x) % res.atlas_tiles_y,
                           tz = l.slot / (res.atlas_tiles_x * res.atlas_tiles_y);
            origin = tx * PAGE_EDGE + ty * PAGE_EDGE * row_words + tz * PAGE_EDGE * slice_words;
        }
        pg_origin[threadIdx.x] = origin;
        pg_probes[threadIdx.x] = l.probes;
    }
    __syncthreads();

    uint32_t n_samples = 0, n_probes = 0, n_misses = 0;
    if (regular) {
        uint32_t* out = p.samples + static_cast<size_t>(jobi) * REGULAR_COUNT;
        for (int linear = threadIdx.x; linear < REGULAR_COUNT; linear += blockDim.x) {
            const int x = linear % REGULAR_EDGE, y = (linear / REGULAR_EDGE) % REGULAR_EDGE, z = linear / (REGULAR_EDGE * REGULAR_EDGE);
            const int cx = x - 1, cy = y - 1, cz = z - 1;  // cell coordinates relative to the page minimum
            const int e = (page_of(cx) + 1) + 4 * (page_of(cy) + 1) + 16 * (page_of(cz) + 1);
            const uint32_t origin = pg_origin[e];
            uint32_t w = AIR;
            if (origin != 0xffffffffu) w = __ldg(p.atlas + origin + local_of(cx) + local_of(cy) * row_words + local_of(cz) * slice_words);
            else ++n_misses;
            out[linear] = w;
            n_probes += pg_probes[e];
            ++n_samples;
        }
    } else {
        uint32_t* out = p.slabs + static_cast<size_t>(jobi) * (6 * FACE_STRIDE) + static_cast<size_t>(face) * FACE_STRIDE;
        for (int linear = threadIdx.x; linear < FACE_STRIDE; linear += blockDim.x) {
            const int layer = linear / (SLAB_EDGE * SLAB_EDGE), rest = linear % (SLAB_EDGE * SLAB_EDGE);
            const int v = rest / SLAB_EDGE, u = rest % SLAB_EDGE;
            int c[3];  // fine-LOD cell coordinates relative to the coarse page minimum
#pragma unroll
            for (int a = 0; a < 3; ++a)
                c[a] = c_origin[face][a] * (2 * PAGE_EDGE) + c_u[face][a] * (u - 1) + c_v[face][a] * (v - 1) + c_out[face][a] * (layer - 1);
            const int e = (page_of(c[0]) + 1) + 4 * (page_of(c[1]) + 1) + 16 * (page_of(c[2]) + 1);
            const uint32_t origin = pg_origin[e];
            uint32_t w = AIR;
            if (origin != 0xffffffffu) w = __ldg(p.atlas + origin + local_of(c[0]) + local_of(c[1]) * row_words + local_of(c[2]) * slice_words);
            else ++n_misses;
            out[linear] = w;
            n_probes += pg_probes[e];
            ++n_samples;
        }
    }
    // ---- counters: one atomic per CTA and counter ------------------------------------------------
    uint32_t vals[3] = {n_samples, n_probes, n_misses};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        uint32_t v = vals[k];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        uint32_t v = 0;
        for (int w = 0; w < 8; ++w) v += red[threadIdx.x][w];
        hvx_gather_counters* c = p.counters + jobi;
        if (threadIdx.x == 0) atomicAdd(regular ? &c->regular_samples : &c->transition_samples, v);
        else if (threadIdx.x == 1) atomicAdd(&c->table_probes, v);
        else if (v != 0u) atomicAdd(&c->page_misses, v);
    }
}

// finalize_gather (surface_gather.wgsl:236-264): one thread per job
__global__ void gather_finalize_kernel(const GatherParams p) {
    const uint32_t jobi = blockIdx.x * blockDim.x + threadIdx.x;
    if (jobi >= p.n_jobs) return;
    const hvx_gather_job job = p.jobs[jobi];
    const hvx_residency res = p.residency;
    hvx_gather_counters* c = p.counters + jobi;
    const Lookup target = lookup_page(p.table, res, job.planet_id, job.relative_lod0_cell_min, job.lod);
    c->table_probes += target.probes;
    const bool current = target.found != 0u && target.slot == job.target_slot && target.generation_low == job.generation_low &&
                         target.generation_high == job.generation_high;
    if (!current || !epoch_matches(res, job)) {
        c->stale_targets = 1u;
        return;
    }
    if (c->page_misses != 0u || c->regular_samples != static_cast<uint32_t>(REGULAR_COUNT) ||
        c->transition_samples != static_cast<uint32_t>(__popc(job.transition_mask & 0x3fu)) * FACE_STRIDE)
        return;
    c->completed = 1u;
    const uint32_t groups[8] = {512u, 128u, 1u, 512u, 96u, 24u, 1u, 96u};
    uint32_t* ind = p.indirect + static_cast<size_t>(jobi) * 24;
    for (int i = 0; i < 8; ++i) {
        ind[3 * i] = groups[i];
        ind[3 * i + 1] = 1u;
        ind[3 * i + 2] = 1u;
    }
}

}  // namespace

cudaError_t launch_gather(const GatherParams& p, const DeviceInfo&, cudaStream_t stream) {
    if (p.n_jobs == 0) return cudaSuccess;
    for (uint32_t first = 0; first < p.n_jobs; first += 65535u) {  // gridDim.y is limited to 65,535
        GatherParams q = p;
        q.job_base = first;
        gather_kernel<<<dim3(7, min(65535u, p.n_jobs - first)), 256, 0, stream>>>(q);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    gather_finalize_kernel<<<(p.n_jobs + 127) / 128, 128, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace hvx
