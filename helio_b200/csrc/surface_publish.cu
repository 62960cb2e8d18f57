// surface_publish.cu -- generation-safe publication into double-banked per-slot arenas.
//
// Replaces copy_regular_surface / copy_transition_surface / publish_surface / refresh_visibility
// (PV/src/surface_publish.wgsl:125-225) for a batch of jobs.  The reference runs two capacity-sized
// copy dispatches per page (393,216 + 491,520 threads for a 32^3 page whatever the mesh size); here
// the copy grid is sized by what was emitted and moves 16 bytes per thread.  The publish step is one
// thread per job (slots are distinct within a batch), feedback counters are atomics, so the batch is
// equivalent to the reference's one-job-at-a-time submissions in any order.
#include "hvx_device.cuh"
#include "hvx_kernels.h"

namespace hvx {

namespace {

__device__ __forceinline__ bool metadata_is_current(const hvx_page_meta* meta, const hvx_surface_job& job) {
    const hvx_page_meta m = meta[job.slot];
    return m.slot == job.slot && m.generation_low == job.generation_low && m.generation_high == job.generation_high;
}

__device__ __forceinline__ hvx_transition_counters transition_counters_of(const PublishParams& p, uint32_t chunk) {
    if (p.has_transition) return p.transition_counters[chunk];
    hvx_transition_counters t{};  // what the reference's transition extractor reports for mask 0
    t.completed = 1u;
    return t;
}

// grid = (blocks per job, jobs, 2): z = 0 regular, 1 transition
__global__ void __launch_bounds__(256) publish_copy_kernel(const PublishParams p) {
    const uint32_t j = p.job_base + blockIdx.y;
    const bool transition = blockIdx.z != 0;
    if (transition && !p.has_transition) return;
    const hvx_surface_job job = p.jobs[j];
    const uint32_t chunk = p.job_chunk[j];
    uint32_t nv, ni;
    bool ok;
    if (!transition) {
        const hvx_emission_counters c = p.regular_counters[chunk];
        ok = c.completed != 0u && c.vertex_overflow == 0u && c.index_overflow == 0u;
        nv = c.emitted_vertices;
        ni = c.emitted_indices;
    } else {
        const hvx_transition_counters c = p.transition_counters[chunk];
        ok = c.completed != 0u && c.vertex_overflow == 0u && c.index_overflow == 0u;
        nv = c.emitted_vertices;
        ni = c.emitted_indices;
    }
    if (!ok || !metadata_is_current(p.meta, job)) return;
    const uint32_t next_bank = 1u - min(p.states[job.slot].active_bank, 1u);
    const uint32_t bank = job.slot * 2u + next_bank;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    // vertices: 32-byte records moved as two 16-byte halves per thread step
    const uint4* sv = reinterpret_cast<const uint4*>(
        transition ? p.src_tvertices + static_cast<size_t>(chunk) * p.src_max_tvertices
                   : p.src_vertices + static_cast<size_t>(chunk) * p.src_max_vertices);
    uint4* dv = reinterpret_cast<uint4*>(transition ? p.tvertices + static_cast<size_t>(bank) * job.transition_max_vertices
                                                    : p.vertices + static_cast<size_t>(bank) * job.regular_max_vertices);
    for (uint32_t q = tid; q < 2u * nv; q += stride) dv[q] = sv[q];
    const uint32_t* si = transition ? p.src_tindices + static_cast<size_t>(chunk) * p.src_max_tindices
                                    : p.src_indices + static_cast<size_t>(chunk) * p.src_max_indices;
    uint32_t* di = transition ? p.tindices + static_cast<size_t>(bank) * job.transition_max_indices
                              : p.indices + static_cast<size_t>(bank) * job.regular_max_indices;
    for (uint32_t q = tid; q < ni; q += stride) di[q] = si[q];
}

// publish_surface (surface_publish.wgsl:166-216), one thread per job
__global__ void publish_state_kernel(const PublishParams p) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p.n_jobs) return;
    const hvx_surface_job job = p.jobs[j];
    const uint32_t chunk = p.job_chunk[j];
    atomicAdd(&p.feedback->submitted_jobs, 1u);
    if (!metadata_is_current(p.meta, job)) {
        atomicAdd(&p.feedback->stale_rejections, 1u);
        return;
    }
    const hvx_emission_counters rc = p.regular_counters[chunk];
    const hvx_transition_counters tc = transition_counters_of(p, chunk);
    if (rc.completed == 0u || tc.completed == 0u) {
        atomicAdd(&p.feedback->incomplete_rejections, 1u);
        return;
    }
    if (rc.vertex_overflow != 0u || rc.index_overflow != 0u || tc.vertex_overflow != 0u || tc.index_overflow != 0u) {
        atomicAdd(&p.feedback->overflow_rejections, 1u);
        return;
    }
    const hvx_surface_state old_state = p.states[job.slot];
    const uint32_t next_bank = 1u - min(old_state.active_bank, 1u);
    hvx_surface_state s{};
    s.generation_low = job.generation_low;
    s.generation_high = job.generation_high;
    s.active_bank = next_bank;
    s.valid = 1u;
    s.regular_vertex_count = rc.emitted_vertices;
    s.regular_index_count = rc.emitted_indices;
    s.transition_vertex_count = tc.emitted_vertices;
    s.transition_index_count = tc.emitted_indices;
    s.regular_meshlet_count = (rc.emitted_indices + 62u) / 63u;
    s.transition_meshlet_count = (tc.emitted_indices + 62u) / 63u;
    p.states[job.slot] = s;
    const uint32_t bank = job.slot * 2u + next_bank;
    hvx_draw_indexed_indirect d;
    d.index_count = rc.emitted_indices;
    d.instance_count = 0u;
    d.first_index = bank * job.regular_max_indices;
    d.base_vertex = static_cast<int32_t>(bank * job.regular_max_vertices);
    d.first_instance = job.slot;
    p.regular_draws[job.slot] = d;
    d.index_count = tc.emitted_indices;
    d.first_index = bank * job.transition_max_indices;
    d.base_vertex = static_cast<int32_t>(bank * job.transition_max_vertices);
    p.transition_draws[job.slot] = d;
    atomicAdd(&p.feedback->published_jobs, 1u);
}

// refresh_visibility (surface_publish.wgsl:218-225)
__global__ void visibility_kernel(const PublishParams p) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.slots) return;
    const uint32_t visible = (p.states[i].valid != 0u && p.draw_pages[i].visible != 0u) ? 1u : 0u;
    p.regular_draws[i].instance_count = visible;
    p.transition_draws[i].instance_count = visible;
}

// hvx_extraction_commit: move reserved pages' meshes from their fixed-stride extraction slots into the
// bounded arenas at the reserved ranges (PV/src/extraction.rs: the ranges a BoundedExtractionPublisher
// reservation holds), and stamp page_ranges[page_slot] with the generation-tagged GpuExtractionRange.
// grid = (blocks per job, jobs).  A job whose reservation does not match what the chunk emitted is
// skipped by every block and counted once.
__global__ void __launch_bounds__(256) commit_kernel(const CommitParams p) {
    const CommitJob job = p.jobs[p.job_base + blockIdx.y];
    const hvx_emission_counters c = p.regular_counters[job.chunk];
    const bool ok = c.completed != 0u && c.vertex_overflow == 0u && c.index_overflow == 0u &&
                    c.emitted_vertices == job.range.vertex_count && c.emitted_indices == job.range.index_count;
    const bool leader = blockIdx.x == 0 && threadIdx.x == 0;
    if (leader) atomicAdd(&p.counters->requests, 1u);
    if (!ok) {
        if (leader) {
            atomicAdd(&p.counters->overflowed, 1u);
            if (c.vertex_overflow != 0u || c.emitted_vertices != job.range.vertex_count) atomicAdd(&p.counters->vertex_overflow, 1u);
            if (c.index_overflow != 0u || c.emitted_indices != job.range.index_count) atomicAdd(&p.counters->index_overflow, 1u);
        }
        return;
    }
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    const uint4* sv = reinterpret_cast<const uint4*>(p.src_vertices + static_cast<size_t>(job.chunk) * p.src_max_vertices);
    uint4* dv = reinterpret_cast<uint4*>(p.vertices + job.range.first_vertex);
    for (uint32_t q = tid; q < 2u * job.range.vertex_count; q += stride) dv[q] = sv[q];
    // index ranges start anywhere in the arena, so only the source is 16-byte aligned: scalar stores, coalesced
    const uint32_t* si = p.src_indices + static_cast<size_t>(job.chunk) * p.src_max_indices;
    uint32_t* di = p.indices + job.range.first_index;
    for (uint32_t q = tid; q < job.range.index_count; q += stride) di[q] = si[q];
    if (leader) {
        p.page_ranges[job.page_slot] = job.range;
        atomicAdd(&p.counters->vertices, job.range.vertex_count);
        atomicAdd(&p.counters->indices, job.range.index_count);
        atomicAdd(&p.counters->meshlets, job.range.meshlet_count);
        atomicAdd(&p.counters->completed, 1u);
    }
}

}  // namespace

cudaError_t launch_commit(const CommitParams& p, const DeviceInfo& dev, cudaStream_t stream) {
    if (p.n_jobs == 0) return cudaSuccess;
    uint32_t per_job = 8;
    while (per_job > 1 && static_cast<uint64_t>(per_job) * p.n_jobs > 64ull * dev.sm_count) per_job >>= 1;
    for (uint32_t first = 0; first < p.n_jobs; first += 65535u) {  // gridDim.y is limited to 65,535
        CommitParams q = p;
        q.job_base = first;
        commit_kernel<<<dim3(per_job, min(65535u, p.n_jobs - first)), 256, 0, stream>>>(q);
    }
    return cudaGetLastError();
}

cudaError_t launch_publish(const PublishParams& p, const DeviceInfo& dev, cudaStream_t stream) {
    if (p.n_jobs == 0) return cudaSuccess;
    // enough blocks per job to cover a typical mesh in one or two grid-stride steps, bounded by the machine
    uint32_t per_job = 8;
    while (per_job > 1 && static_cast<uint64_t>(per_job) * p.n_jobs * 2 > 64ull * dev.sm_count) per_job >>= 1;
    for (uint32_t first = 0; first < p.n_jobs; first += 65535u) {  // gridDim.y is limited to 65,535
        PublishParams q = p;
        q.job_base = first;
        publish_copy_kernel<<<dim3(per_job, min(65535u, p.n_jobs - first), 2), 256, 0, stream>>>(q);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    publish_state_kernel<<<(p.n_jobs + 127) / 128, 128, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_visibility(const PublishParams& p, const DeviceInfo&, cudaStream_t stream) {
    if (p.slots == 0) return cudaSuccess;
    visibility_kernel<<<(p.slots + 127) / 128, 128, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace hvx
