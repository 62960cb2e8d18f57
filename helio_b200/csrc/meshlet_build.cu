// meshlet_build.cu -- fixed 63-index terrain meshlets with bounding sphere and normal cone
// (the step after extraction in the reference's planetary pass).
//
// Replaces build_regular / build_transition (PV/src/terrain_meshlet_build.wgsl:133-261); the
// arithmetic follows the reference's CPU twin build_terrain_meshlets / compute_bounds
// (PV/src/terrain_meshlet.rs:83-249) operation by operation, so descriptors AND bounds are
// bit-identical to the oracle (the reference only holds its WGSL to 1e-5 / 1e-4).
//
// One thread per meshlet, grid = (ceil(max_meshlets / 64), chunks).  The "bank" of the reference
// is the chunk's fixed-stride slot: first_index = slot * max_indices + 63 m, first_vertex =
// slot * max_vertices, bounds_offset = slot * max_meshlets + m, _pad = 0 regular / 1 transition.
// Chunks whose extraction overflowed (or never completed) publish no meshlets, as in the WGSL.
// The 63 gathered positions live in per-thread local memory (interleaved across lanes, served by
// L1/L2: every vertex of a chunk is referenced by ~1.5 triangles, so this kernel re-reads the mesh
// once and writes 80 bytes per 63 indices).
#include "hvx_device.cuh"
#include "hvx_kernels.h"

namespace hvx {

namespace {

struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 vsub(V3 a, V3 b) { return {fsub(a.x, b.x), fsub(a.y, b.y), fsub(a.z, b.z)}; }
__device__ __forceinline__ float vdot(V3 a, V3 b) { return fadd(fadd(fmul(a.x, b.x), fmul(a.y, b.y)), fmul(a.z, b.z)); }
__device__ __forceinline__ V3 vcross(V3 a, V3 b) {
    return {fsub(fmul(a.y, b.z), fmul(a.z, b.y)), fsub(fmul(a.z, b.x), fmul(a.x, b.z)), fsub(fmul(a.x, b.y), fmul(a.y, b.x))};
}
// normalize (terrain_meshlet.rs:309-312): magnitude > f32::EPSILON, scale by magnitude.recip()
__device__ __forceinline__ bool vnormalize(V3 v, V3& out) {
    const float magnitude = fsqrt(vdot(v, v));
    if (!(magnitude > 1.1920929e-7f)) return false;
    const float inv = fdiv(1.0f, magnitude);
    out = {fmul(v.x, inv), fmul(v.y, inv), fmul(v.z, inv)};
    return true;
}

__global__ void __launch_bounds__(64) meshlet_build_kernel(const MeshletParams p) {
    const uint32_t chunk = p.chunk_base + blockIdx.y;
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t emitted_indices, completed, overflow;
    if (p.transition) {
        const hvx_transition_counters c = p.transition_counters[chunk];
        emitted_indices = c.emitted_indices;
        completed = c.completed;
        overflow = c.vertex_overflow | c.index_overflow;
    } else {
        const hvx_emission_counters c = p.regular_counters[chunk];
        emitted_indices = c.emitted_indices;
        completed = c.completed;
        overflow = c.vertex_overflow | c.index_overflow;
    }
    const uint32_t count = (completed && !overflow) ? (emitted_indices + 62u) / 63u : 0u;
    if (m == 0) p.meshlet_counts[chunk] = count;
    if (m >= count) return;

    const uint32_t n = min(63u, emitted_indices - 63u * m);
    const uint32_t first_index = chunk * p.max_indices + 63u * m;
    const uint32_t first_vertex = chunk * p.max_vertices;
    const uint32_t* idx = p.indices + first_index;
    const hvx_vertex* verts = p.vertices + first_vertex;

    uint32_t ids[63];
    V3 pos[63];
    for (uint32_t i = 0; i < n; ++i) {
        ids[i] = idx[i];
        const float4 q = *reinterpret_cast<const float4*>(&verts[ids[i]]);
        pos[i] = {q.x, q.y, q.z};
    }
    // unique_index_count (terrain_meshlet.rs:171-182)
    uint32_t unique = 0;
    for (uint32_t i = 0; i < n; ++i) {
        bool seen = false;
        for (uint32_t j = 0; j < i; ++j) seen |= ids[j] == ids[i];
        unique += seen ? 0u : 1u;
    }
    // compute_bounds (terrain_meshlet.rs:184-249)
    V3 mn = pos[0], mx = pos[0];
    for (uint32_t i = 0; i < n; ++i) {
        mn = {fminf(mn.x, pos[i].x), fminf(mn.y, pos[i].y), fminf(mn.z, pos[i].z)};
        mx = {fmaxf(mx.x, pos[i].x), fmaxf(mx.y, pos[i].y), fmaxf(mx.z, pos[i].z)};
    }
    const V3 center = {fmul(fadd(mn.x, mx.x), 0.5f), fmul(fadd(mn.y, mx.y), 0.5f), fmul(fadd(mn.z, mx.z), 0.5f)};
    float radius = 0.0f;
    for (uint32_t i = 0; i < n; ++i) {
        const V3 d = vsub(pos[i], center);
        radius = fmaxf(radius, fsqrt(vdot(d, d)));
    }
    V3 apex = center, axis = {0.0f, 0.0f, 0.0f}, stored_axis = {0.0f, 0.0f, 0.0f};
    float cutoff = 1.0f;
    V3 sum = {0.0f, 0.0f, 0.0f};
    for (uint32_t t = 0; t + 2 < n; t += 3) {
        V3 nrm;
        if (vnormalize(vcross(vsub(pos[t + 1], pos[t]), vsub(pos[t + 2], pos[t])), nrm))
            sum = {fadd(sum.x, nrm.x), fadd(sum.y, nrm.y), fadd(sum.z, nrm.z)};
    }
    if (vnormalize(sum, axis)) {
        float min_dot = 1.0f;
        for (uint32_t t = 0; t + 2 < n; t += 3) {
            V3 nrm;
            if (vnormalize(vcross(vsub(pos[t + 1], pos[t]), vsub(pos[t + 2], pos[t])), nrm))
                min_dot = fminf(min_dot, vdot(nrm, axis));
        }
        if (min_dot > 0.1f) {
            float apex_distance = 0.0f;
            for (uint32_t t = 0; t + 2 < n; t += 3) {
                V3 nrm;
                if (!vnormalize(vcross(vsub(pos[t + 1], pos[t]), vsub(pos[t + 2], pos[t])), nrm)) continue;
                const float denominator = vdot(axis, nrm);
                if (denominator > 0.0f)
                    apex_distance = fmaxf(apex_distance, fdiv(vdot(vsub(center, pos[t]), nrm), denominator));
            }
            apex = {fsub(center.x, fmul(axis.x, apex_distance)), fsub(center.y, fmul(axis.y, apex_distance)),
                    fsub(center.z, fmul(axis.z, apex_distance))};
            cutoff = fminf(fadd(fsqrt(fmaxf(fsub(1.0f, fmul(min_dot, min_dot)), 0.0f)), 1.0e-4f), 1.0f);
            stored_axis = axis;
        }
    }
    const uint32_t out = chunk * p.max_meshlets + m;
    const ChunkDesc desc = p.descs[chunk];
    uint4* md = reinterpret_cast<uint4*>(&p.meshlets[out]);
    md[0] = make_uint4(first_index, n, first_vertex, unique);
    md[1] = make_uint4(out, static_cast<uint32_t>(desc.generation), static_cast<uint32_t>(desc.generation >> 32),
                       p.transition ? 1u : 0u);
    float4* bd = reinterpret_cast<float4*>(&p.bounds[out]);
    bd[0] = make_float4(center.x, center.y, center.z, radius);
    bd[1] = make_float4(apex.x, apex.y, apex.z, cutoff);
    bd[2] = make_float4(stored_axis.x, stored_axis.y, stored_axis.z, 0.0f);
}

}  // namespace

cudaError_t launch_meshlets(const MeshletParams& p, const DeviceInfo&, cudaStream_t stream) {
    if (p.n_chunks == 0 || p.max_meshlets == 0) return cudaSuccess;
    // gridDim.y is limited to 65,535: larger batches go in several launches
    for (uint32_t first = 0; first < p.n_chunks; first += 65535u) {
        MeshletParams q = p;
        q.chunk_base = first;
        const dim3 grid((p.max_meshlets + 63u) / 64u, min(65535u, p.n_chunks - first));
        meshlet_build_kernel<<<grid, 64, 0, stream>>>(q);
    }
    return cudaGetLastError();
}

}  // namespace hvx
