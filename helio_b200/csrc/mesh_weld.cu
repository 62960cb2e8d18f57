// mesh_weld.cu -- optional vertex-reuse output mode (hvx_weld_meshes).
//
// The reference emits every cell's vertices separately: a crossing edge shared by four cells appears four times in
// the vertex buffer, and the reuse byte of Lengyel's vertex code is never consumed (PV/src/transvoxel_emit.wgsl:322-358
// reads only `code & 0xff`; PV/src/transvoxel.rs:99-101 exposes reuse() and nothing calls it).  The default output of
// this library is that mesh, byte for byte.  This pass turns it, in place, into the indexed mesh with shared vertices
// that edge-ownership reuse would have produced: a vertex is a pure function of its edge (the two corner samples,
// their gradients, the chunk's transition mask), so the copies of one edge are bit-identical 32-byte records, and
// merging bit-identical records IS merging by edge -- without threading edge identities through the extraction kernel.
//   * kept vertex   = the first occurrence of each distinct record, in the original order
//   * index buffer  = the original one, every entry redirected to the kept copy (triangle order and winding unchanged)
//   * records       = range.vertex_count and counters.emitted_vertices become the number of kept vertices
// The oracle is oracle/weld.py (numpy unique + first-occurrence order); the reference has nothing to compare with.
//
// One CTA per chunk (persistent over a self-rearming ticket counter).  Per chunk: (A) every vertex is inserted into an
// open-addressed table of vertex indices in global scratch -- a slot is claimed with atomicCAS, a record equal to the
// slot's resolves to the smaller index with atomicMin; (B) every vertex looks its representative up; (C) an ordered
// compaction moves the kept vertices forward, block by block (a block is loaded into registers before any of its
// stores, and stores only go backwards, so in place is safe); (D) the indices are redirected.

#include "hvx_device.cuh"
#include "hvx_kernels.h"

namespace hvx {
namespace {

constexpr int WELD_NT = 512;
constexpr uint32_t WELD_EMPTY = 0xffffffffu;

__device__ __forceinline__ uint32_t hash_record(const uint4& a, const uint4& b) {
    uint32_t h = 0x811C9DC5u;
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        h = (h ^ w[k]) * 0x01000193u;
        h ^= h >> 15;
    }
    h *= 0x2C1B3C6Du;
    h ^= h >> 12;
    return h;
}

__device__ __forceinline__ bool same_record(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1) {
    return a0.x == b0.x && a0.y == b0.y && a0.z == b0.z && a0.w == b0.w && a1.x == b1.x && a1.y == b1.y && a1.z == b1.z &&
           a1.w == b1.w;
}

__global__ void __launch_bounds__(WELD_NT) weld_kernel(const WeldParams p) {
    __shared__ uint32_t s_chunk;
    __shared__ uint32_t s_sums[WELD_NT / 32], s_prefix[WELD_NT / 32 + 1];
    const uint32_t tid = threadIdx.x;
    uint32_t* const table = p.scratch + static_cast<size_t>(blockIdx.x) * p.scratch_words_per_cta;
    uint32_t* const rep = table + p.table_words;       // [max_vertices] representative (smallest equal index)
    uint32_t* const moved = rep + p.max_vertices;      // [max_vertices] where a kept vertex went

    for (;;) {
        __syncthreads();
        if (tid == 0) s_chunk = atomicAdd(p.work_counter, 1u);
        __syncthreads();
        const uint32_t chunk = s_chunk;
        if (chunk >= p.n_chunks) {
            if (tid == 0) rearm_work_counter(p.work_counter);
            break;
        }
        const hvx_range rg = p.ranges[chunk];
        const uint32_t nv = rg.vertex_count, ni = rg.index_count;
        if (nv == 0) continue;  // an empty chunk, or one that overflowed and published nothing
        uint4* const v = reinterpret_cast<uint4*>(p.vertices + static_cast<size_t>(chunk) * p.max_vertices);
        uint32_t* const idx = p.indices + static_cast<size_t>(chunk) * p.max_indices;
        uint32_t size = 64;
        while (size < 2u * nv) size <<= 1;  // <= table_words (a power of two >= 2 * max_vertices)
        const uint32_t mask = size - 1u;
        for (uint32_t i = tid; i < size; i += WELD_NT) table[i] = WELD_EMPTY;
        __syncthreads();

        // (A) insert: one slot per distinct record, holding the smallest index that carries it
        for (uint32_t i = tid; i < nv; i += WELD_NT) {
            const uint4 a0 = __ldcg(&v[2 * i]), a1 = __ldcg(&v[2 * i + 1]);
            uint32_t s = hash_record(a0, a1) & mask;
            for (;;) {
                const uint32_t cur = atomicCAS(&table[s], WELD_EMPTY, i);
                if (cur == WELD_EMPTY) break;
                const uint4 b0 = __ldcg(&v[2 * cur]), b1 = __ldcg(&v[2 * cur + 1]);
                if (same_record(a0, a1, b0, b1)) {
                    if (i < cur) atomicMin(&table[s], i);
                    break;
                }
                s = (s + 1u) & mask;
            }
        }
        __syncthreads();
        // (B) look up: the probe sequence of a record ends at its slot (slots only ever go from empty to taken)
        for (uint32_t i = tid; i < nv; i += WELD_NT) {
            const uint4 a0 = __ldcg(&v[2 * i]), a1 = __ldcg(&v[2 * i + 1]);
            uint32_t s = hash_record(a0, a1) & mask;
            for (;;) {
                const uint32_t cur = __ldcg(&table[s]);
                const uint4 b0 = __ldcg(&v[2 * cur]), b1 = __ldcg(&v[2 * cur + 1]);
                if (same_record(a0, a1, b0, b1)) {
                    rep[i] = cur;
                    break;
                }
                s = (s + 1u) & mask;
            }
        }
        __syncthreads();
        // (C) ordered compaction of the kept vertices, in place
        uint32_t kept = 0;
        for (uint32_t b0 = 0; b0 < nv; b0 += WELD_NT) {
            const uint32_t i = b0 + tid;
            const bool keep = i < nv && rep[i] == i;
            uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0;
            if (keep) {
                r0 = __ldcg(&v[2 * i]);
                r1 = __ldcg(&v[2 * i + 1]);
            }
            uint32_t total;
            const uint32_t pos = block_exclusive_scan<WELD_NT>(keep ? 1u : 0u, s_sums, s_prefix, total);  // syncs: the block is loaded
            if (keep) {
                const uint32_t d = kept + pos;
                moved[i] = d;
                v[2 * d] = r0;
                v[2 * d + 1] = r1;
            }
            kept += total;
            __syncthreads();
        }
        // (D) indices follow their vertex
        for (uint32_t j = tid; j < ni; j += WELD_NT) idx[j] = moved[rep[idx[j]]];
        if (tid == 0) {
            p.ranges[chunk].vertex_count = kept;
            *reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(p.counters) + static_cast<size_t>(chunk) * p.counter_stride +
                                         p.emitted_vertices_offset) = kept;
        }
    }
}

}  // namespace

cudaError_t launch_weld(const WeldParams& p, const DeviceInfo& dev, uint32_t ctas, cudaStream_t stream) {
    if (p.n_chunks == 0) return cudaSuccess;
    (void)dev;
    weld_kernel<<<ctas, WELD_NT, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace hvx
