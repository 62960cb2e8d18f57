// regular_records.cu -- the per-cell debug records of the regular path, for contexts created with
// HVX_CFG_DEBUG_RECORDS:
//   GpuTransvoxelCell        PV/src/transvoxel_gpu.rs:77-84      (classify_regular_cells, transvoxel_classify.wgsl:75-120)
//   GpuTransvoxelCellOffset  PV/src/transvoxel_emit.rs:14-21     (scan_regular_cells,     transvoxel_emit.wgsl:131-172)
//   GpuTransvoxelScanBlock   PV/src/transvoxel_emit.rs:29-36     (scan_regular_blocks,    transvoxel_emit.wgsl:174-202)
// The extraction kernel (regular_extract.cu) never materialises per-cell words: it classifies from sign bits
// and places vertices with running prefixes.  The records the reference exposes for inspection (and its tests
// compare cell by cell) are therefore produced here, by two small launches that run after the extraction and
// read the same samples: `cell_records_kernel` is the reference's classify + per-block scan (one CTA per
// 256-cell scan block), `block_prefix_kernel` its block scan (one CTA per chunk).  Only cells of dirty
// microbricks are visited; the others keep whatever generation they held (the reference's contract,
// PV/tests/gpu_transvoxel_emission.rs:154-214).  Not on the hot path: nothing here runs unless the caller
// asked for the records.
#include "hvx_device.cuh"
#include "hvx_kernels.h"

namespace hvx {

namespace {

#define HVX_TABLE static __device__ const
#include "transvoxel_tables.inc"

template <int E>
__global__ void __launch_bounds__(256) cell_records_kernel(const RegularParams p) {
    constexpr int S = E + 2, QW = E / 4, NB = E * E * E / 256;
    __shared__ uint16_t case_info[256];
    __shared__ uint64_t sums[8], prefix[9];
    const int tid = threadIdx.x;
    case_info[tid] = HVX_REGULAR_CASE_INFO[tid];
    __syncthreads();
    const uint32_t chunk = blockIdx.x / NB, block = blockIdx.x % NB;
    const ChunkDesc desc = p.descs[chunk];
    const uint32_t lin = block * 256u + static_cast<uint32_t>(tid);
    const int x = lin % E, y = (lin / E) % E, z = lin / (E * E);
    // a chunk flagged uniform was not uploaded: its cells keep their old records, its blocks report nothing
    const bool visited = ((desc.dirty_microbricks >> ((x / QW) + 4 * (y / QW) + 16 * (z / QW))) & 1ull) && !(desc.flags & HVX_CHUNK_UNIFORM);
    uint32_t c = 0, nv = 0, nt = 0, cls = 0;
    if (visited) {
        const uint32_t* s0 = p.samples + static_cast<size_t>(chunk) * S * S * S + (static_cast<size_t>(z + 1) * S + (y + 1)) * S + (x + 1);
        const uint32_t* s1 = s0 + S * S;
        c = (cw_solid(s0[0]) ? 1u : 0u) | (cw_solid(s0[1]) ? 2u : 0u) | (cw_solid(s0[S]) ? 4u : 0u) |
            (cw_solid(s0[S + 1]) ? 8u : 0u) | (cw_solid(s1[0]) ? 16u : 0u) | (cw_solid(s1[1]) ? 32u : 0u) |
            (cw_solid(s1[S]) ? 64u : 0u) | (cw_solid(s1[S + 1]) ? 128u : 0u);
        const uint32_t info = case_info[c];
        nv = info & 15u;
        nt = (info >> 4) & 15u;
        cls = info >> 8;
    }
    uint64_t total;
    const uint64_t before = block_exclusive_scan<256, uint64_t>(nv | (static_cast<uint64_t>(3u * nt) << 32), sums, prefix, total);
    const size_t cell = static_cast<size_t>(chunk) * E * E * E + lin;
    if (visited) {
        const uint32_t glo = static_cast<uint32_t>(desc.generation), ghi = static_cast<uint32_t>(desc.generation >> 32);
        *reinterpret_cast<uint4*>(&p.cells[cell]) = make_uint4(c | (cls << 8) | (nv << 16) | (nt << 24) | 0x80000000u, glo, ghi, 0u);
        *reinterpret_cast<uint4*>(&p.offsets[cell]) =
            make_uint4(static_cast<uint32_t>(before), static_cast<uint32_t>(before >> 32), glo, ghi);
    }
    if (tid == 0) {
        hvx_scan_block blk;
        blk.vertex_count = static_cast<uint32_t>(total);
        blk.index_count = static_cast<uint32_t>(total >> 32);
        blk.first_vertex = 0u;  // block_prefix_kernel
        blk.first_index = 0u;
        p.blocks[static_cast<size_t>(chunk) * NB + block] = blk;
    }
}

template <int E>
__global__ void __launch_bounds__(256) block_prefix_kernel(const RegularParams p) {
    constexpr int NB = E * E * E / 256, K = (NB + 255) / 256;  // scan blocks per chunk, per thread
    __shared__ uint64_t sums[8], prefix[9];
    hvx_scan_block* blocks = p.blocks + static_cast<size_t>(blockIdx.x) * NB;
    const int first = threadIdx.x * K;
    uint64_t mine = 0;
#pragma unroll
    for (int k = 0; k < K; ++k)
        if (first + k < NB) mine += blocks[first + k].vertex_count | (static_cast<uint64_t>(blocks[first + k].index_count) << 32);
    uint64_t total;
    uint64_t run = block_exclusive_scan<256, uint64_t>(mine, sums, prefix, total);
#pragma unroll
    for (int k = 0; k < K; ++k)
        if (first + k < NB) {
            hvx_scan_block& b = blocks[first + k];
            b.first_vertex = static_cast<uint32_t>(run);
            b.first_index = static_cast<uint32_t>(run >> 32);
            run += b.vertex_count | (static_cast<uint64_t>(b.index_count) << 32);
        }
}

// chunks the caller flagged HVX_CHUNK_UNIFORM: nothing was read, the slot reports an empty, completed mesh
__global__ void __launch_bounds__(128) uniform_records_kernel(const RegularParams p, const uint32_t* __restrict__ ids, uint32_t n, uint32_t microbrick_cells) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t chunk = ids[i];
    hvx_emission_counters ec{};
    ec.completed = 1u;
    p.counters[chunk] = ec;
    hvx_classify_counters cc{};
    cc.visited_cells = static_cast<uint32_t>(__popcll(p.descs[chunk].dirty_microbricks)) * microbrick_cells;
    p.classify[chunk] = cc;
    hvx_range rg{};
    rg.first_vertex = (p.chunk_base + chunk) * p.max_vertices;
    rg.first_index = (p.chunk_base + chunk) * p.max_indices;
    p.ranges[chunk] = rg;
}

}  // namespace

cudaError_t launch_uniform_records(int edge, const RegularParams& p, const uint32_t* ids, uint32_t n, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const uint32_t q = static_cast<uint32_t>(edge) / 4u;
    uniform_records_kernel<<<(n + 127u) / 128u, 128, 0, stream>>>(p, ids, n, q * q * q);
    return cudaGetLastError();
}

cudaError_t launch_regular_records(int edge, const RegularParams& p, cudaStream_t stream) {
    if (p.n_chunks == 0 || p.cells == nullptr) return cudaSuccess;
    if (edge == 64) {
        cell_records_kernel<64><<<p.n_chunks * 1024u, 256, 0, stream>>>(p);
        block_prefix_kernel<64><<<p.n_chunks, 256, 0, stream>>>(p);
    } else if (edge == 32) {
        cell_records_kernel<32><<<p.n_chunks * 128u, 256, 0, stream>>>(p);
        block_prefix_kernel<32><<<p.n_chunks, 256, 0, stream>>>(p);
    } else {
        return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace hvx
