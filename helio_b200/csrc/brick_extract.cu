// brick_extract.cu -- legacy 8^3-brick marching-cubes surface extraction for sm_100a.
//
// Replaces crates/passes/3d/helio-pass-voxel-mesh/shaders/voxel_surface_extract.wgsl:123-264 (one 64-thread
// workgroup per dirty brick, slots handed out by two workgroup atomics) for a batch of dirty bricks:
//   * one WARP per brick, eight bricks per CTA, grid-stride over the dirty list;
//   * the padded 9^3-byte block (183 words) is read from HBM once, coalesced, into shared memory;
//   * cells are visited 32 at a time in linear order (x fastest); a warp scan of the per-cell index counts
//     replaces the atomics, so the brick's vertex order is deterministic: cell-linear order, one of the orders
//     the reference's atomics can produce.  The overflow rule is the reference's: a cell whose range would
//     pass 2,048 entries is dropped but still advances the counter, the published counts are clamped;
//   * emission is one lane per OUTPUT entry (owner cell found by a 5-step search of the scanned prefix), so
//     the three output streams (position|material, normal, index) are written as full 512 / 128-byte rows.
// HBM-bound: 732 B read + 36 B per emitted entry + 52 B of descriptor / indirect draw per brick.
// Arithmetic follows oracle/brick_oracle.c operation by operation (separate IEEE multiply / add, 1 / sqrt).
#include "hvx_device.cuh"
#include "hvx_kernels.h"

namespace hvx {

namespace {

#define HVX_TABLE static __device__ const
#include "mc_tables.inc"

constexpr int BRICK_WARPS = 8;
constexpr uint32_t BRICK_WORDS = 183;      // VOXEL_MESH_BRICK_VOXEL_WORDS, helio-pass-voxel-mesh/src/lib.rs:34
constexpr uint32_t BRICK_MAX_ENTRIES = 2048;  // MAX_SURFACE_{VERTS,INDICES}_PER_BRICK, helio-voxel-core/src/constants.rs:15-16

// edge_vertex (voxel_surface_extract.wgsl:71-87) in half units: 2 bits per axis, x | y<<2 | z<<4
__device__ __forceinline__ uint32_t edge_mid_halves(uint32_t edge) {
    // edges 0..11: (1,0,0) (2,1,0) (1,2,0) (0,1,0) (1,0,2) (2,1,2) (1,2,2) (0,1,2) (0,0,1) (2,0,1) (2,2,1) (0,2,1)
    constexpr uint64_t packed = (uint64_t(0x01) << 0) | (uint64_t(0x06) << 6) | (uint64_t(0x09) << 12) | (uint64_t(0x04) << 18) |
                                (uint64_t(0x21) << 24) | (uint64_t(0x26) << 30) | (uint64_t(0x29) << 36) | (uint64_t(0x24) << 42) |
                                (uint64_t(0x10) << 48) | (uint64_t(0x12) << 54);
    if (edge == 10u) return 0x1Au;
    if (edge == 11u) return 0x18u;
    return static_cast<uint32_t>(packed >> (6u * edge)) & 0x3fu;
}

__device__ __forceinline__ uint32_t brick_voxel(const uint32_t* __restrict__ words, uint32_t x, uint32_t y, uint32_t z) {
    const uint32_t linear = z * 81u + y * 9u + x;
    return (words[linear >> 2] >> ((linear & 3u) * 8u)) & 0xffu;
}

__device__ __forceinline__ float brick_occupancy(const uint32_t* __restrict__ words, int x, int y, int z) {
    x = min(max(x, 0), 8);
    y = min(max(y, 0), 8);
    z = min(max(z, 0), 8);
    return brick_voxel(words, x, y, z) > 0u ? 1.0f : -1.0f;
}

__global__ void __launch_bounds__(BRICK_WARPS * 32) brick_extract_kernel(const BrickParams p) {
    __shared__ uint32_t s_vox[BRICK_WARPS][BRICK_WORDS + 1];
    __shared__ uint16_t s_pref[BRICK_WARPS][34];
    __shared__ uint32_t s_cell[BRICK_WARPS][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* const vox = s_vox[warp];
    uint16_t* const pref = s_pref[warp];
    uint32_t* const cells = s_cell[warp];

    for (uint32_t b = blockIdx.x * BRICK_WARPS + warp; b < p.n_dirty; b += gridDim.x * BRICK_WARPS) {
        const hvx_dirty_brick brick = p.dirty[b];
        const uint32_t slot = brick.brick_slot;
        // host lists were validated by the API; device-resident ones are checked here (warp-uniform)
        const uint32_t data_offset = slot < p.slot_limit ? p.meta[slot].data_offset : 0xffffffffu;
        if (slot >= p.slot_limit || static_cast<uint64_t>(data_offset) + BRICK_WORDS > p.n_words) {
            if (lane == 0) atomicAdd(p.rejected, 1u);
            continue;
        }
        const float vs = brick.origin_size[3];
        __syncwarp();
        // a brick with no empty or no occupied voxel has no surface cell: most bricks of a volume take this exit
        // after their one coalesced read (bytes 729..731 of the last word are padding)
        bool any_set = false, any_clear = false;
        for (uint32_t i = lane; i < BRICK_WORDS; i += 32u) {
            const uint32_t w = p.voxels[data_offset + i];
            vox[i] = w;
            const uint32_t valid = i == BRICK_WORDS - 1u ? 0x000000ffu : 0xffffffffu;
            any_set |= (w & valid) != 0u;
            any_clear |= (((w | ~valid) - 0x01010101u) & ~(w | ~valid) & 0x80808080u) != 0u;  // some byte is zero
        }
        const bool has_surface = __any_sync(0xffffffffu, any_set) && __any_sync(0xffffffffu, any_clear);
        __syncwarp();
        float4* const out_v = p.vertices + static_cast<size_t>(slot) * BRICK_MAX_ENTRIES;
        float4* const out_n = p.normals + static_cast<size_t>(slot) * BRICK_MAX_ENTRIES;
        uint32_t* const out_i = p.indices + static_cast<size_t>(slot) * BRICK_MAX_ENTRIES;
        uint32_t counter = 0;  // the reference's wg_vertex_count == wg_index_count (both advance together)
        for (uint32_t round = 0; has_surface && round < 16u; ++round) {
            const uint32_t cell = round * 32u + static_cast<uint32_t>(lane);
            const uint32_t cx = cell & 7u, cy = (cell >> 3) & 7u, cz = cell >> 6;
            // corner order of the shader (:153-160): (0,0,0) (1,0,0) (1,1,0) (0,1,0) (0,0,1) (1,0,1) (1,1,1) (0,1,1)
            const uint32_t c[8] = {brick_voxel(vox, cx, cy, cz),         brick_voxel(vox, cx + 1, cy, cz),
                                   brick_voxel(vox, cx + 1, cy + 1, cz), brick_voxel(vox, cx, cy + 1, cz),
                                   brick_voxel(vox, cx, cy, cz + 1),     brick_voxel(vox, cx + 1, cy, cz + 1),
                                   brick_voxel(vox, cx + 1, cy + 1, cz + 1), brick_voxel(vox, cx, cy + 1, cz + 1)};
            uint32_t cube = 0, material = 0;
#pragma unroll
            for (int i = 7; i >= 0; --i)
                if (c[i] != 0u) {
                    cube |= 1u << i;
                    material = c[i];  // ends as the first non-zero corner (:202-208)
                }
            const uint32_t n = HVX_MC_INDEX_COUNT[cube];  // 0 for the empty and the full cube
            uint32_t incl = n;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += up;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            if (total == 0u) continue;  // warp-uniform
            pref[lane] = static_cast<uint16_t>(incl - n);
            if (lane == 31) pref[32] = static_cast<uint16_t>(total);
            cells[lane] = cube | (material << 8) | (n << 16);
            __syncwarp();
            for (uint32_t e = lane; e < total; e += 32u) {
                uint32_t o = 0;  // largest cell with pref <= e: the one whose range holds entry e
#pragma unroll
                for (uint32_t step = 16; step != 0; step >>= 1)
                    if (pref[o + step] <= e) o += step;
                const uint32_t info = cells[o], first = counter + pref[o], i = e - pref[o];
                if (first + (info >> 16) > BRICK_MAX_ENTRIES) continue;  // the whole cell is dropped (:193-195)
                const uint32_t oc = round * 32u + o, ox = oc & 7u, oy = (oc >> 3) & 7u, oz = oc >> 6;
                const uint32_t case_index = info & 0xffu;
                const uint32_t h = edge_mid_halves(HVX_MC_EDGES[case_index][i]);
                const uint32_t hx = h & 3u, hy = (h >> 2) & 3u, hz = h >> 4;
                // world = (cell * vs + origin) + local * vs, local in {0, 0.5, 1}
                float4 v;
                v.x = fadd(fadd(fmul(static_cast<float>(ox), vs), brick.origin_size[0]), fmul(0.5f * static_cast<float>(hx), vs));
                v.y = fadd(fadd(fmul(static_cast<float>(oy), vs), brick.origin_size[1]), fmul(0.5f * static_cast<float>(hy), vs));
                v.z = fadd(fadd(fmul(static_cast<float>(oz), vs), brick.origin_size[2]), fmul(0.5f * static_cast<float>(hz), vs));
                v.w = static_cast<float>((info >> 8) & 0xffu);
                // normal at the nearest corner voxel; round() ties to even, so only a full unit moves (:222-228)
                const int nx = static_cast<int>(ox + (hx >> 1)), ny = static_cast<int>(oy + (hy >> 1)), nz = static_cast<int>(oz + (hz >> 1));
                const float sx = fsub(brick_occupancy(vox, nx + 1, ny, nz), brick_occupancy(vox, nx - 1, ny, nz));
                const float sy = fsub(brick_occupancy(vox, nx, ny + 1, nz), brick_occupancy(vox, nx, ny - 1, nz));
                const float sz = fsub(brick_occupancy(vox, nx, ny, nz + 1), brick_occupancy(vox, nx, ny, nz - 1));
                const float m2 = fadd(fadd(fmul(sx, sx), fmul(sy, sy)), fmul(sz, sz));
                float4 nn = make_float4(0.0f, 1.0f, 0.0f, 0.0f);
                if (m2 >= 0.000001f) {
                    const float inv = fdiv(1.0f, fsqrt(m2 > 0.000001f ? m2 : 0.000001f));
                    nn.x = fmul(sx, inv);
                    nn.y = fmul(sy, inv);
                    nn.z = fmul(sz, inv);
                }
                const uint32_t vi = first + i;
                out_v[vi] = v;
                out_n[vi] = nn;
                out_i[vi] = vi;
            }
            counter += total;
            __syncwarp();
        }
        if (lane == 0) {
            const uint32_t count = min(counter, BRICK_MAX_ENTRIES);
            hvx_brick_meshlet d;
            d.vertex_offset = slot * BRICK_MAX_ENTRIES;
            d.index_offset = slot * BRICK_MAX_ENTRIES;
            d.vertex_count = count;
            d.index_count = count;
            d.brick_index = slot;
            d.volume_id = brick.volume_id;
            d._pad[0] = 0u;
            d._pad[1] = 0u;
            p.descriptors[slot] = d;
            hvx_draw_indexed_indirect draw;
            draw.index_count = count;
            draw.instance_count = count > 0u ? 1u : 0u;
            draw.first_index = slot * BRICK_MAX_ENTRIES;
            draw.base_vertex = static_cast<int32_t>(slot * BRICK_MAX_ENTRIES);
            draw.first_instance = 0u;
            p.draws[slot] = draw;
        }
    }
}

}  // namespace

cudaError_t launch_bricks(const BrickParams& p, const DeviceInfo& dev, cudaStream_t stream) {
    if (p.n_dirty == 0) return cudaSuccess;
    const uint32_t need = (p.n_dirty + BRICK_WARPS - 1) / BRICK_WARPS;
    const uint32_t grid = min(need, static_cast<uint32_t>(dev.sm_count) * 8u);  // 8 CTAs of 8 warps per SM
    brick_extract_kernel<<<grid, BRICK_WARPS * 32, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace hvx
