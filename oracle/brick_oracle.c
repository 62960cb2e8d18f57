/* brick_oracle.c -- TEST INFRASTRUCTURE: CPU restatement of the reference's legacy 8^3-brick marching-cubes
 * extractor, crates/passes/3d/helio-pass-voxel-mesh/shaders/voxel_surface_extract.wgsl:64-264.
 *
 * The reference allocates vertex / index slots with workgroup atomics, so the ORDER of cells inside a brick's
 * range is whatever the GPU scheduler produced; every order that keeps a thread's own cells ascending is a
 * valid reference result.  This restatement (and the CUDA kernel) uses cell-linear order, x fastest, which is
 * one of them and makes the output deterministic.  parity unpinned: the reference holds no golden vectors or
 * tests for this shader; tests/test_brick_extract.py pins the restatement on hand-derivable cases instead
 * (single voxel, half-filled brick, all 256 cube cases against the table, the overflow rule).
 *
 * Float arithmetic: separate IEEE multiplies and adds in the shader's order (WGSL leaves fusing to the
 * driver); inverseSqrt(x) is restated as 1 / sqrt(x), both correctly rounded.
 * Only tests/ and bench tooling may link this file. */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define HVXO_TABLE static const
#include "mc_tables.inc"

#define MAX_VERTS 2048u   /* helio-voxel-core/src/constants.rs:15-16 */
#define PADDED 9u

/* edge_vertex, voxel_surface_extract.wgsl:71-87 */
static const float EDGE_MID[12][3] = {{0.5f, 0.0f, 0.0f}, {1.0f, 0.5f, 0.0f}, {0.5f, 1.0f, 0.0f}, {0.0f, 0.5f, 0.0f},
                                      {0.5f, 0.0f, 1.0f}, {1.0f, 0.5f, 1.0f}, {0.5f, 1.0f, 1.0f}, {0.0f, 0.5f, 1.0f},
                                      {0.0f, 0.0f, 0.5f}, {1.0f, 0.0f, 0.5f}, {1.0f, 1.0f, 0.5f}, {0.0f, 1.0f, 0.5f}};

/* read_voxel :91-96 */
static uint32_t read_voxel(const uint32_t* words, uint32_t data_offset, uint32_t x, uint32_t y, uint32_t z) {
    const uint32_t linear = z * (PADDED * PADDED) + y * PADDED + x;
    return (words[data_offset + linear / 4u] >> ((linear % 4u) * 8u)) & 0xffu;
}

static float occupancy(const uint32_t* words, uint32_t off, int x, int y, int z) {
    x = x < 0 ? 0 : (x > 8 ? 8 : x);   /* clamped_voxel :105-107 */
    y = y < 0 ? 0 : (y > 8 ? 8 : y);
    z = z < 0 ? 0 : (z > 8 ? 8 : z);
    return read_voxel(words, off, (uint32_t)x, (uint32_t)y, (uint32_t)z) > 0u ? 1.0f : -1.0f;
}

/* compute_normal :109-121 */
static void compute_normal(const uint32_t* words, uint32_t off, int cx, int cy, int cz, float out[3]) {
    const float sx = occupancy(words, off, cx + 1, cy, cz) - occupancy(words, off, cx - 1, cy, cz);
    const float sy = occupancy(words, off, cx, cy + 1, cz) - occupancy(words, off, cx, cy - 1, cz);
    const float sz = occupancy(words, off, cx, cy, cz + 1) - occupancy(words, off, cx, cy, cz - 1);
    const float m2 = sx * sx + sy * sy + sz * sz;
    if (m2 >= 0.000001f) {
        const float inv = 1.0f / sqrtf(m2 > 0.000001f ? m2 : 0.000001f);
        out[0] = sx * inv;
        out[1] = sy * inv;
        out[2] = sz * inv;
    } else {
        out[0] = 0.0f;
        out[1] = 1.0f;
        out[2] = 0.0f;
    }
}

/* main :123-264 for one brick, cells in linear order.  vertices / normals: [2048][4] floats, indices: [2048].
 * *raw_count receives the un-clamped counter (the shader publishes min(counter, 2048)). */
int hvxo_brick_extract(const uint32_t* voxel_words, uint32_t data_offset, const float origin_size[4], float* vertices,
                       float* normals, uint32_t* indices, uint32_t* raw_count) {
    const float vs = origin_size[3];
    uint32_t counter = 0;
    for (uint32_t cell = 0; cell < 512u; ++cell) {
        const uint32_t cz = cell / 64u, cy = (cell / 8u) % 8u, cx = cell % 8u;
        const uint32_t corner[8] = {
            read_voxel(voxel_words, data_offset, cx, cy, cz),         read_voxel(voxel_words, data_offset, cx + 1, cy, cz),
            read_voxel(voxel_words, data_offset, cx + 1, cy + 1, cz), read_voxel(voxel_words, data_offset, cx, cy + 1, cz),
            read_voxel(voxel_words, data_offset, cx, cy, cz + 1),     read_voxel(voxel_words, data_offset, cx + 1, cy, cz + 1),
            read_voxel(voxel_words, data_offset, cx + 1, cy + 1, cz + 1), read_voxel(voxel_words, data_offset, cx, cy + 1, cz + 1)};
        uint32_t cube = 0, material = 0;
        for (int i = 7; i >= 0; --i)
            if (corner[i] != 0u) {
                cube |= 1u << i;
                material = corner[i];   /* ends as the FIRST non-zero corner, :202-208 */
            }
        if (cube == 0u || cube == 0xffu) continue;
        const uint8_t* row = HVXO_MC_EDGES[cube];      /* the packed_tri_table row, one edge per entry */
        uint32_t n = 0;
        while (n < 15u && row[n] != 0xffu) ++n;        /* :174-180, 0xF nibble == 255 here */
        if (n == 0u) continue;
        const uint32_t base = counter;
        counter += n;                                   /* both atomics advance by the same amount */
        if (base + n > MAX_VERTS) continue;             /* :193-195 -- the cell is dropped, the counter keeps its value */
        const float cell_world[3] = {(float)cx * vs + origin_size[0], (float)cy * vs + origin_size[1],
                                     (float)cz * vs + origin_size[2]};
        for (uint32_t i = 0; i < n; ++i) {
            const uint32_t edge = row[i];
            const float* mid = EDGE_MID[edge];
            float* v = vertices + 4u * (base + i);
            v[0] = cell_world[0] + mid[0] * vs;
            v[1] = cell_world[1] + mid[1] * vs;
            v[2] = cell_world[2] + mid[2] * vs;
            v[3] = (float)material;
            /* round() ties to even: 0.5 -> 0, so only a 1.0 component moves to the far corner, :222-228 */
            float nrm[3];
            compute_normal(voxel_words, data_offset, (int)cx + (mid[0] == 1.0f), (int)cy + (mid[1] == 1.0f),
                           (int)cz + (mid[2] == 1.0f), nrm);
            float* nn = normals + 4u * (base + i);
            nn[0] = nrm[0];
            nn[1] = nrm[1];
            nn[2] = nrm[2];
            nn[3] = 0.0f;
            indices[base + i] = base + i;
        }
    }
    *raw_count = counter;
    return 0;
}
