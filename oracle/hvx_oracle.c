/* hvx_oracle.c -- CPU ORACLE (test infrastructure, never shipped; see hvx_oracle.h).
 *
 * Plain-C restatement of the reference's CPU Transvoxel extractor.  Every
 * function cites the reference lines it follows (paths relative to
 * /root/reference/, PV = crates/passes/3d/helio-pass-planetary-voxel).
 * Float arithmetic is written one IEEE operation per C expression node, in the
 * reference's order, and must be compiled with -ffp-contract=off.
 */
#include "hvx_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "tables.inc"

/* ------------------------------------------------------------------------- */
/* CellWord: helio-planet-voxel-core/src/types.rs:328-354                     */

uint32_t hvxo_cellword(int16_t density, uint8_t material, uint8_t flags) {
    return (uint32_t)(uint16_t)density | ((uint32_t)material << 16) | ((uint32_t)flags << 24);
}
static inline int16_t cw_density(uint32_t w) { return (int16_t)(uint16_t)(w & 0xffffu); }
static inline uint32_t cw_material(uint32_t w) { return (w >> 16) & 0xffu; }
static inline int cw_solid(uint32_t w) { return cw_density(w) <= 0; }

/* ------------------------------------------------------------------------- */
/* Table decode + audit: PV/src/transvoxel.rs:201-238, 263-411                */

#define FNV_OFFSET 0xcbf29ce484222325ull
#define FNV_PRIME 0x00000100000001b3ull

static inline uint64_t fnv_byte(uint64_t f, uint8_t b) { return (f ^ (uint64_t)b) * FNV_PRIME; }

typedef struct {
    int kind; /* 0 regular, 1 transition */
    uint32_t case_index, class_index, reverse, nv, nt;
    const uint16_t* codes;
    const uint8_t* tris;
} topo_t;

/* PV/src/transvoxel.rs:201-216 */
static topo_t regular_case(uint32_t c) {
    topo_t t;
    t.kind = 0;
    t.case_index = c;
    t.class_index = HVXO_REGULAR_CELL_CLASS[c];
    uint8_t counts = HVXO_REGULAR_CELL_GEOMETRY_COUNTS[t.class_index];
    t.nv = counts >> 4;
    t.nt = counts & 0x0f;
    t.reverse = 0;
    t.codes = HVXO_REGULAR_VERTEX_DATA[c];
    t.tris = HVXO_REGULAR_CELL_VERTEX_INDEX[t.class_index];
    return t;
}

/* PV/src/transvoxel.rs:222-238 */
static topo_t transition_case(uint32_t c) {
    topo_t t;
    t.kind = 1;
    t.case_index = c;
    uint8_t code = HVXO_TRANSITION_CELL_CLASS[c];
    t.class_index = code & 0x7f;
    uint8_t counts = HVXO_TRANSITION_CELL_GEOMETRY_COUNTS[t.class_index];
    t.nv = counts >> 4;
    t.nt = counts & 0x0f;
    t.reverse = (code & 0x80) != 0;
    t.codes = HVXO_TRANSITION_VERTEX_DATA[c];
    t.tris = HVXO_TRANSITION_CELL_VERTEX_INDEX[t.class_index];
    return t;
}

/* PV/src/transvoxel.rs:148-156 triangle() with the inverse flip applied */
static inline void topo_triangle(const topo_t* t, uint32_t i, uint8_t out[3]) {
    const uint8_t* r = t->tris + 3 * i;
    out[0] = r[0];
    if (t->reverse) { out[1] = r[2]; out[2] = r[1]; } else { out[1] = r[1]; out[2] = r[2]; }
}

/* PV/src/transvoxel.rs:158-182 */
static uint64_t topo_fingerprint(const topo_t* t) {
    uint64_t f = FNV_OFFSET;
    f = fnv_byte(f, (uint8_t)t->kind);
    f = fnv_byte(f, (uint8_t)(t->case_index & 0xff));
    f = fnv_byte(f, (uint8_t)(t->case_index >> 8));
    f = fnv_byte(f, (uint8_t)t->class_index);
    f = fnv_byte(f, (uint8_t)t->reverse);
    f = fnv_byte(f, (uint8_t)t->nv);
    f = fnv_byte(f, (uint8_t)t->nt);
    for (uint32_t v = 0; v < t->nv; ++v) {
        f = fnv_byte(f, (uint8_t)(t->codes[v] & 0xff));
        f = fnv_byte(f, (uint8_t)(t->codes[v] >> 8));
    }
    for (uint32_t i = 0; i < t->nt; ++i) {
        uint8_t tri[3];
        topo_triangle(t, i, tri);
        for (int k = 0; k < 3; ++k) f = fnv_byte(f, tri[k]);
    }
    return f;
}

/* PV/src/transvoxel.rs:20-22 */
static const uint16_t CASE_WEIGHTS[9] = {0x001, 0x002, 0x004, 0x080, 0x100, 0x008, 0x040, 0x020, 0x010};
/* PV/src/transvoxel.rs:26 */
static const uint8_t DUPLICATE_CORNERS[4] = {0, 2, 6, 8};

/* PV/src/transvoxel.rs:40-49 */
static int transition_corner_is_solid(uint32_t case_index, uint32_t corner) {
    uint32_t full = corner < 9 ? corner : DUPLICATE_CORNERS[corner - 9];
    return (case_index & CASE_WEIGHTS[full]) != 0;
}

/* PV/src/transvoxel.rs:334-379 */
static int validate_topology(const topo_t* t) {
    uint32_t class_count = t->kind ? 56 : 16, max_v = 12, max_t = t->kind ? 12 : 5;
    uint32_t corners = t->kind ? 13 : 8;
    if (t->class_index >= class_count) return -1;
    if (t->nv > max_v || t->nt > max_t) return -2;
    for (uint32_t v = 0; v < t->nv; ++v) {
        uint32_t a = (t->codes[v] & 0xff) >> 4, b = t->codes[v] & 0x0f;
        if (a >= corners || b >= corners || a == b) return -3;
    }
    for (uint32_t i = 0; i < t->nt; ++i) {
        uint8_t tri[3];
        topo_triangle(t, i, tri);
        for (int k = 0; k < 3; ++k)
            if (tri[k] >= t->nv) return -4;
    }
    return 0;
}

int hvxo_validate_tables(hvxo_table_audit* out) {
    hvxo_table_audit a;
    memset(&a, 0, sizeof a);
    a.regular_cases = 256;
    a.transition_cases = 512;
    a.fingerprint = FNV_OFFSET;
    for (uint32_t c = 0; c < 256; ++c) {
        topo_t t = regular_case(c);
        int rc = validate_topology(&t);
        if (rc) return rc;
        for (uint32_t v = 0; v < t.nv; ++v) {
            uint32_t p = (t.codes[v] & 0xff) >> 4, q = t.codes[v] & 0x0f;
            if (((c >> p) & 1) == ((c >> q) & 1)) return -5; /* edge does not cross the surface */
        }
        a.regular_vertices += t.nv;
        a.regular_triangles += t.nt;
        if (t.nv > a.max_regular_vertices) a.max_regular_vertices = t.nv;
        if (t.nt > a.max_regular_triangles) a.max_regular_triangles = t.nt;
        uint64_t fp = topo_fingerprint(&t);
        for (int b = 0; b < 8; ++b) a.fingerprint = fnv_byte(a.fingerprint, (uint8_t)(fp >> (8 * b)));
    }
    for (uint32_t c = 0; c < 512; ++c) {
        topo_t t = transition_case(c);
        int rc = validate_topology(&t);
        if (rc) return rc;
        for (uint32_t v = 0; v < t.nv; ++v) {
            uint32_t p = (t.codes[v] & 0xff) >> 4, q = t.codes[v] & 0x0f;
            if (transition_corner_is_solid(c, p) == transition_corner_is_solid(c, q)) return -6;
        }
        a.transition_vertices += t.nv;
        a.transition_triangles += t.nt;
        if (t.nv > a.max_transition_vertices) a.max_transition_vertices = t.nv;
        if (t.nt > a.max_transition_triangles) a.max_transition_triangles = t.nt;
        uint64_t fp = topo_fingerprint(&t);
        for (int b = 0; b < 8; ++b) a.fingerprint = fnv_byte(a.fingerprint, (uint8_t)(fp >> (8 * b)));
    }
    if (out) *out = a;
    return 0;
}

const char* hvxo_table_revision(void) { return HVXO_TABLE_REVISION; }

int hvxo_case_topology(int kind, uint32_t case_index, uint32_t* class_index, uint32_t* reverse,
                       uint32_t* vertex_count, uint32_t* triangle_count, uint16_t codes[12],
                       uint8_t triangles[36]) {
    if ((kind == 0 && case_index >= 256) || (kind == 1 && case_index >= 512) || kind < 0 || kind > 1)
        return -1;
    topo_t t = kind ? transition_case(case_index) : regular_case(case_index);
    *class_index = t.class_index;
    *reverse = t.reverse;
    *vertex_count = t.nv;
    *triangle_count = t.nt;
    for (uint32_t v = 0; v < 12; ++v) codes[v] = v < t.nv ? t.codes[v] : 0;
    memset(triangles, 0, 36);
    for (uint32_t i = 0; i < t.nt; ++i) topo_triangle(&t, i, triangles + 3 * i);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* fBm terrain: crates/passes/3d/helio-pass-sdf/src/noise.rs                  */

/* Rust f32::fract = self - self.trunc() */
static inline float rs_fract(float x) { return x - truncf(x); }

/* noise.rs:9-32 */
static float hash3(float px, float py, float pz) {
    float qx = rs_fract(px * 0.3183099f + 0.1f);
    float qy = rs_fract(py * 0.3183099f + 0.1f);
    float qz = rs_fract(pz * 0.3183099f + 0.1f);
    if (qx < 0.0f) qx += 1.0f;
    if (qy < 0.0f) qy += 1.0f;
    if (qz < 0.0f) qz += 1.0f;
    qx *= 17.0f;
    qy *= 17.0f;
    qz *= 17.0f;
    float v = qx * qy * qz * (qx + qy + qz);
    float r = rs_fract(v);
    return r < 0.0f ? r + 1.0f : r;
}

static inline float lerpf(float a, float b, float t) { return a + (b - a) * t; }

/* noise.rs:36-67 */
static float noise3(float px, float py, float pz) {
    float ix = floorf(px), iy = floorf(py), iz = floorf(pz);
    float fx = px - ix, fy = py - iy, fz = pz - iz;
    float ux = fx * fx * fx * (fx * (fx * 6.0f - 15.0f) + 10.0f);
    float uy = fy * fy * fy * (fy * (fy * 6.0f - 15.0f) + 10.0f);
    float uz = fz * fz * fz * (fz * (fz * 6.0f - 15.0f) + 10.0f);
    float a = hash3(ix, iy, iz);
    float b = hash3(ix + 1.0f, iy, iz);
    float c = hash3(ix, iy + 1.0f, iz);
    float d = hash3(ix + 1.0f, iy + 1.0f, iz);
    float e = hash3(ix, iy, iz + 1.0f);
    float f = hash3(ix + 1.0f, iy, iz + 1.0f);
    float g = hash3(ix, iy + 1.0f, iz + 1.0f);
    float h = hash3(ix + 1.0f, iy + 1.0f, iz + 1.0f);
    float val = lerpf(lerpf(lerpf(a, b, ux), lerpf(c, d, ux), uy),
                      lerpf(lerpf(e, f, ux), lerpf(g, h, ux), uy), uz);
    return val * 2.0f - 1.0f;
}

/* noise.rs:78-96 (fbm_rotate inlined from :70-75) */
static float fbm2(float x, float z, uint32_t octaves, float lac, float persistence) {
    float value = 0.0f, amplitude = 1.0f, max_amp = 0.0f;
    float sx = x, sy = 0.0f, sz = z;
    for (uint32_t o = 0; o < octaves; ++o) {
        value += amplitude * noise3(sx, sy, sz);
        max_amp += amplitude;
        amplitude *= persistence;
        float rx = lac * (0.00f * sx + 0.80f * sy + 0.60f * sz);
        float ry = lac * (-0.80f * sx + 0.36f * sy - 0.48f * sz);
        float rz = lac * (-0.60f * sx - 0.48f * sy + 0.64f * sz);
        sx = rx;
        sy = ry;
        sz = rz;
    }
    return value / max_amp;
}

/* noise.rs:139-208 with TerrainStyle::Rolling and terrain.rs:40-51 rolling():
 * height -2, amplitude 4, frequency 0.08, 5 octaves, lacunarity 2, persistence 0.5 */
float hvxo_terrain_sdf_rolling(float x, float y, float z) {
    float fx = x * 0.08f;
    float fz = z * 0.08f;
    float terrain_height = fbm2(fx, fz, 5, 2.0f, 0.5f) * 4.0f;
    return y - (-2.0f + terrain_height);
}

/* ------------------------------------------------------------------------- */
/* Density fields: PV/src/fixture.rs:43-71                                    */

static inline int64_t sat_add(int64_t a, int64_t b) {
    int64_t r;
    if (__builtin_add_overflow(a, b, &r)) return b > 0 ? INT64_MAX : INT64_MIN;
    return r;
}
static inline int64_t sat_sub(int64_t a, int64_t b) {
    int64_t r;
    if (__builtin_sub_overflow(a, b, &r)) return b < 0 ? INT64_MAX : INT64_MIN;
    return r;
}
static inline int64_t sat_mul(int64_t a, int64_t b) {
    int64_t r;
    if (__builtin_mul_overflow(a, b, &r)) return ((a < 0) != (b < 0)) ? INT64_MIN : INT64_MAX;
    return r;
}
static inline int64_t sat_abs(int64_t a) { return a == INT64_MIN ? INT64_MAX : (a < 0 ? -a : a); }

static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

uint32_t hvxo_sample_canonical(int kind, const int64_t position[3], uint32_t lod) {
    int64_t x = position[0], y = position[1], z = position[2];
    int64_t density;
    switch (kind) {
    case HVXO_FIELD_PLANE:
    case HVXO_FIELD_MATERIAL_SEAM:
        density = sat_add(y, 1);
        break;
    case HVXO_FIELD_SPHERE:
        density = sat_sub(sat_add(sat_add(sat_mul(x, x), sat_mul(y, y)), sat_mul(z, z)), 144);
        break;
    case HVXO_FIELD_CAVE:
        density = sat_sub(144, sat_add(sat_add(sat_mul(x, x), sat_mul(y, y)), sat_mul(z, z)));
        break;
    case HVXO_FIELD_SHARP_CORNER:
        density = x > y ? x : y;
        density = density > z ? density : z;
        break;
    case HVXO_FIELD_THIN_SLAB:
        density = sat_sub(sat_abs(y), 1);
        break;
    case HVXO_FIELD_TERRAIN_FBM: {
        /* OURS: metres = lod0 cell * 0.1; density in 1/256 of a cell of this LOD,
         * round-half-even, clamped to i16.  No reference counterpart. */
        float sdf = hvxo_terrain_sdf_rolling((float)x * 0.1f, (float)y * 0.1f, (float)z * 0.1f);
        float cell_m = 0.1f * (float)(1u << (lod > 30 ? 30 : lod));
        float q = rintf((sdf / cell_m) * 256.0f);
        density = q < -32768.0f ? -32768 : (q > 32767.0f ? 32767 : (int64_t)q);
        int16_t d16 = (int16_t)density;
        return hvxo_cellword(d16, d16 <= 0 ? 1 : 0, 0);
    }
    case HVXO_FIELD_DENSE_RANDOM: {
        /* OURS (SURVEY 8d "dense adversarial"): position-hashed so page halos agree. */
        uint64_t h = splitmix64(((uint64_t)x * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)y * 0xC2B2AE3D27D4EB4Full) ^
                                ((uint64_t)z * 0x165667B19E3779F9ull) ^ 0xC0FFEEull);
        int16_t d16 = (int16_t)((int64_t)(h % 65535ull) - 32767);
        return hvxo_cellword(d16, d16 <= 0 ? (uint8_t)(1 + ((h >> 32) % 255ull)) : 0, 0);
    }
    default:
        density = INT16_MAX;
        break;
    }
    if (density < INT16_MIN) density = INT16_MIN;
    if (density > INT16_MAX) density = INT16_MAX;
    int16_t d16 = (int16_t)density;
    uint8_t material = 0;
    if (d16 <= 0) material = (kind == HVXO_FIELD_MATERIAL_SEAM && x >= 0) ? 2 : 1;
    return hvxo_cellword(d16, material, 0);
}

static int edge_ok(int edge) { return edge == 32 || edge == 64 || edge == 16 || edge == 8; }

/* PV/src/fixture.rs:95-124; page_min per helio-planet-voxel-core/src/types.rs:258-280 */
int hvxo_fixture_fill(int kind, int edge, uint32_t lod, const int64_t page_xyz[3], uint32_t* samples) {
    if (!edge_ok(edge) || lod > 57) return -1;
    int64_t span = (int64_t)edge << lod;
    int64_t scale = (int64_t)1 << lod;
    int64_t page_min[3];
    for (int a = 0; a < 3; ++a)
        if (__builtin_mul_overflow(page_xyz[a], span, &page_min[a])) return -2;
    size_t i = 0;
    for (int z = -1; z < edge + 1; ++z)
        for (int y = -1; y < edge + 1; ++y)
            for (int x = -1; x < edge + 1; ++x) {
                int local[3] = {x, y, z};
                int64_t p[3];
                for (int a = 0; a < 3; ++a)
                    if (__builtin_add_overflow(page_min[a], sat_mul((int64_t)local[a], scale), &p[a])) return -2;
                samples[i++] = hvxo_sample_canonical(kind, p, lod);
            }
    return 0;
}

/* PV/src/fixture.rs:200-209 cyclic corner order + PV/src/transvoxel.rs:187-199 adapter.
 * Net effect: regular case bit i <=> corner (i&1, i>>1&1, i>>2&1) solid. */
static const int CUBE_CORNERS[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0},
                                       {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
static const uint8_t FIXTURE_BIT_FOR_REGULAR_CORNER[8] = {0, 1, 3, 2, 4, 5, 7, 6};
/* PV/src/transvoxel.rs:54-63 */
static const int REGULAR_CORNERS[8][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {1, 1, 0},
                                          {0, 0, 1}, {1, 0, 1}, {0, 1, 1}, {1, 1, 1}};

typedef struct {
    int edge, s;
    const uint32_t* samples;
} grid_t;

/* PV/src/fixture.rs:211-220 fixture_index (bounds are the caller's business here) */
static inline uint32_t grid_sample(const grid_t* g, int x, int y, int z) {
    return g->samples[(size_t)(x + 1) + (size_t)(y + 1) * g->s + (size_t)(z + 1) * g->s * g->s];
}

/* PV/src/fixture.rs:146-162 cell_case, then PV/src/transvoxel.rs:187-199 */
static uint32_t regular_case_of_cell(const grid_t* g, int x, int y, int z) {
    uint32_t fixture_case = 0;
    for (int c = 0; c < 8; ++c)
        if (cw_solid(grid_sample(g, x + CUBE_CORNERS[c][0], y + CUBE_CORNERS[c][1], z + CUBE_CORNERS[c][2])))
            fixture_case |= 1u << c;
    uint32_t regular = 0;
    for (int rc = 0; rc < 8; ++rc)
        if (fixture_case & (1u << FIXTURE_BIT_FOR_REGULAR_CORNER[rc])) regular |= 1u << rc;
    return regular;
}

int hvxo_fixture_metrics_of(int edge, const uint32_t* samples, hvxo_fixture_metrics* out) {
    if (!edge_ok(edge)) return -1;
    grid_t g = {edge, edge + 2, samples};
    hvxo_fixture_metrics m;
    memset(&m, 0, sizeof m);
    uint64_t f = FNV_OFFSET;
    size_t n = (size_t)g.s * g.s * g.s;
    for (size_t i = 0; i < n; ++i) {
        if (cw_solid(samples[i])) m.solid_samples++; else m.air_samples++;
        for (int b = 0; b < 4; ++b) f = fnv_byte(f, (uint8_t)(samples[i] >> (8 * b)));
    }
    int mb = edge / 4;
    for (int z = 0; z < edge; ++z)
        for (int y = 0; y < edge; ++y)
            for (int x = 0; x < edge; ++x) {
                uint32_t c = regular_case_of_cell(&g, x, y, z);
                if (c != 0 && c != 255) {
                    m.active_cells++;
                    m.active_microbrick_mask |= 1ull << (x / mb + (y / mb) * 4 + (z / mb) * 16);
                }
            }
    m.fingerprint = f;
    *out = m;
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Regular extractor: PV/tests/gpu_transvoxel_emission.rs:272-450             */

static inline float density_at(const grid_t* g, int x, int y, int z) {
    return (float)cw_density(grid_sample(g, x, y, z));
}

/* PV/tests/gpu_transvoxel_emission.rs:402-420 */
static void gradient(const grid_t* g, const int p[3], float out[3]) {
    for (int a = 0; a < 3; ++a) {
        int lo[3] = {p[0], p[1], p[2]}, hi[3] = {p[0], p[1], p[2]};
        if (p[a] <= -1) {
            hi[a] += 1;
            out[a] = density_at(g, hi[0], hi[1], hi[2]) - density_at(g, p[0], p[1], p[2]);
        } else if (p[a] >= g->edge) {
            lo[a] -= 1;
            out[a] = density_at(g, p[0], p[1], p[2]) - density_at(g, lo[0], lo[1], lo[2]);
        } else {
            lo[a] -= 1;
            hi[a] += 1;
            out[a] = (density_at(g, hi[0], hi[1], hi[2]) - density_at(g, lo[0], lo[1], lo[2])) * 0.5f;
        }
    }
}

/* Rust f32::clamp: comparison based, keeps -0.0 */
static inline float rs_clamp01(float x) {
    if (x < 0.0f) x = 0.0f;
    if (x > 1.0f) x = 1.0f;
    return x;
}

/* interpolation parameter, emission.rs:305-312 / transvoxel_transition.rs:213-220 */
static inline float edge_t(float d0, float d1) {
    float den = d0 - d1;
    return fabsf(den) > 1.0e-12f ? rs_clamp01(d0 / den) : 0.5f;
}

/* PV/tests/gpu_transvoxel_emission.rs:346-400, thresholds generalised 31 -> edge-1 */
static void secondary_position(float p[3], const float n[3], uint32_t mask, float hi) {
    uint32_t near = 0;
    if (p[0] < 1.0f) near |= 1u;
    if (p[0] > hi) near |= 2u;
    if (p[1] < 1.0f) near |= 4u;
    if (p[1] > hi) near |= 8u;
    if (p[2] < 1.0f) near |= 16u;
    if (p[2] > hi) near |= 32u;
    if (near == 0 || (near & ~mask) != 0) return;
    float off[3] = {0.0f, 0.0f, 0.0f};
    for (int a = 0; a < 3; ++a) {
        if (near & (1u << (2 * a))) off[a] = (1.0f - p[a]) * 0.25f;
        else if (near & (2u << (2 * a))) off[a] = (hi - p[a]) * 0.25f;
    }
    float nc = ((off[0] * n[0]) + (off[1] * n[1])) + (off[2] * n[2]);
    for (int a = 0; a < 3; ++a) p[a] = (p[a] + off[a]) - (n[a] * nc);
}

int hvxo_extract_regular(int edge, const uint32_t* samples, uint64_t generation,
                         uint64_t dirty_microbricks, uint32_t transition_mask,
                         uint32_t max_vertices, uint32_t max_indices,
                         hvxo_vertex* vertices, uint32_t vertex_cap,
                         uint32_t* indices, uint32_t index_cap,
                         uint32_t* cell_words, uint32_t* cell_ranges,
                         uint32_t classify[4], uint32_t emission[8]) {
    if (!edge_ok(edge)) return -1;
    grid_t g = {edge, edge + 2, samples};
    const int mb = edge / 4;
    const float hi = (float)(edge - 1);
    uint32_t nv_total = 0, ni_total = 0;
    uint32_t cls[4] = {0, 0, 0, 0};
    for (int z = 0; z < edge; ++z)
        for (int y = 0; y < edge; ++y)
            for (int x = 0; x < edge; ++x) {
                size_t linear = (size_t)x + (size_t)y * edge + (size_t)z * edge * edge;
                uint32_t microbrick = (uint32_t)(x / mb + (y / mb) * 4 + (z / mb) * 16);
                if (!((dirty_microbricks >> microbrick) & 1ull)) continue;
                cls[0]++;
                topo_t t = regular_case(regular_case_of_cell(&g, x, y, z));
                if (cell_words) {
                    /* PV/src/transvoxel_gpu.rs:88-109 GpuTransvoxelCell::new */
                    cell_words[4 * linear + 0] = t.case_index | (t.class_index << 8) | (t.nv << 16) | (t.nt << 24) | 0x80000000u;
                    cell_words[4 * linear + 1] = (uint32_t)generation;
                    cell_words[4 * linear + 2] = (uint32_t)(generation >> 32);
                    cell_words[4 * linear + 3] = 0;
                }
                if (cell_ranges) {
                    cell_ranges[2 * linear + 0] = nv_total;
                    cell_ranges[2 * linear + 1] = ni_total;
                }
                if (t.nv) {
                    cls[1]++;
                    cls[2] += t.nv;
                    cls[3] += t.nt;
                }
                uint32_t first_vertex = nv_total;
                for (uint32_t k = 0; k < t.nv; ++k) {
                    uint32_t code = t.codes[k] & 0xff;
                    uint32_t c0 = code >> 4, c1 = code & 0x0f;
                    int A[3] = {x + REGULAR_CORNERS[c0][0], y + REGULAR_CORNERS[c0][1], z + REGULAR_CORNERS[c0][2]};
                    int B[3] = {x + REGULAR_CORNERS[c1][0], y + REGULAR_CORNERS[c1][1], z + REGULAR_CORNERS[c1][2]};
                    uint32_t wA = grid_sample(&g, A[0], A[1], A[2]);
                    uint32_t wB = grid_sample(&g, B[0], B[1], B[2]);
                    float d0 = (float)cw_density(wA), d1 = (float)cw_density(wB);
                    float t01 = edge_t(d0, d1);
                    float gA[3], gB[3], gm[3], p[3], n[3];
                    gradient(&g, A, gA);
                    gradient(&g, B, gB);
                    for (int a = 0; a < 3; ++a) {
                        float fa = (float)A[a], fb = (float)B[a];
                        p[a] = fa + (fb - fa) * t01;
                        gm[a] = gA[a] + (gB[a] - gA[a]) * t01;
                    }
                    /* normalize_or_up, emission.rs:442-450 */
                    float s = ((gm[0] * gm[0]) + (gm[1] * gm[1])) + (gm[2] * gm[2]);
                    if (s > 1.0e-12f) {
                        float inv = 1.0f / sqrtf(s);
                        for (int a = 0; a < 3; ++a) n[a] = gm[a] * inv;
                    } else {
                        n[0] = 0.0f; n[1] = 1.0f; n[2] = 0.0f;
                    }
                    secondary_position(p, n, transition_mask, hi);
                    if (nv_total < vertex_cap && vertices) {
                        hvxo_vertex* v = &vertices[nv_total];
                        for (int a = 0; a < 3; ++a) { v->position[a] = p[a]; v->normal[a] = n[a]; }
                        v->material = d0 <= 0.0f ? cw_material(wA) : cw_material(wB);
                        v->flags = 0;
                    }
                    nv_total++;
                }
                for (uint32_t i = 0; i < t.nt; ++i) {
                    uint8_t tri[3];
                    topo_triangle(&t, i, tri);
                    for (int k = 0; k < 3; ++k) {
                        if (ni_total < index_cap && indices) indices[ni_total] = first_vertex + tri[k];
                        ni_total++;
                    }
                }
            }
    if (classify) memcpy(classify, cls, sizeof cls);
    if (emission) {
        /* PV/src/transvoxel_emit.wgsl:191-200 + :297-299,368-369 */
        uint32_t vo = nv_total > max_vertices, io = ni_total > max_indices;
        emission[0] = nv_total;
        emission[1] = ni_total;
        emission[2] = (vo || io) ? 0 : nv_total;
        emission[3] = (vo || io) ? 0 : ni_total;
        emission[4] = vo;
        emission[5] = io;
        emission[6] = 1;
        emission[7] = 0;
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Transition faces: PV/src/transvoxel_transition.rs                          */

typedef struct { int origin[3], u[3], v[3], o[3]; } basis_t;
/* PV/src/transvoxel_transition.rs:463-502 (float twin :63-103) */
static const basis_t FACE_BASIS[6] = {
    {{0, 0, 1}, {0, 1, 0}, {0, 0, -1}, {-1, 0, 0}},
    {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 0}},
    {{1, 0, 0}, {0, 0, 1}, {-1, 0, 0}, {0, -1, 0}},
    {{0, 1, 0}, {0, 0, 1}, {1, 0, 0}, {0, 1, 0}},
    {{0, 1, 0}, {1, 0, 0}, {0, -1, 0}, {0, 0, -1}},
    {{0, 0, 1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}},
};
/* PV/src/transvoxel_transition.rs:28-38 */
static const int FULL_UV[9][2] = {{0, 0}, {1, 0}, {2, 0}, {0, 1}, {1, 1}, {2, 1}, {0, 2}, {1, 2}, {2, 2}};

/* PV/src/transvoxel_transition.rs:399-410 canonical_position */
static void face_canonical_position(const int64_t page_min[3], int64_t coarse, int64_t fine, int edge, int face,
                                    int fu, int fv, int layer, int64_t out[3]) {
    const basis_t* b = &FACE_BASIS[face];
    int64_t page_span = coarse * edge;
    for (int a = 0; a < 3; ++a)
        out[a] = page_min[a] + (int64_t)b->origin[a] * page_span + (int64_t)b->u[a] * fu * fine +
                 (int64_t)b->v[a] * fv * fine + (int64_t)b->o[a] * layer * fine;
}

int hvxo_slab_fill(int kind, int edge, uint32_t lod, const int64_t page_xyz[3], uint32_t* slabs) {
    if (!edge_ok(edge) || lod == 0 || lod > 57) return -1; /* FinestLodHasNoFinerNeighbor */
    int64_t coarse = (int64_t)1 << lod, fine = coarse / 2;
    int64_t page_min[3];
    for (int a = 0; a < 3; ++a)
        if (__builtin_mul_overflow(page_xyz[a], (int64_t)edge << lod, &page_min[a])) return -2;
    int se = 2 * edge + 1; /* TRANSITION_FACE_SAMPLE_EDGE */
    size_t i = 0;
    for (int face = 0; face < 6; ++face)
        for (int layer = -1; layer <= 1; ++layer)
            for (int v = -1; v <= se; ++v)
                for (int u = -1; u <= se; ++u) {
                    int64_t p[3];
                    face_canonical_position(page_min, coarse, fine, edge, face, u, v, layer, p);
                    slabs[i++] = hvxo_sample_canonical(kind, p, lod - 1);
                }
    return 0;
}

typedef struct {
    uint32_t word;
    float gradient[3];
} tsample_t;

/* One transition cell: PV/src/transvoxel_transition.rs:189-270 */
static void extract_transition_cell(int edge, int face, int cu, int cv, const tsample_t full[9], const topo_t* t,
                                    hvxo_vertex* out_v, uint32_t* out_i) {
    const basis_t* b = &FACE_BASIS[face];
    float origin[3], ua[3], va[3], oa[3];
    for (int a = 0; a < 3; ++a) {
        origin[a] = (float)b->origin[a] * (float)edge; /* transition_face_basis: 0.0 or edge */
        ua[a] = (float)b->u[a];
        va[a] = (float)b->v[a];
        oa[a] = (float)b->o[a];
    }
    for (uint32_t k = 0; k < t->nv; ++k) {
        uint32_t code = t->codes[k] & 0xff;
        uint32_t c0 = code >> 4, c1 = code & 0x0f;
        uint32_t f0 = c0 < 9 ? c0 : DUPLICATE_CORNERS[c0 - 9];
        uint32_t f1 = c1 < 9 ? c1 : DUPLICATE_CORNERS[c1 - 9];
        const tsample_t* s0 = &full[f0];
        const tsample_t* s1 = &full[f1];
        float d0 = (float)cw_density(s0->word), d1 = (float)cw_density(s1->word);
        float t01 = edge_t(d0, d1);
        /* transition_corner_uv :504-512, transition_corner_depth :514-520 */
        float u0 = (float)FULL_UV[f0][0] * 0.5f, v0 = (float)FULL_UV[f0][1] * 0.5f;
        float u1 = (float)FULL_UV[f1][0] * 0.5f, v1 = (float)FULL_UV[f1][1] * 0.5f;
        float fu = (float)cu + (u0 + (u1 - u0) * t01);
        float fv = (float)cv + (v0 + (v1 - v0) * t01);
        float dp0 = c0 < 9 ? 0.0f : 1.0f, dp1 = c1 < 9 ? 0.0f : 1.0f;
        float depth = dp0 + (dp1 - dp0) * t01;
        float gm[3], n[3];
        for (int a = 0; a < 3; ++a) gm[a] = s0->gradient[a] + (s1->gradient[a] - s0->gradient[a]) * t01;
        /* normalize_or_outward :531-539 */
        float s = ((gm[0] * gm[0]) + (gm[1] * gm[1])) + (gm[2] * gm[2]);
        if (s > 1.0e-12f) {
            float inv = 1.0f / sqrtf(s);
            for (int a = 0; a < 3; ++a) n[a] = gm[a] * inv;
        } else {
            for (int a = 0; a < 3; ++a) n[a] = oa[a];
        }
        /* basis.map(u, v, 0.0) :51-58 */
        float primary[3], inward[3];
        for (int a = 0; a < 3; ++a) {
            primary[a] = origin[a] + (((ua[a] * fu) + (va[a] * fv)) - (oa[a] * 0.0f));
            inward[a] = ((-oa[a]) * 0.25f) * depth;
        }
        /* project_onto_tangent :522-529, add3 */
        float nc = ((inward[0] * n[0]) + (inward[1] * n[1])) + (inward[2] * n[2]);
        for (int a = 0; a < 3; ++a) {
            out_v[k].position[a] = primary[a] + (inward[a] - (n[a] * nc));
            out_v[k].normal[a] = n[a];
        }
        out_v[k].material = d0 <= 0.0f ? cw_material(s0->word) : cw_material(s1->word);
        out_v[k].flags = 1u << face;
    }
    for (uint32_t i = 0; i < t->nt; ++i) {
        uint8_t tri[3];
        topo_triangle(t, i, tri); /* per-case inverse flip */
        out_i[3 * i + 0] = tri[0]; /* then the global flip [a, c, b], :262-267 */
        out_i[3 * i + 1] = tri[2];
        out_i[3 * i + 2] = tri[1];
    }
}

static uint32_t transition_case_of(const tsample_t full[9]) {
    uint32_t c = 0;
    for (int i = 0; i < 9; ++i)
        if (cw_solid(full[i].word)) c |= CASE_WEIGHTS[i];
    return c;
}

int hvxo_extract_transition(int edge, const uint32_t* slabs, uint32_t transition_mask,
                            uint64_t generation, uint32_t max_vertices, uint32_t max_indices,
                            hvxo_vertex* vertices, uint32_t vertex_cap,
                            uint32_t* indices, uint32_t index_cap,
                            uint32_t* cell_words, uint32_t* cell_ranges, uint32_t counters[12]) {
    if (!edge_ok(edge)) return -1;
    if (transition_mask & ~0x3fu) return -3; /* TransvoxelTransitionGpuError::TransitionMask */
    const int w = 2 * edge + 3;
    const size_t face_words = (size_t)w * w * 3;
    uint32_t nv_total = 0, ni_total = 0, active_cells = 0;
    for (int face = 0; face < 6; ++face) {
        size_t cell_base = (size_t)face * edge * edge;
        if (!((transition_mask >> face) & 1u)) {
            if (cell_words) memset(cell_words + 4 * cell_base, 0, sizeof(uint32_t) * 4 * edge * edge);
            continue;
        }
        const uint32_t* slab = slabs + face * face_words;
        const basis_t* b = &FACE_BASIS[face];
        for (int cv = 0; cv < edge; ++cv)
            for (int cu = 0; cu < edge; ++cu) {
                tsample_t full[9];
                for (int i = 0; i < 9; ++i) {
                    int su = 2 * cu + FULL_UV[i][0] + 1, sv = 2 * cv + FULL_UV[i][1] + 1;
                    full[i].word = slab[(size_t)su + (size_t)sv * w + (size_t)w * w];
                    /* gradient(): xyz central differences at +-fine_scale, *0.5 (:412-424), read
                     * from the slab halo: axis a lies along exactly one of u, v, outward. */
                    for (int a = 0; a < 3; ++a) {
                        int du = b->u[a], dv = b->v[a], dl = b->o[a];
                        uint32_t up = slab[(size_t)(su + du) + (size_t)(sv + dv) * w + (size_t)(1 + dl) * w * w];
                        uint32_t lo = slab[(size_t)(su - du) + (size_t)(sv - dv) * w + (size_t)(1 - dl) * w * w];
                        full[i].gradient[a] = ((float)cw_density(up) - (float)cw_density(lo)) * 0.5f;
                    }
                }
                topo_t t = transition_case(transition_case_of(full));
                size_t linear = cell_base + (size_t)cu + (size_t)cv * edge;
                if (cell_words) {
                    /* PV/src/transvoxel_transition_gpu.rs:76-95 */
                    uint32_t class_code = HVXO_TRANSITION_CELL_CLASS[t.case_index];
                    cell_words[4 * linear + 0] = t.case_index | (class_code << 9) | (t.nv << 17) | (t.nt << 21) | 0x80000000u;
                    cell_words[4 * linear + 1] = (uint32_t)generation;
                    cell_words[4 * linear + 2] = (uint32_t)(generation >> 32);
                    cell_words[4 * linear + 3] = 0;
                }
                if (cell_ranges) {
                    cell_ranges[2 * linear + 0] = nv_total;
                    cell_ranges[2 * linear + 1] = ni_total;
                }
                if (t.nv) active_cells++;
                hvxo_vertex cv_out[12];
                uint32_t ci_out[36];
                extract_transition_cell(edge, face, cu, cv, full, &t, cv_out, ci_out);
                for (uint32_t k = 0; k < t.nv; ++k)
                    if (nv_total + k < vertex_cap && vertices) vertices[nv_total + k] = cv_out[k];
                for (uint32_t i = 0; i < 3 * t.nt; ++i)
                    if (ni_total + i < index_cap && indices) indices[ni_total + i] = nv_total + ci_out[i];
                nv_total += t.nv;
                ni_total += 3 * t.nt;
            }
    }
    if (counters) {
        uint32_t vo = nv_total > max_vertices, io = ni_total > max_indices;
        memset(counters, 0, 12 * sizeof(uint32_t));
        counters[0] = active_cells;
        counters[1] = (uint32_t)__builtin_popcount(transition_mask & 0x3fu);
        counters[2] = nv_total;
        counters[3] = ni_total;
        counters[4] = (vo || io) ? 0 : nv_total;
        counters[5] = (vo || io) ? 0 : ni_total;
        counters[6] = vo;
        counters[7] = io;
        counters[8] = 1;
    }
    return 0;
}

int hvxo_extract_transition_face_analytic(int kind, int edge, uint32_t lod, const int64_t page_xyz[3], int face,
                                          hvxo_vertex* vertices, uint32_t vertex_cap,
                                          uint32_t* indices, uint32_t index_cap,
                                          uint32_t* vertex_count, uint32_t* index_count) {
    if (!edge_ok(edge) || lod == 0 || lod > 57 || face < 0 || face > 5) return -1;
    int64_t coarse = (int64_t)1 << lod, fine = coarse / 2;
    int64_t page_min[3];
    for (int a = 0; a < 3; ++a)
        if (__builtin_mul_overflow(page_xyz[a], (int64_t)edge << lod, &page_min[a])) return -2;
    uint32_t nv_total = 0, ni_total = 0;
    for (int cv = 0; cv < edge; ++cv)
        for (int cu = 0; cu < edge; ++cu) {
            tsample_t full[9];
            for (int i = 0; i < 9; ++i) {
                int64_t p[3];
                face_canonical_position(page_min, coarse, fine, edge, face, 2 * cu + FULL_UV[i][0],
                                        2 * cv + FULL_UV[i][1], 0, p);
                full[i].word = hvxo_sample_canonical(kind, p, lod - 1);
                for (int a = 0; a < 3; ++a) {
                    int64_t lo[3] = {p[0], p[1], p[2]}, hi[3] = {p[0], p[1], p[2]};
                    lo[a] -= fine;
                    hi[a] += fine;
                    full[i].gradient[a] = ((float)cw_density(hvxo_sample_canonical(kind, hi, lod - 1)) -
                                           (float)cw_density(hvxo_sample_canonical(kind, lo, lod - 1))) * 0.5f;
                }
            }
            topo_t t = transition_case(transition_case_of(full));
            hvxo_vertex cv_out[12];
            uint32_t ci_out[36];
            extract_transition_cell(edge, face, cu, cv, full, &t, cv_out, ci_out);
            for (uint32_t k = 0; k < t.nv; ++k)
                if (nv_total + k < vertex_cap) vertices[nv_total + k] = cv_out[k];
            for (uint32_t i = 0; i < 3 * t.nt; ++i)
                if (ni_total + i < index_cap) indices[ni_total + i] = nv_total + ci_out[i];
            nv_total += t.nv;
            ni_total += 3 * t.nt;
        }
    *vertex_count = nv_total;
    *index_count = ni_total;
    return 0;
}

/* ------------------------------------------------------------------------- */
/* CPU baseline batch driver (OpenMP over chunks)                  */

int hvxo_batch_fill(int kind, int edge, uint32_t lod, const int64_t* page_xyz, uint32_t n, int threads, uint32_t* out) {
    if (!edge_ok(edge)) return -1;
    const size_t s = (size_t)edge + 2, chunk_words = s * s * s;
    int rc = 0;
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(dynamic) num_threads(threads)
    for (uint32_t c = 0; c < n; ++c) {
        int r = hvxo_fixture_fill(kind, edge, lod, page_xyz + 3 * (size_t)c, out + (size_t)c * chunk_words);
        if (r) {
#pragma omp critical
            rc = r;
        }
    }
    return rc;
}

int hvxo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int64_t hvxo_batch_regular(int kind, int edge, uint32_t lod, const int64_t* page_xyz, uint32_t n, int do_fill,
                           const uint32_t* samples_or_null, int threads, uint64_t totals[4]) {
    if (!edge_ok(edge)) return -1;
    const size_t s = (size_t)edge + 2, chunk_words = s * s * s;
    uint64_t tv = 0, ti = 0, ta = 0, tc = 0;
    if (threads < 1) threads = 1;
#pragma omp parallel num_threads(threads) reduction(+ : tv, ti, ta, tc)
    {
        uint32_t* scratch = do_fill ? (uint32_t*)malloc(chunk_words * sizeof(uint32_t)) : NULL;
        /* worst case 12 vertices / 15 indices per cell would be 100 MB at edge 64; the baseline
         * keeps a bounded mesh buffer and counts the rest (writes beyond the cap are skipped). */
        const uint32_t vcap = 1u << 18, icap = 3u << 17;
        hvxo_vertex* vbuf = (hvxo_vertex*)malloc(sizeof(hvxo_vertex) * vcap);
        uint32_t* ibuf = (uint32_t*)malloc(sizeof(uint32_t) * icap);
#pragma omp for schedule(dynamic)
        for (uint32_t c = 0; c < n; ++c) {
            const uint32_t* smp;
            if (do_fill) {
                hvxo_fixture_fill(kind, edge, lod, page_xyz + 3 * (size_t)c, scratch);
                smp = scratch;
            } else {
                smp = samples_or_null + (size_t)c * chunk_words;
            }
            uint32_t cls[4], em[8];
            hvxo_extract_regular(edge, smp, 1, ~0ull, 0, 0xffffffffu, 0xffffffffu, vbuf, vcap, ibuf, icap, NULL, NULL,
                                 cls, em);
            tv += em[0];
            ti += em[1];
            ta += cls[1];
            tc += em[0] != 0;
        }
        free(scratch);
        free(vbuf);
        free(ibuf);
    }
    if (totals) {
        totals[0] = tv;
        totals[1] = ti;
        totals[2] = ta;
        totals[3] = tc;
    }
    return (int64_t)n * edge * edge * edge;
}

/* ------------------------------------------------------------------------- */
/* Meshlet build: PV/src/terrain_meshlet.rs:83-249              */

static inline float m_dot(const float a[3], const float b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void m_sub(const float a[3], const float b[3], float o[3]) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }
static inline void m_cross(const float a[3], const float b[3], float o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
/* normalize (:309-312): magnitude > f32::EPSILON, scale by magnitude.recip() */
static inline int m_normalize(const float v[3], float o[3]) {
    float magnitude = sqrtf(m_dot(v, v));
    if (!(magnitude > 1.1920929e-7f)) return 0;
    float inv = 1.0f / magnitude;
    o[0] = v[0] * inv; o[1] = v[1] * inv; o[2] = v[2] * inv;
    return 1;
}
static int m_triangle_normal(const hvxo_vertex* v, const uint32_t* tri, float n[3]) {
    float ab[3], ac[3], c[3];
    m_sub(v[tri[1]].position, v[tri[0]].position, ab);
    m_sub(v[tri[2]].position, v[tri[0]].position, ac);
    m_cross(ab, ac, c);
    return m_normalize(c, n);
}

/* compute_bounds (:184-249) for one chunk of <= 63 indices */
static void meshlet_bounds(const hvxo_vertex* v, const uint32_t* idx, uint32_t n, hvxo_meshlet_bounds* out) {
    float mn[3], mx[3];
    for (int a = 0; a < 3; ++a) mn[a] = mx[a] = v[idx[0]].position[a];
    for (uint32_t i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) {
            mn[a] = fminf(mn[a], v[idx[i]].position[a]);
            mx[a] = fmaxf(mx[a], v[idx[i]].position[a]);
        }
    float center[3];
    for (int a = 0; a < 3; ++a) center[a] = (mn[a] + mx[a]) * 0.5f;
    float radius = 0.0f;
    for (uint32_t i = 0; i < n; ++i) {
        float d[3];
        m_sub(v[idx[i]].position, center, d);
        radius = fmaxf(radius, sqrtf(m_dot(d, d)));
    }
    memset(out, 0, sizeof *out);
    for (int a = 0; a < 3; ++a) { out->center[a] = center[a]; out->cone_apex[a] = center[a]; }
    out->radius = radius;
    out->cone_cutoff = 1.0f; /* disabled_bounds (:251-260) until proven otherwise */
    float sum[3] = {0.0f, 0.0f, 0.0f}, axis[3];
    for (uint32_t t = 0; t + 2 < n; t += 3) {
        float nrm[3];
        if (m_triangle_normal(v, idx + t, nrm))
            for (int a = 0; a < 3; ++a) sum[a] = sum[a] + nrm[a];
    }
    if (!m_normalize(sum, axis)) return;
    float min_dot = 1.0f;
    for (uint32_t t = 0; t + 2 < n; t += 3) {
        float nrm[3];
        if (m_triangle_normal(v, idx + t, nrm)) min_dot = fminf(min_dot, m_dot(nrm, axis));
    }
    if (min_dot <= 0.1f) return;
    float apex_distance = 0.0f;
    for (uint32_t t = 0; t + 2 < n; t += 3) {
        float nrm[3], ca[3];
        if (!m_triangle_normal(v, idx + t, nrm)) continue;
        float denominator = m_dot(axis, nrm);
        if (denominator > 0.0f) {
            m_sub(center, v[idx[t]].position, ca);
            apex_distance = fmaxf(apex_distance, m_dot(ca, nrm) / denominator);
        }
    }
    for (int a = 0; a < 3; ++a) {
        out->cone_apex[a] = center[a] - axis[a] * apex_distance;
        out->cone_axis[a] = axis[a];
    }
    out->cone_cutoff = fminf(sqrtf(fmaxf(1.0f - min_dot * min_dot, 0.0f)) + 1.0e-4f, 1.0f);
}

int hvxo_build_meshlets(const hvxo_vertex* vertices, uint32_t vertex_count, const uint32_t* indices,
                        uint32_t index_count, uint32_t first_index, uint32_t first_vertex, uint32_t first_bounds,
                        uint64_t generation, uint32_t flags, hvxo_meshlet* meshlets, hvxo_meshlet_bounds* bounds,
                        uint32_t capacity) {
    if (index_count % 3 != 0) return -1;                                   /* IncompleteTriangle */
    for (uint32_t i = 0; i < vertex_count; ++i)
        for (int a = 0; a < 3; ++a)
            if (!isfinite(vertices[i].position[a])) return -3;              /* NonFinitePosition */
    for (uint32_t i = 0; i < index_count; ++i)
        if (indices[i] >= vertex_count) return -2;                          /* IndexOutOfBounds */
    uint32_t count = (index_count + 62) / 63;
    if (count > capacity) return -4;
    for (uint32_t m = 0; m < count; ++m) {
        const uint32_t* chunk = indices + 63 * m;
        uint32_t n = index_count - 63 * m < 63 ? index_count - 63 * m : 63;
        uint32_t unique = 0;                                                /* unique_index_count (:171-182) */
        for (uint32_t i = 0; i < n; ++i) {
            int seen = 0;
            for (uint32_t j = 0; j < i; ++j)
                if (chunk[j] == chunk[i]) { seen = 1; break; }
            unique += !seen;
        }
        hvxo_meshlet* d = &meshlets[m];
        d->first_index = first_index + 63 * m;
        d->index_count = n;
        d->first_vertex = first_vertex;
        d->vertex_count = unique;
        d->bounds_offset = first_bounds + m;
        d->generation_low = (uint32_t)generation;
        d->generation_high = (uint32_t)(generation >> 32);
        d->_pad = flags;
        meshlet_bounds(vertices, chunk, n, &bounds[m]);
    }
    return (int)count;
}

/* ================================================================================================
 * Surface gather.  Follows PV/src/surface_gather.wgsl line by line: every gathered
 * sample walks the open-addressed page table (lookup_page, :125-149), missing pages read as AIR and
 * count a miss (:181-186), counters are the shader's atomics, finalize_gather (:236-264) publishes
 * `completed` and the eight DispatchIndirectArgs.  i32 arithmetic wraps like WGSL's.
 * ============================================================================================== */
static uint32_t gather_mix_hash(uint32_t hash, uint32_t value) { /* table.rs:163-167, wgsl:102-105 */
    uint32_t mixed = (hash ^ value) * 0x045d9f3bu;
    return mixed ^ (mixed >> 16);
}

uint32_t hvxo_page_hash(const uint32_t planet_id[4], const int32_t relative_min[3], uint32_t lod) {
    uint32_t hash = 0x811c9dc5u;
    for (int i = 0; i < 4; ++i) hash = gather_mix_hash(hash, planet_id[i]);
    for (int i = 0; i < 3; ++i) hash = gather_mix_hash(hash, (uint32_t)relative_min[i]);
    return gather_mix_hash(hash, lod);
}

typedef struct {
    uint32_t slot, generation_low, generation_high, probes, found;
} gather_lookup;

static gather_lookup gather_lookup_page(const hvxo_residency* res, const hvxo_page_table_entry* table,
                                        const hvxo_gather_job* job, const int32_t rel[3], uint32_t lod) {
    const uint32_t start = hvxo_page_hash(job->planet_id, rel, lod) & res->table_mask;
    uint32_t probe = 0;
    for (;;) {
        if (probe >= res->max_probe) break;
        const hvxo_page_table_entry* e = &table[(start + probe) & res->table_mask];
        if (e->state == 0u) return (gather_lookup){0u, 0u, 0u, probe + 1u, 0u};
        if (e->state == 1u && memcmp(e->planet_id, job->planet_id, 16) == 0 && memcmp(e->relative_lod0_cell_min, rel, 12) == 0 &&
            e->lod == lod)
            return (gather_lookup){e->slot, e->generation_low, e->generation_high, probe + 1u, 1u};
        probe += 1u;
    }
    return (gather_lookup){0u, 0u, 0u, probe, 0u};
}

static int32_t wrap_add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static int32_t wrap_sub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
static int32_t wrap_mul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
static int32_t gather_floor_div(int32_t value, int32_t divisor) { /* wgsl:151-157 */
    int32_t q = value / divisor;
    if (value % divisor < 0) q -= 1;
    return q;
}

static uint32_t gather_sample(const hvxo_residency* res, const hvxo_page_table_entry* table, const uint32_t* atlas,
                              const hvxo_gather_job* job, const int32_t pos[3], uint32_t lod, hvxo_gather_counters* c) {
    const int32_t scale = (int32_t)(1u << lod), span = 32 * scale; /* wgsl:166-168 */
    int32_t rel[3];
    for (int a = 0; a < 3; ++a) {
        const int32_t off = wrap_sub(pos[a], job->relative_lod0_cell_min[a]);
        rel[a] = wrap_add(job->relative_lod0_cell_min[a], wrap_mul(gather_floor_div(off, span), span));
    }
    const gather_lookup l = gather_lookup_page(res, table, job, rel, lod);
    c->table_probes += l.probes;
    if (!l.found) {
        c->page_misses += 1u;
        return 0x00007fffu;
    }
    uint32_t local[3];
    for (int a = 0; a < 3; ++a) local[a] = (uint32_t)(wrap_sub(pos[a], rel[a]) / scale);
    const uint32_t tx = l.slot % res->atlas_tiles_x, ty = (l.slot / res->atlas_tiles_x) % res->atlas_tiles_y,
                   tz = l.slot / (res->atlas_tiles_x * res->atlas_tiles_y); /* slot_origin, wgsl:159-164 */
    const uint64_t row = (uint64_t)res->atlas_tiles_x * 32u, slice = row * res->atlas_tiles_y * 32u;
    return atlas[(tx * 32u + local[0]) + (ty * 32u + local[1]) * row + (tz * 32u + local[2]) * slice];
}

void hvxo_gather_surface(const hvxo_residency* res, const hvxo_page_table_entry* table, const uint32_t* atlas,
                         const hvxo_gather_job* job, uint32_t* regular, uint32_t* transition,
                         hvxo_gather_counters* c, uint32_t* indirect) {
    static const int32_t ORIGIN[6][3] = {{0, 0, 1}, {1, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 1, 0}, {0, 0, 1}};
    static const int32_t U[6][3] = {{0, 1, 0}, {0, 1, 0}, {0, 0, 1}, {0, 0, 1}, {1, 0, 0}, {1, 0, 0}};
    static const int32_t V[6][3] = {{0, 0, -1}, {0, 0, 1}, {-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}};
    static const int32_t OUT[6][3] = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}};
    memset(c, 0, sizeof(*c));
    const int epoch_ok = res->publication_epoch_low == job->residency_epoch_low &&
                         res->publication_epoch_high == job->residency_epoch_high;
    if (epoch_ok) {
        /* gather_regular, wgsl:200-212 */
        const int32_t scale = (int32_t)(1u << job->lod);
        for (int32_t linear = 0; linear < 39304; ++linear) {
            const int32_t x = linear % 34, y = (linear / 34) % 34, z = linear / (34 * 34);
            const int32_t local[3] = {x - 1, y - 1, z - 1};
            int32_t pos[3];
            for (int a = 0; a < 3; ++a) pos[a] = wrap_add(job->relative_lod0_cell_min[a], wrap_mul(local[a], scale));
            regular[linear] = gather_sample(res, table, atlas, job, pos, job->lod, c);
            c->regular_samples += 1u;
        }
        /* gather_transition, wgsl:214-241 */
        if (job->lod != 0u) {
            const int32_t fine = (int32_t)(1u << (job->lod - 1u)), coarse_span = 32 * fine * 2;
            for (int32_t linear = 0; linear < 80802; ++linear) {
                const int32_t face = linear / 13467;
                if ((job->transition_mask & (1u << face)) == 0u) continue;
                const int32_t fl = linear % 13467, layer = fl / (67 * 67), ll = fl % (67 * 67), v = ll / 67, u = ll % 67;
                int32_t pos[3];
                for (int a = 0; a < 3; ++a)
                    pos[a] = wrap_add(wrap_add(wrap_add(wrap_add(job->relative_lod0_cell_min[a], wrap_mul(ORIGIN[face][a], coarse_span)),
                                                        wrap_mul(wrap_mul(U[face][a], u - 1), fine)),
                                               wrap_mul(wrap_mul(V[face][a], v - 1), fine)),
                                      wrap_mul(wrap_mul(OUT[face][a], layer - 1), fine));
                transition[linear] = gather_sample(res, table, atlas, job, pos, job->lod - 1u, c);
                c->transition_samples += 1u;
            }
        }
    }
    /* finalize_gather, wgsl:236-264 */
    const gather_lookup target = gather_lookup_page(res, table, job, job->relative_lod0_cell_min, job->lod);
    c->table_probes += target.probes;
    const int current = target.found && target.slot == job->target_slot && target.generation_low == job->generation_low &&
                        target.generation_high == job->generation_high;
    if (!current || !epoch_ok) {
        c->stale_targets = 1u;
        return;
    }
    uint32_t faces = 0;
    for (int f = 0; f < 6; ++f) faces += (job->transition_mask >> f) & 1u;
    if (c->page_misses != 0u || c->regular_samples != 39304u || c->transition_samples != faces * 13467u) return;
    c->completed = 1u;
    static const uint32_t GROUPS[8] = {512u, 128u, 1u, 512u, 96u, 24u, 1u, 96u};
    for (int i = 0; i < 8; ++i) {
        indirect[3 * i] = GROUPS[i];
        indirect[3 * i + 1] = 1u;
        indirect[3 * i + 2] = 1u;
    }
}

/* ================================================================================================
 * Surface publication, PV/src/surface_publish.wgsl:104-225 line by line.
 * ============================================================================================== */
static int publish_metadata_is_current(const hvxo_surface_job* job, const hvxo_page_meta* meta) { /* :107-112 */
    const hvxo_page_meta* m = &meta[job->slot];
    return m->slot == job->slot && m->generation_low == job->generation_low && m->generation_high == job->generation_high;
}

void hvxo_publish_surface(const hvxo_surface_job* job, const hvxo_page_meta* meta, const uint32_t* rc, const uint32_t* tc,
                          const hvxo_vertex* src_v, const uint32_t* src_i, const hvxo_vertex* src_tv, const uint32_t* src_ti,
                          hvxo_surface_state* states, hvxo_vertex* v, uint32_t* idx, hvxo_vertex* tv, uint32_t* tidx,
                          hvxo_draw_args* rdraw, hvxo_draw_args* tdraw, hvxo_surface_feedback* fb) {
    /* emission counters: required_v, required_i, emitted_v, emitted_i, v_overflow, i_overflow, completed, pad
     * transition counters: active_cells, active_faces, required_v, required_i, emitted_v, emitted_i, v_ovf, i_ovf, completed */
    const uint32_t r_ev = rc[2], r_ei = rc[3], r_vo = rc[4], r_io = rc[5], r_done = rc[6];
    const uint32_t t_ev = tc[4], t_ei = tc[5], t_vo = tc[6], t_io = tc[7], t_done = tc[8];
    const int current = publish_metadata_is_current(job, meta);
    const uint32_t active = states[job->slot].active_bank;
    const uint32_t next_bank = 1u - (active < 1u ? active : 1u);
    const uint32_t bank = job->slot * 2u + next_bank;
    /* copy_regular_surface, :125-136 */
    if (r_done != 0u && r_vo == 0u && r_io == 0u && current) {
        if (v && src_v) memcpy(v + (size_t)bank * job->regular_max_vertices, src_v, (size_t)r_ev * sizeof(hvxo_vertex));
        if (idx && src_i) memcpy(idx + (size_t)bank * job->regular_max_indices, src_i, (size_t)r_ei * 4u);
    }
    /* copy_transition_surface, :151-162 */
    if (t_done != 0u && t_vo == 0u && t_io == 0u && current) {
        if (tv && src_tv) memcpy(tv + (size_t)bank * job->transition_max_vertices, src_tv, (size_t)t_ev * sizeof(hvxo_vertex));
        if (tidx && src_ti) memcpy(tidx + (size_t)bank * job->transition_max_indices, src_ti, (size_t)t_ei * 4u);
    }
    /* publish_surface, :166-216 */
    fb->submitted_jobs += 1u;
    if (!current) {
        fb->stale_rejections += 1u;
        return;
    }
    if (r_done == 0u || t_done == 0u) {
        fb->incomplete_rejections += 1u;
        return;
    }
    if (r_vo != 0u || r_io != 0u || t_vo != 0u || t_io != 0u) {
        fb->overflow_rejections += 1u;
        return;
    }
    hvxo_surface_state s;
    memset(&s, 0, sizeof(s));
    s.generation_low = job->generation_low;
    s.generation_high = job->generation_high;
    s.active_bank = next_bank;
    s.valid = 1u;
    s.regular_vertex_count = r_ev;
    s.regular_index_count = r_ei;
    s.transition_vertex_count = t_ev;
    s.transition_index_count = t_ei;
    s.regular_meshlet_count = (r_ei + 62u) / 63u;
    s.transition_meshlet_count = (t_ei + 62u) / 63u;
    states[job->slot] = s;
    rdraw[job->slot] = (hvxo_draw_args){r_ei, 0u, bank * job->regular_max_indices, (int32_t)(bank * job->regular_max_vertices), job->slot};
    tdraw[job->slot] = (hvxo_draw_args){t_ei, 0u, bank * job->transition_max_indices, (int32_t)(bank * job->transition_max_vertices), job->slot};
    fb->published_jobs += 1u;
}

void hvxo_refresh_visibility(uint32_t slots, const hvxo_surface_state* states, const uint32_t* visible,
                             hvxo_draw_args* rdraw, hvxo_draw_args* tdraw) {
    for (uint32_t i = 0; i < slots; ++i) {
        const uint32_t vis = (states[i].valid != 0u && visible[i] != 0u) ? 1u : 0u;
        rdraw[i].instance_count = vis;
        tdraw[i].instance_count = vis;
    }
}
