// regular_extract.cu -- fused regular-cell Transvoxel extraction for sm_100a.
//
// Replaces the reference's four dispatches per page
//   classify_regular_cells  PV/src/transvoxel_classify.wgsl:75-120
//   scan_regular_cells      PV/src/transvoxel_emit.wgsl:131-172
//   scan_regular_blocks     PV/src/transvoxel_emit.wgsl:174-202
//   emit_regular_cells      PV/src/transvoxel_emit.wgsl:291-370
// with ONE persistent kernel over a batch of chunks.  Arithmetic follows the reference's CPU
// extractor (PV/tests/gpu_transvoxel_emission.rs:272-450), not the WGSL, so positions and
// normals are bit-identical to the oracle.
//
// Design:
//   * one CTA per SM pulls chunks from an atomic queue and walks each chunk front to back in z.
//     x-fastest cell order == walk order, so vertex/index placement is a running prefix inside
//     the CTA: order preserving by construction, no cross-CTA scan on the data path.
//   * sample layers ((E+2)^2 words, contiguous, 16-byte aligned) are streamed HBM -> shared
//     memory with cp.async.bulk (TMA 1-D, SASS UBLKCP) into a ring of R layers, completion on
//     one mbarrier per slot; the producer runs up to R-(ZB+3) layers ahead, across chunk
//     boundaries.  Every sample is read from HBM exactly once.
//   * classification is bit-parallel: warp ballots turn a layer into one solid-bit row mask
//     per sample row, and a cell row's 8-corner signs are 8 shifted copies of 4 row masks; the
//     active-cell mask of a 64-cell row costs one thread ~40 integer ops.  Occupancy counts
//     are popc of those masks.
//   * active cells are compacted in order (row popc -> block scan -> rank scatter), their
//     (vertex, index) counts scanned with warp shuffles, and vertices are emitted one thread
//     per VERTEX (not per cell) from the shared-memory bricks: 14 smem loads, no HBM re-read.
//   * no single-address atomics: the reference's 5 atomicAdd per cell become one counter
//     record written once per chunk.
#include <cstdlib>

#include "hvx_device.cuh"
#include "hvx_kernels.h"

namespace hvx {

namespace {

// Lengyel's tables live in device global memory (statically initialised); every CTA copies the
// 3.8 KB it needs into shared memory once (the kernel is persistent).
#define HVX_TABLE static __device__ const
#include "transvoxel_tables.inc"

template <int E_, int ZB_, int R_, int NT_>
struct Cfg {
    static constexpr int E = E_;          // cells per chunk edge
    static constexpr int S = E_ + 2;      // samples per edge (1-sample halo)
    static constexpr int ZB = ZB_;        // cell layers per step
    static constexpr int R = R_;          // ring slots (sample layers resident or in flight)
    static constexpr int NT = NT_;        // threads per CTA
    static constexpr int NW = NT_ / 32;
    static constexpr int WIN = ZB_ + 3;   // sample layers a step touches (corners + gradient halo)
    static constexpr int LAYER_WORDS = S * S;
    static constexpr int LAYER_BYTES = LAYER_WORDS * 4;
    static constexpr int NB = (LAYER_WORDS + 31) / 32;  // 32-sample ballot blocks per layer
    static constexpr int BW = NB + 3;                   // + zero padding for the 4-word row window
    static constexpr int ROWS = ZB_ * E_;  // cell rows per step
    static constexpr int ROW_WARPS = ROWS / 32;
    static constexpr int CB = NT_;         // active cells per emission batch
    static constexpr int QW = E_ / 4;      // microbrick edge == quarter-row width
    static constexpr int RPB = 256 / E_;   // rows per 256-cell scan block
    static_assert(R_ >= ZB_ + 4, "ring must hold one window plus at least one layer in flight");
    static_assert(R_ < E_ + 2, "producer may be at most one chunk ahead");
    static_assert(LAYER_BYTES % 16 == 0, "cp.async.bulk needs 16-byte multiples");
    static_assert(E_ % ZB_ == 0 && ROWS <= NT_ && ROWS % 32 == 0 && (E_ == 32 || E_ == 64), "unsupported tiling");
};

template <class C>
struct Smem {
    alignas(128) uint32_t ring[C::R][C::LAYER_WORDS];
    alignas(8) uint64_t full_bar[C::R];
    uint64_t active[C::ROWS];       // active-cell bits of each cell row of the step
    uint64_t dirty_row[16];         // [my + 4*mz] -> x mask of dirty microbricks
    union {
        uint16_t owner[C::CB * 12];        // emission: vertex -> cell slot | k<<10
        uint64_t row_pref64[C::ROWS + 1];  // debug records: per-row exclusive (vertices | indices<<32)
    };
    uint64_t scan64[2][34];
    uint32_t scan32[2][34];
    uint32_t bits[C::R][C::BW];     // solid bit of every sample of a layer, flat x-fastest order
    uint32_t row_off[C::ROWS + 1];  // exclusive prefix of popc(active)
    uint32_t cell_rec[C::CB];       // x | row<<8 | case<<16
    uint32_t chunk_ids[4];
    uint16_t case_info[256];
    uint8_t vertex_edge[256 * 12];
    uint8_t class_index[16 * 16];
};

// Solid bits of the 8 corners of every cell of one cell row: bit x of a** is the corner at
// sample x+1 (cell-local x), bit x of b** the corner at x+2; 10 = next sample row, 01 = next layer.
struct RowCorners {
    uint64_t a00, b00, a10, b10, a01, b01, a11, b11;
};

// Bits [o+1, o+1+E) and [o+2, o+2+E) of a layer's flat solid-bit array, o = row * S.
template <class C>
__device__ __forceinline__ void row_window(const uint32_t* __restrict__ bits, int row, uint64_t& a, uint64_t& b) {
    const int p = row * C::S + 1, w = p >> 5, sh = p & 31;
    const uint32_t x0 = bits[w], x1 = bits[w + 1], x2 = bits[w + 2];
    if (C::E == 64) {
        const uint32_t x3 = bits[w + 3];
        const uint32_t lo = __funnelshift_r(x0, x1, sh), hi = __funnelshift_r(x1, x2, sh);
        const uint32_t nx = __funnelshift_r(x2, x3, sh);
        a = static_cast<uint64_t>(lo) | (static_cast<uint64_t>(hi) << 32);
        b = (a >> 1) | (static_cast<uint64_t>(nx & 1u) << 63);
    } else {
        const uint32_t lo = __funnelshift_r(x0, x1, sh), nx = __funnelshift_r(x1, x2, sh);
        a = lo;
        b = (lo >> 1) | ((nx & 1u) << 31);
    }
}

template <class C>
__device__ __forceinline__ RowCorners load_row_corners(const Smem<C>& sm, int slot0, int slot1, int y) {
    RowCorners rc;
    row_window<C>(sm.bits[slot0], y + 1, rc.a00, rc.b00);
    row_window<C>(sm.bits[slot0], y + 2, rc.a10, rc.b10);
    row_window<C>(sm.bits[slot1], y + 1, rc.a01, rc.b01);
    row_window<C>(sm.bits[slot1], y + 2, rc.a11, rc.b11);
    return rc;
}

__device__ __forceinline__ uint32_t case_at(const RowCorners& rc, int x) {
    return static_cast<uint32_t>((rc.a00 >> x) & 1) | static_cast<uint32_t>((rc.b00 >> x) & 1) << 1 |
           static_cast<uint32_t>((rc.a10 >> x) & 1) << 2 | static_cast<uint32_t>((rc.b10 >> x) & 1) << 3 |
           static_cast<uint32_t>((rc.a01 >> x) & 1) << 4 | static_cast<uint32_t>((rc.b01 >> x) & 1) << 5 |
           static_cast<uint32_t>((rc.a11 >> x) & 1) << 6 | static_cast<uint32_t>((rc.b11 >> x) & 1) << 7;
}

// Exclusive scan over the first NWS warps' values (threads >= 32*NWS pass 0 and only read the
// total): one warp-shuffle scan + one barrier.  `buf` is a 2-deep ping-pong so consecutive scans
// need no trailing barrier.
template <int NWS, typename T>
__device__ __forceinline__ T scan_front_warps(T value, T (*buf)[34], uint32_t& flip, T& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T* sums = buf[flip & 1u];
    flip ^= 1u;
    T incl = value;
    if (warp < NWS) {
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            T up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += up;
        }
        if (lane == 31) sums[warp] = incl;
    }
    __syncthreads();
    T before = 0, all = 0;
#pragma unroll
    for (int w = 0; w < NWS; ++w) {
        const T s = sums[w];
        if (w < warp) before += s;
        all += s;
    }
    total = all;
    return incl - value + before;
}

// One vertex of an active cell, from the shared-memory bricks.  (x,y,z) cell coords, slot_of(zi)
// maps a sample-layer index to its ring slot.  Arithmetic: SURVEY.md Appendix A.1.
template <class C, class SlotOf>
__device__ __forceinline__ void emit_regular_vertex(const Smem<C>& sm, SlotOf slot_of, int x, int y, int z,
                                                    uint32_t code, uint32_t transition_mask, hvx_vertex* dst) {
    const int c0 = code >> 4, c1 = code & 15;
    // sample-index coordinates of both edge endpoints (local + 1 for the halo)
    const int ax = x + 1 + (c0 & 1), ay = y + 1 + ((c0 >> 1) & 1), az = z + 1 + ((c0 >> 2) & 1);
    const int bx = x + 1 + (c1 & 1), by = y + 1 + ((c1 >> 1) & 1), bz = z + 1 + ((c1 >> 2) & 1);
    auto word = [&](int xi, int yi, int zi) -> uint32_t { return sm.ring[slot_of(zi)][yi * C::S + xi]; };
    auto dens = [&](int xi, int yi, int zi) -> float { return cw_density(word(xi, yi, zi)); };
    // central difference * 0.5, one-sided (no 0.5) on the +face where local == E
    auto grad = [&](int xi, int yi, int zi, float d, float g[3]) {
        g[0] = xi >= C::E + 1 ? fsub(d, dens(xi - 1, yi, zi)) : fmul(fsub(dens(xi + 1, yi, zi), dens(xi - 1, yi, zi)), 0.5f);
        g[1] = yi >= C::E + 1 ? fsub(d, dens(xi, yi - 1, zi)) : fmul(fsub(dens(xi, yi + 1, zi), dens(xi, yi - 1, zi)), 0.5f);
        g[2] = zi >= C::E + 1 ? fsub(d, dens(xi, yi, zi - 1)) : fmul(fsub(dens(xi, yi, zi + 1), dens(xi, yi, zi - 1)), 0.5f);
    };
    const uint32_t wa = word(ax, ay, az), wb = word(bx, by, bz);
    const float d0 = cw_density(wa), d1 = cw_density(wb);
    const float t = edge_parameter(d0, d1);
    float ga[3], gb[3];
    grad(ax, ay, az, d0, ga);
    grad(bx, by, bz, d1, gb);
    float p[3], n[3];
    p[0] = fmix(static_cast<float>(ax - 1), static_cast<float>(bx - 1), t);
    p[1] = fmix(static_cast<float>(ay - 1), static_cast<float>(by - 1), t);
    p[2] = fmix(static_cast<float>(az - 1), static_cast<float>(bz - 1), t);
    const float gx = fmix(ga[0], gb[0], t), gy = fmix(ga[1], gb[1], t), gz = fmix(ga[2], gb[2], t);
    const float s = fadd(fadd(fmul(gx, gx), fmul(gy, gy)), fmul(gz, gz));
    if (s > 1.0e-12f) {
        const float inv = fdiv(1.0f, fsqrt(s));
        n[0] = fmul(gx, inv);
        n[1] = fmul(gy, inv);
        n[2] = fmul(gz, inv);
    } else {
        n[0] = 0.0f;
        n[1] = 1.0f;
        n[2] = 0.0f;
    }
    // Transvoxel secondary position on faces that own a transition mesh
    // (PV/tests/gpu_transvoxel_emission.rs:346-400, thresholds 1 and E-1)
    if (transition_mask != 0) {
        const float hi = static_cast<float>(C::E - 1);
        uint32_t near = 0;
        near |= p[0] < 1.0f ? 1u : 0u;
        near |= p[0] > hi ? 2u : 0u;
        near |= p[1] < 1.0f ? 4u : 0u;
        near |= p[1] > hi ? 8u : 0u;
        near |= p[2] < 1.0f ? 16u : 0u;
        near |= p[2] > hi ? 32u : 0u;
        if (near != 0 && (near & ~transition_mask) == 0) {
            float off[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                off[a] = 0.0f;
                if (near & (1u << (2 * a))) off[a] = fmul(fsub(1.0f, p[a]), 0.25f);
                else if (near & (2u << (2 * a))) off[a] = fmul(fsub(hi, p[a]), 0.25f);
            }
            const float nc = fadd(fadd(fmul(off[0], n[0]), fmul(off[1], n[1])), fmul(off[2], n[2]));
#pragma unroll
            for (int a = 0; a < 3; ++a) p[a] = fsub(fadd(p[a], off[a]), fmul(n[a], nc));
        }
    }
    const uint32_t material = d0 <= 0.0f ? cw_material(wa) : cw_material(wb);
    float4* out = reinterpret_cast<float4*>(dst);
    out[0] = make_float4(p[0], p[1], p[2], __uint_as_float(material));
    out[1] = make_float4(n[0], n[1], n[2], __uint_as_float(0u));
}

template <class C>
__global__ void __launch_bounds__(C::NT, 1) regular_extract_kernel(const RegularParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem<C>& sm = *reinterpret_cast<Smem<C>*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int E = C::E, S = C::S, R = C::R, ZB = C::ZB, NT = C::NT;
    constexpr uint64_t ROWMASK = E == 64 ? ~0ull : 0xffffffffull;
    const size_t chunk_words = static_cast<size_t>(S) * S * S;

    // ---- one-time setup -------------------------------------------------------------
    for (int i = tid; i < 256; i += NT) sm.case_info[i] = HVX_REGULAR_CASE_INFO[i];
    for (int i = tid; i < 256 * 12; i += NT) sm.vertex_edge[i] = HVX_REGULAR_VERTEX_EDGE[i / 12][i % 12];
    for (int i = tid; i < 256; i += NT) sm.class_index[i] = HVX_REGULAR_CLASS_INDEX[i / 16][i % 16];
    for (int i = tid; i < R * C::BW; i += NT) sm.bits[i / C::BW][i % C::BW] = 0u;  // incl. the zero padding
    if (tid == 0) {
        for (int i = 0; i < R; ++i) mbar_init(&sm.full_bar[i], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // ---- producer state (thread 0 only) ---------------------------------------------
    // The CTA consumes a flat stream of sample layers: local chunk k contributes layers
    // seq = k*S .. k*S+S-1; slot = seq % R, barrier parity = (seq / R) & 1.
    uint32_t prod_seq = 0, prod_k = 0, prod_zi = 0, prod_slot = 0;
    bool prod_done = false;
    auto produce = [&](uint32_t oldest_needed_seq) {
        while (!prod_done && prod_seq < oldest_needed_seq + R) {
            if (prod_zi == 0) {
                uint32_t id = atomicAdd(p.work_counter, 1u);
                sm.chunk_ids[prod_k & 3] = id;
                if (id >= p.n_chunks) {
                    prod_done = true;
                    break;
                }
            }
            const uint32_t id = sm.chunk_ids[prod_k & 3];
            const uint32_t* src = p.samples + static_cast<size_t>(id) * chunk_words +
                                  static_cast<size_t>(prod_zi) * C::LAYER_WORDS;
            mbar_arrive_expect_tx(&sm.full_bar[prod_slot], C::LAYER_BYTES);
            bulk_g2s(&sm.ring[prod_slot][0], src, C::LAYER_BYTES, &sm.full_bar[prod_slot]);
            ++prod_seq;
            prod_slot = prod_slot + 1 == R ? 0 : prod_slot + 1;
            if (++prod_zi == S) {
                prod_zi = 0;
                ++prod_k;
            }
        }
    };

    // ---- consumer state (uniform across the CTA) -------------------------------------
    // (win_slot, win_q) = (seq % R, seq / R) of sample layer z0 of the current step, kept
    // incrementally so the hot loop has no division.
    uint32_t base_seq = 0;
    int win_slot = 0;
    uint32_t win_q = 0, flip32 = 0, flip64 = 0;
    auto advance = [&](int layers) {
        win_slot += layers;
        while (win_slot >= R) {
            win_slot -= R;
            ++win_q;
        }
    };
    for (uint32_t kc = 0;; ++kc, base_seq += S) {
        if (tid == 0) produce(base_seq);
        __syncthreads();
        const uint32_t chunk = sm.chunk_ids[kc & 3];
        if (chunk >= p.n_chunks) break;
        const ChunkDesc desc = p.descs[chunk];
        const uint64_t dirty = desc.dirty_microbricks;
        const uint32_t tmask = desc.transition_mask & 0x3fu;
        if (tid < 16) {
            // x mask of dirty microbricks for (my, mz) = (tid & 3, tid >> 2)
            uint64_t m = 0;
            for (int mx = 0; mx < 4; ++mx)
                if ((dirty >> (mx + 4 * tid)) & 1ull) m |= ((1ull << C::QW) - 1ull) << (mx * C::QW);
            sm.dirty_row[tid] = m;
        }
        hvx_vertex* const out_v = p.vertices + static_cast<size_t>(chunk) * p.max_vertices;
        uint32_t* const out_i = p.indices + static_cast<size_t>(chunk) * p.max_indices;
        const bool debug = p.cells != nullptr;
        const bool do_emit = p.mode == MODE_EXTRACT;

        uint32_t v_base = 0, i_base = 0;  // running chunk-local placement
        uint32_t active_cells = 0;        // classify counters
        int waited = 0;                   // sample layers [0, waited) of this chunk have landed

        for (int z0 = 0; z0 < E; z0 += ZB) {
            if (z0 != 0 && tid == 0) produce(base_seq + z0);
            // ring slot of sample layer z0 + d, d in [0, R)
            auto slot_of = [&](int d) -> int {
                const int s = win_slot + d;
                return s >= R ? s - R : s;
            };
            auto slot_of_layer = [&](int zi) -> int { return slot_of(zi - z0); };
            // ---- P0: wait for the window's layers ---------------------------------------
            const int need = min(S, z0 + C::WIN);
            for (int zi = waited; zi < need; ++zi) {
                const int s = win_slot + (zi - z0);
                const bool wrap = s >= R;
                mbar_wait(&sm.full_bar[wrap ? s - R : s], (win_q + (wrap ? 1u : 0u)) & 1u);
            }
            if (p.mode == MODE_STREAM_ONLY) {  // diagnostics: the bare HBM -> smem pipeline
                waited = need;
                __syncthreads();
                advance(ZB);
                continue;
            }
            // ---- P1: solid bits of the new corner layers, 32 samples per warp ballot -------
            {
                const int first = z0 == 0 ? 1 : z0 + 2, last = z0 + ZB + 1;  // inclusive, <= E+1
                constexpr int FULL = C::LAYER_WORDS / 32, TAIL = C::LAYER_WORDS % 32;
                for (int zi = first; zi <= last; ++zi) {
                    const int slot = slot_of(zi - z0);
                    // each lane reads the low (density) half of its sample: LDS.S16, ISETP, VOTE, STS
                    const short* src = reinterpret_cast<const short*>(&sm.ring[slot][0]) + 2 * (32 * warp + lane);
                    uint32_t* dst = sm.bits[slot] + warp;
#pragma unroll
                    for (int k = 0; k < (FULL + C::NW - 1) / C::NW; ++k) {
                        if (k * C::NW + C::NW <= FULL || warp < FULL - k * C::NW) {
                            const short d = src[64 * C::NW * k];
                            const uint32_t b = __ballot_sync(0xffffffffu, d <= 0);
                            if (lane == 0) dst[C::NW * k] = b;
                        }
                    }
                    if (TAIL != 0 && warp == C::NW - 1) {
                        const short* tail = reinterpret_cast<const short*>(&sm.ring[slot][0]) + 2 * (32 * FULL);
                        const short d = lane < TAIL ? tail[2 * lane] : short(1);
                        const uint32_t b = __ballot_sync(0xffffffffu, d <= 0);
                        if (lane == 0) sm.bits[slot][FULL] = b;
                    }
                }
            }
            waited = need;
            __syncthreads();
            // ---- P2: per-row active masks + ordered ranks --------------------------------
            uint32_t my_count = 0;
            if (tid < C::ROWS) {
                const int zl = tid / E, y = tid % E, z = z0 + zl;
                const RowCorners rc = load_row_corners<C>(sm, slot_of(zl + 1), slot_of(zl + 2), y);
                const uint64_t any = rc.a00 | rc.b00 | rc.a10 | rc.b10 | rc.a01 | rc.b01 | rc.a11 | rc.b11;
                const uint64_t all = rc.a00 & rc.b00 & rc.a10 & rc.b10 & rc.a01 & rc.b01 & rc.a11 & rc.b11;
                const uint64_t act = any & ~all & ROWMASK & sm.dirty_row[(y / C::QW) + 4 * (z / C::QW)];
                sm.active[tid] = act;
                my_count = __popcll(act);
            }
            // Most steps of most chunks see no surface at all: one barrier-with-OR decides, and only
            // steps with active cells pay for the rank scan.  (Debug records need every step.)
            if (!__syncthreads_or(my_count != 0) && !debug) {
                advance(ZB);
                continue;
            }
            uint32_t n_active;
            const uint32_t my_off = scan_front_warps<C::ROW_WARPS>(my_count, sm.scan32, flip32, n_active);
            if (tid < C::ROWS) sm.row_off[tid] = my_off;

            if (debug) {
                // ---- debug records: per-cell case words, block-relative offsets, scan blocks ----
                // (GpuTransvoxelCell / GpuTransvoxelCellOffset / GpuTransvoxelScanBlock)
                uint64_t row_tot = 0;
                RowCorners rc;
                uint64_t dirty_x = 0;
                int zl = 0, y = 0, z = 0;
                if (tid < C::ROWS) {
                    zl = tid / E;
                    y = tid % E;
                    z = z0 + zl;
                    rc = load_row_corners<C>(sm, slot_of(zl + 1), slot_of(zl + 2), y);
                    dirty_x = sm.dirty_row[(y / C::QW) + 4 * (z / C::QW)];
                    uint64_t act = sm.active[tid];
                    while (act) {
                        const int x = __ffsll(static_cast<long long>(act)) - 1;
                        act &= act - 1;
                        const uint32_t info = sm.case_info[case_at(rc, x)];
                        row_tot += (info & 15u) | (static_cast<uint64_t>(3u * ((info >> 4) & 15u)) << 32);
                    }
                }
                uint64_t step_tot;
                const uint64_t row_pref = scan_front_warps<C::ROW_WARPS>(row_tot, sm.scan64, flip64, step_tot);
                if (tid < C::ROWS) sm.row_pref64[tid] = row_pref;
                if (tid == 0) sm.row_pref64[C::ROWS] = step_tot;
                __syncthreads();
                if (tid < C::ROWS) {
                    const int block_row = (tid / C::RPB) * C::RPB;  // first row of this cell's 256-block
                    const uint64_t bp = sm.row_pref64[block_row];
                    uint32_t rv = static_cast<uint32_t>(row_pref) - static_cast<uint32_t>(bp);
                    uint32_t ri = static_cast<uint32_t>(row_pref >> 32) - static_cast<uint32_t>(bp >> 32);
                    const size_t cell0 = static_cast<size_t>(chunk) * E * E * E + static_cast<size_t>(z) * E * E +
                                         static_cast<size_t>(y) * E;
                    const uint32_t glo = static_cast<uint32_t>(desc.generation);
                    const uint32_t ghi = static_cast<uint32_t>(desc.generation >> 32);
                    for (int x = 0; x < E; ++x) {
                        if (!((dirty_x >> x) & 1ull)) continue;
                        const uint32_t c = case_at(rc, x);
                        const uint32_t info = sm.case_info[c];
                        const uint32_t nv = info & 15u, nt = (info >> 4) & 15u, cls = info >> 8;
                        uint4 rec = make_uint4(c | (cls << 8) | (nv << 16) | (nt << 24) | 0x80000000u, glo, ghi, 0u);
                        *reinterpret_cast<uint4*>(&p.cells[cell0 + x]) = rec;
                        uint4 off = make_uint4(rv, ri, glo, ghi);
                        *reinterpret_cast<uint4*>(&p.offsets[cell0 + x]) = off;
                        rv += nv;
                        ri += 3u * nt;
                    }
                    if (tid % C::RPB == 0) {
                        const uint64_t nx = sm.row_pref64[block_row + C::RPB];
                        hvx_scan_block blk;
                        blk.vertex_count = static_cast<uint32_t>(nx) - static_cast<uint32_t>(bp);
                        blk.index_count = static_cast<uint32_t>(nx >> 32) - static_cast<uint32_t>(bp >> 32);
                        blk.first_vertex = v_base + static_cast<uint32_t>(bp);
                        blk.first_index = i_base + static_cast<uint32_t>(bp >> 32);
                        const size_t b = static_cast<size_t>(chunk) * (E * E * E / 256) +
                                         (static_cast<size_t>(z) * E * E + static_cast<size_t>(y) * E) / 256;
                        p.blocks[b] = blk;
                    }
                }
            }

            if (n_active == 0) {  // uniform: nothing on the surface in these layers
                advance(ZB);
                continue;
            }
            active_cells += n_active;

            // ---- P3: emission in ordered batches of CB active cells -----------------------
            for (uint32_t b0 = 0; b0 < n_active; b0 += C::CB) {
                const uint32_t nb = min(static_cast<uint32_t>(C::CB), n_active - b0);
                __syncthreads();  // row_off / previous batch's cell_rec, owner are free
                // P3a: rank scatter, one thread per quarter row
                for (int q = tid; q < C::ROWS * 4; q += NT) {
                    const int r = q >> 2, part = q & 3;
                    const uint64_t m = sm.active[r];
                    uint32_t sub = static_cast<uint32_t>((m >> (C::QW * part)) & ((1ull << C::QW) - 1ull));
                    if (!sub) continue;
                    uint32_t rank = sm.row_off[r] + __popcll(m & ((1ull << (C::QW * part)) - 1ull));
                    while (sub) {
                        const int bit = __ffs(sub) - 1;
                        sub &= sub - 1;
                        const uint32_t rel = rank - b0;  // wraps for rank < b0
                        if (rel < nb) sm.cell_rec[rel] = static_cast<uint32_t>(C::QW * part + bit) | (r << 8);
                        ++rank;
                    }
                }
                __syncthreads();
                // P3b: case lookup, per-cell counts, ordered offsets, index emission
                uint32_t packed = 0, rec = 0, info = 0;
                if (tid < nb) {
                    rec = sm.cell_rec[tid];
                    const int x = rec & 63, r = rec >> 8, zl = r / E, y = r % E;
                    const RowCorners rc = load_row_corners<C>(sm, slot_of(zl + 1), slot_of(zl + 2), y);
                    const uint32_t c = case_at(rc, x);
                    info = sm.case_info[c];
                    rec |= c << 16;
                    packed = (info & 15u) | ((3u * ((info >> 4) & 15u)) << 16);
                }
                uint32_t batch_tot;
                const uint32_t off = scan_front_warps<C::NW>(packed, sm.scan32, flip32, batch_tot);
                const uint32_t batch_v = batch_tot & 0xffffu, batch_i = batch_tot >> 16;
                if (tid < nb && do_emit) {
                    const uint32_t nv = info & 15u, ni = 3u * ((info >> 4) & 15u), cls = info >> 8;
                    const uint32_t vo = off & 0xffffu, io = off >> 16;
                    sm.cell_rec[tid] = rec;
                    for (uint32_t k = 0; k < nv; ++k) sm.owner[vo + k] = static_cast<uint16_t>(tid | (k << 10));
                    const uint32_t first_vertex = v_base + vo;
                    const uint32_t dst = i_base + io;
                    const uint8_t* tri = &sm.class_index[cls * 16];
                    for (uint32_t j = 0; j < ni; ++j)
                        if (dst + j < p.max_indices) out_i[dst + j] = first_vertex + tri[j];
                }
                __syncthreads();
                // P3c: one thread per vertex
                if (do_emit) {
                    for (uint32_t v = tid; v < batch_v; v += NT) {
                        const uint32_t o = sm.owner[v];
                        const uint32_t slot = o & 1023u, k = o >> 10;
                        const uint32_t cr = sm.cell_rec[slot];
                        const int x = cr & 63, r = (cr >> 8) & 255, c = cr >> 16;
                        const int zl = r / E, y = r % E, z = z0 + zl;
                        const uint32_t code = sm.vertex_edge[c * 12 + k];
                        if (v_base + v < p.max_vertices)
                            emit_regular_vertex<C>(sm, slot_of_layer, x, y, z, code, tmask, out_v + v_base + v);
                    }
                }
                v_base += batch_v;
                i_base += batch_i;
            }
            __syncthreads();  // ring reads of this step are done before the producer refills
            advance(ZB);
        }
        advance(S - E);  // the chunk's last two sample layers

        // ---- chunk epilogue: one counter record, no atomics --------------------------------
        if (tid == 0) {
            const uint32_t vo = v_base > p.max_vertices ? 1u : 0u, io = i_base > p.max_indices ? 1u : 0u;
            const bool ok = !(vo | io) && do_emit;
            hvx_emission_counters ec;
            ec.required_vertices = v_base;
            ec.required_indices = i_base;
            ec.emitted_vertices = ok ? v_base : 0u;
            ec.emitted_indices = ok ? i_base : 0u;
            ec.vertex_overflow = vo;
            ec.index_overflow = io;
            ec.completed = 1u;
            ec._pad = 0u;
            p.counters[chunk] = ec;
            hvx_classify_counters cc;
            cc.visited_cells = static_cast<uint32_t>(__popcll(dirty)) * (C::QW * C::QW * C::QW);
            cc.active_cells = active_cells;
            cc.vertices = v_base;
            cc.triangles = i_base / 3u;
            p.classify[chunk] = cc;
            hvx_range rg;
            rg.first_vertex = chunk * p.max_vertices;
            rg.vertex_count = ok ? v_base : 0u;
            rg.first_index = chunk * p.max_indices;
            rg.index_count = ok ? i_base : 0u;
            p.ranges[chunk] = rg;
        }
    }
}

template <class C>
cudaError_t launch_cfg(const RegularParams& p, const DeviceInfo& dev, cudaStream_t stream) {
    const size_t smem = sizeof(Smem<C>);
    cudaError_t err = cudaFuncSetAttribute(regular_extract_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem));
    if (err != cudaSuccess) return err;
    int ctas_per_sm = 1;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, regular_extract_kernel<C>, C::NT, smem);
    if (err != cudaSuccess) return err;
    if (ctas_per_sm < 1) return cudaErrorInvalidConfiguration;
    const uint32_t grid = static_cast<uint32_t>(
        min(static_cast<long long>(p.n_chunks), static_cast<long long>(dev.sm_count) * ctas_per_sm));
    regular_extract_kernel<C><<<grid, C::NT, smem, stream>>>(p);
    return cudaGetLastError();
}

using Cfg64 = Cfg<64, 2, 11, 512>;
using Cfg32 = Cfg<32, 4, 16, 256>;

// Tuning variants (HVX_REGULAR_VARIANT=<n>, default 0); all produce identical output.
int variant_from_env() {
    const char* v = getenv("HVX_REGULAR_VARIANT");
    return v ? atoi(v) : 0;
}

}  // namespace

size_t regular_smem_bytes(int edge) { return edge == 64 ? sizeof(Smem<Cfg64>) : sizeof(Smem<Cfg32>); }

cudaError_t launch_regular(int edge, const RegularParams& p, const DeviceInfo& dev, cudaStream_t stream) {
    if (p.n_chunks == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(p.work_counter, 0, sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    const int variant = variant_from_env();
    if (edge == 64) {
        switch (variant) {
            case 1: return launch_cfg<Cfg<64, 4, 11, 512>>(p, dev, stream);
            case 2: return launch_cfg<Cfg<64, 2, 11, 256>>(p, dev, stream);
            case 3: return launch_cfg<Cfg<64, 4, 11, 256>>(p, dev, stream);
            case 4: return launch_cfg<Cfg<64, 1, 10, 256>>(p, dev, stream);
            case 5: return launch_cfg<Cfg<64, 2, 9, 512>>(p, dev, stream);
            default: return launch_cfg<Cfg64>(p, dev, stream);
        }
    }
    if (edge == 32) {
        switch (variant) {
            case 1: return launch_cfg<Cfg<32, 8, 20, 256>>(p, dev, stream);
            case 2: return launch_cfg<Cfg<32, 4, 16, 128>>(p, dev, stream);
            case 3: return launch_cfg<Cfg<32, 2, 12, 128>>(p, dev, stream);
            default: return launch_cfg<Cfg32>(p, dev, stream);
        }
    }
    return cudaErrorInvalidValue;
}

}  // namespace hvx
