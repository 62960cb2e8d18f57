// regular_extract.cu -- fused regular-cell Transvoxel extraction for sm_100a.
//
// Replaces the reference's four dispatches per page
//   classify_regular_cells  PV/src/transvoxel_classify.wgsl:75-120
//   scan_regular_cells      PV/src/transvoxel_emit.wgsl:131-172
//   scan_regular_blocks     PV/src/transvoxel_emit.wgsl:174-202
//   emit_regular_cells      PV/src/transvoxel_emit.wgsl:291-370
// with ONE persistent, warp-specialised kernel over a batch of chunks.  Arithmetic follows the
// reference's CPU extractor (PV/tests/gpu_transvoxel_emission.rs:272-450), not the WGSL, so
// positions and normals are bit-identical to the oracle.
//
// Design:
//   * one CTA per SM pulls chunks from an atomic queue and walks each chunk front to back in z.
//     x-fastest cell order == walk order, so vertex/index placement is a running prefix inside
//     the CTA: order preserving by construction, no cross-CTA scan on the data path.
//   * a PRODUCER warp streams SLABS of two sample layers (2*(E+2)^2 words, contiguous, 16-byte
//     aligned) from HBM into a shared-memory ring with cp.async.bulk (TMA 1-D, SASS UBLKCP):
//     full[slot] mbarriers carry the byte count, empty[slot] mbarriers hand slots back.  It runs
//     ahead across chunk boundaries.  Every sample is read from HBM exactly once.
//   * CONSUMER warps never meet at a CTA barrier while the chunk is empty.  Iteration j (slab j):
//     every warp turns its 1/NW of the slab into solid bits (all loads first, then one ballot per
//     32 samples); a rotating group of warps classifies step j-1 (two cell layers) -- a cell row's
//     8 corner signs are 8 shifted copies of 4 bit rows, so one thread gets the 64-cell active
//     mask of a row with funnel shifts and AND/OR -- and publishes a verdict through an mbarrier;
//     everybody reads the verdict of step j-2 (ready for a whole iteration) and releases the
//     oldest ring slot.
//   * steps that do contain surface cells are batched (up to EBS consecutive steps) and emitted
//     collectively: quarter-row popc ranks -> warp-shuffle scan -> ordered compaction -> one
//     thread per active CELL (case from the bricks, counts, second scan, indices) -> one thread
//     per VERTEX reading its 14 samples from the shared-memory bricks (no HBM re-read).
//   * no single-address atomics: the reference's 5 atomicAdd per cell become one counter record
//     written once per chunk.
#include <cstdlib>

#include "hvx_device.cuh"
#include "hvx_kernels.h"

namespace hvx {

namespace {

// Lengyel's tables live in device global memory (statically initialised); every CTA copies the
// 3.8 KB it needs into shared memory once (the kernel is persistent).
#define HVX_TABLE static __device__ const
#include "transvoxel_tables.inc"

template <int E_, int EBS_, int RS_, int NW_>
struct Cfg {
    static constexpr int E = E_;          // cells per chunk edge
    static constexpr int S = E_ + 2;      // samples per edge (1-sample halo)
    static constexpr int G = 2;           // sample layers per slab (one TMA copy, one hand-off)
    static constexpr int NSLAB = S / G;   // slabs per chunk
    static constexpr int EBS = EBS_;      // max steps (pairs of cell layers) per emission batch
    static constexpr int RS = RS_;        // ring slots, in slabs
    static constexpr int NW = NW_;        // consumer warps
    static constexpr int NT = NW_ * 32;   // consumer threads
    static constexpr int NT_ALL = NT + 32;  // + producer warp
    static constexpr int LAYER_WORDS = S * S;
    static constexpr int SLAB_WORDS = G * LAYER_WORDS;
    static constexpr int SLAB_BYTES = SLAB_WORDS * 4;
    static constexpr int FULL = SLAB_WORDS / 32;        // whole 32-sample ballot blocks per slab
    static constexpr int TAIL = SLAB_WORDS % 32;
    static constexpr int BPW = (FULL + NW_ - 1) / NW_;  // ballot blocks per warp per slab (last one guarded)
    static constexpr int BW = FULL + 1 + 3;             // bit words per slab + zero padding
    static constexpr int BR = 4;                        // slab-bits / verdict / active-mask ring depth
    static constexpr int PWS = (G * E_) / 32;           // warps classifying one step (G cell layers)
    static constexpr int PG = NW_ / PWS;                // classification groups (left-over warps never classify)
    static constexpr int STEP_ROWS = G * E_;            // cell rows per step
    static constexpr int ROWS = EBS_ * STEP_ROWS;       // cell rows per emission batch
    static constexpr int CB = NT < 512 ? NT : 512;      // active cells per emission sub-batch (one thread per cell)
    static constexpr int QW = E_ / 4;                   // microbrick edge == quarter-row width
    static constexpr int RPB = 256 / E_;                // rows per 256-cell scan block
    static_assert(S % G == 0, "chunk must split evenly into slabs");
    static_assert(RS_ >= EBS_ + 3 + 1, "ring must hold an emission window, the slab being classified and one in flight");
    static_assert(RS_ < NSLAB, "producer may be at most one chunk ahead");
    static_assert(EBS_ + 2 <= BR && EBS_ <= 2, "bit / verdict rings too small");
    static_assert(SLAB_BYTES % 16 == 0, "cp.async.bulk needs 16-byte multiples");
    static_assert(ROWS * 2 <= NT && ROWS <= 256 && NW_ >= PWS && (E_ == 32 || E_ == 64), "unsupported tiling");
};

template <class C>
struct Smem {
    alignas(128) uint32_t ring[C::RS][C::SLAB_WORDS];
    alignas(8) uint64_t full_bar[C::RS];
    uint64_t empty_bar[C::RS];
    uint64_t bits_bar[C::BR];
    uint64_t verdict_bar[C::BR];
    uint64_t active[C::BR][C::STEP_ROWS];  // active-cell bits per cell row of a step
    union {
        uint64_t row_pref[C::ROWS + 1];    // debug records: per-row exclusive (vertices | indices<<32)
        uint8_t owner[C::CB * 12];         // emission: vertex -> (cell slot of the sub-batch) >> 1
    };
    uint32_t layer_off[8];                 // ring word offset of sample layer z0 + d of the batch being emitted
    uint64_t scan64[2][34];
    uint32_t scan32[2][34];
    uint32_t bits[C::BR][C::BW];           // solid bit of every sample of a slab, flat x-fastest order
    uint32_t verdict_flag[C::BR][C::PWS];
    uint32_t cell_rec[C::CB];              // emission sub-batch: x | row<<8 | case<<16
    uint16_t cell_vo[C::CB];               // sub-batch-relative first vertex of each cell
    uint32_t chunk_ids[4];
    uint16_t case_info[256];
    uint8_t vertex_edge[256 * 12];
    uint8_t class_index[16 * 16];
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// barrier among the consumer warps only (the producer warp never joins)
template <int NT>
__device__ __forceinline__ void consumer_sync() {
    asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
}

// Solid bits of the 8 corners of every cell of one cell row: bit x of a** is the corner at
// sample x+1 (cell-local x), bit x of b** the corner at x+2; 10 = next sample row, 01 = next layer.
struct RowCorners {
    uint64_t a00, b00, a10, b10, a01, b01, a11, b11;
};

// Bits [o+1, o+1+E) and [o+2, o+2+E) of a flat solid-bit array, o = base + row * S.
template <class C>
__device__ __forceinline__ void row_window(const uint32_t* __restrict__ bits, int base, int row, uint64_t& a, uint64_t& b) {
    const int p = base + row * C::S + 1, w = p >> 5, sh = p & 31;
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

// bits0/base0: slab bit array and bit offset of sample layer z+1; bits1/base1: of sample layer z+2
template <class C>
__device__ __forceinline__ RowCorners load_row_corners(const uint32_t* __restrict__ bits0, int base0,
                                                       const uint32_t* __restrict__ bits1, int base1, int y) {
    RowCorners rc;
    row_window<C>(bits0, base0, y + 1, rc.a00, rc.b00);
    row_window<C>(bits0, base0, y + 2, rc.a10, rc.b10);
    row_window<C>(bits1, base1, y + 1, rc.a01, rc.b01);
    row_window<C>(bits1, base1, y + 2, rc.a11, rc.b11);
    return rc;
}

// Exclusive scan over the values of the first `nws` consumer warps (other threads pass 0 and only
// read the total): one warp-shuffle scan + one consumer barrier.  `buf` is a 2-deep ping-pong so
// consecutive scans need no trailing barrier.
template <int NT, typename T>
__device__ __forceinline__ T scan_front_warps(T value, int nws, T (*buf)[34], uint32_t& flip, T& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T* sums = buf[flip & 1u];
    flip ^= 1u;
    T incl = value;
    if (warp < nws) {
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            T up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += up;
        }
        if (lane == 31) sums[warp] = incl;
    }
    consumer_sync<NT>();
    T before = 0, all = 0;
    if constexpr (sizeof(T) == 4) {
        // second level in one shot: lane w holds warp w's sum, two REDUX.ADD give prefix and total
        const uint32_t mine = lane < nws ? sums[lane] : 0u;
        before = __reduce_add_sync(0xffffffffu, lane < warp ? mine : 0u);
        all = __reduce_add_sync(0xffffffffu, mine);
    } else {
        for (int w = 0; w < nws; ++w) {
            const T s = sums[w];
            if (w < warp) before += s;
            all += s;
        }
    }
    total = all;
    return incl - value + before;
}

// One vertex of an active cell, from the shared-memory bricks.  (x,y) cell coords, zl the cell
// layer relative to the batch's first layer, z its absolute layer; layer_words(d) is the word
// offset inside the ring of sample layer (first layer + d).  Arithmetic: SURVEY.md Appendix A.1.
template <class C, class LayerOff>
__device__ __forceinline__ void emit_regular_vertex(const uint32_t* __restrict__ ring, LayerOff layer_words, int x, int y,
                                                    int zl, int z, uint32_t code, uint32_t transition_mask,
                                                    hvx_vertex* dst) {
    constexpr int S = C::S;
    const int c0 = code >> 4, c1 = code & 15;
    // sample-index coordinates of both edge endpoints (local + 1 for the halo)
    const int ax = x + 1 + (c0 & 1), ay = y + 1 + ((c0 >> 1) & 1), adz = 1 + ((c0 >> 2) & 1);
    const int bx = x + 1 + (c1 & 1), by = y + 1 + ((c1 >> 1) & 1), bdz = 1 + ((c1 >> 2) & 1);
    const int az = z + adz, bz = z + bdz;  // absolute sample-layer indices
    // one endpoint: its word, the six neighbours, central difference * 0.5 (one-sided, no 0.5, on
    // the +face where local == E, i.e. sample index E+1)
    auto endpoint = [&](int xi, int yi, int dz, int zi, uint32_t& w, float& d, float g[3]) {
        const int off = yi * S + xi;
        const uint32_t* l0 = ring + layer_words(zl + dz) + off;
        const uint32_t* lm = ring + layer_words(zl + dz - 1) + off;
        w = l0[0];
        d = cw_density(w);
        const float xm = cw_density(l0[-1]), ym = cw_density(l0[-S]), zm = cw_density(lm[0]);
        if (xi >= C::E + 1) g[0] = fsub(d, xm); else g[0] = fmul(fsub(cw_density(l0[1]), xm), 0.5f);
        if (yi >= C::E + 1) g[1] = fsub(d, ym); else g[1] = fmul(fsub(cw_density(l0[S]), ym), 0.5f);
        if (zi >= C::E + 1) g[2] = fsub(d, zm);
        else g[2] = fmul(fsub(cw_density((ring + layer_words(zl + dz + 1) + off)[0]), zm), 0.5f);
    };
    uint32_t wa, wb;
    float d0, d1, ga[3], gb[3];
    endpoint(ax, ay, adz, az, wa, d0, ga);
    endpoint(bx, by, bdz, bz, wb, d1, gb);
    const float t = edge_parameter(d0, d1);
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
__global__ void __launch_bounds__(C::NT_ALL, C::E == 32 ? 2 : 1) regular_extract_kernel(const RegularParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem<C>& sm = *reinterpret_cast<Smem<C>*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int E = C::E, S = C::S, RS = C::RS, NT = C::NT, NW = C::NW, LW = C::LAYER_WORDS;
    constexpr uint64_t ROWMASK = E == 64 ? ~0ull : 0xffffffffull;
    const size_t chunk_words = static_cast<size_t>(S) * S * S;

    // ---- one-time setup (all warps) ---------------------------------------------------------
    for (int i = tid; i < 256; i += C::NT_ALL) sm.case_info[i] = HVX_REGULAR_CASE_INFO[i];
    for (int i = tid; i < 256 * 12; i += C::NT_ALL) sm.vertex_edge[i] = HVX_REGULAR_VERTEX_EDGE[i / 12][i % 12];
    for (int i = tid; i < 256; i += C::NT_ALL) sm.class_index[i] = HVX_REGULAR_CLASS_INDEX[i / 16][i % 16];
    for (int i = tid; i < C::BR * C::BW; i += C::NT_ALL) sm.bits[i / C::BW][i % C::BW] = 0u;  // incl. zero padding
    if (tid == 0) {
        for (int i = 0; i < RS; ++i) {
            mbar_init(&sm.full_bar[i], 1);
            mbar_init(&sm.empty_bar[i], NW);
        }
        for (int i = 0; i < C::BR; ++i) {
            mbar_init(&sm.bits_bar[i], NW);
            mbar_init(&sm.verdict_bar[i], C::PWS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    // =========================================================================================
    // PRODUCER warp: chunk queue + TMA bulk copies, one slab (G sample layers, contiguous in HBM
    // and in the ring) per copy.  The n-th use of a ring slot waits for its (n-1)-th release.
    // =========================================================================================
    if (warp == NW) {
        if (lane == 0) {
            int slot = 0;
            uint32_t round = 0;  // how many times the ring has wrapped
            for (uint32_t k = 0;; ++k) {
                const uint32_t id = atomicAdd(p.work_counter, 1u);
                // chunk_ids[k & 3] was last read for local chunk k - 4, long retired (RS < NSLAB)
                sm.chunk_ids[k & 3] = id;
                if (id >= p.n_chunks) {
                    // sentinel: complete the slot's phase without data so the consumers wake up and exit
                    mbar_wait(&sm.empty_bar[slot], (round & 1u) ^ 1u);
                    mbar_arrive(&sm.full_bar[slot]);
                    break;
                }
                const uint32_t* src = p.samples + static_cast<size_t>(id) * chunk_words;
                for (int j = 0; j < C::NSLAB; ++j) {
                    mbar_wait(&sm.empty_bar[slot], (round & 1u) ^ 1u);
                    mbar_arrive_expect_tx(&sm.full_bar[slot], C::SLAB_BYTES);
                    bulk_g2s(&sm.ring[slot][0], src + static_cast<size_t>(j) * C::SLAB_WORDS, C::SLAB_BYTES,
                             &sm.full_bar[slot]);
                    if (++slot == RS) {
                        slot = 0;
                        ++round;
                    }
                }
            }
        }
        return;
    }

    // =========================================================================================
    // CONSUMER warps.  Iteration j handles slab j (sample layers 2j, 2j+1) of the chunk:
    //   P1  every warp turns its 1/NW of the slab into solid bits;
    //   P2  one group of PWS warps classifies step j = cell layers 2j-2, 2j-1 (needs bits of
    //       sample layers 2j-1 .. 2j+1) and publishes active masks + a verdict;
    //   P3  everybody consumes the verdict of step j-1 (published one iteration ago, so the wait
    //       is normally free); its emission window is slabs j-2 .. j, all resident by now.
    // =========================================================================================
    const uint32_t* const ring_flat = &sm.ring[0][0];
    int slot = 0;            // ring slot of the slab being consumed
    uint32_t round = 0;      // ring wraps seen by the consume pointer
    int rel_slot = 0;        // ring slot of the oldest slab not yet handed back
    uint32_t sc = 0;         // running slab counter (bits ring index sc & 3)
    uint32_t stc = 0;        // running step counter (verdict / active ring index stc & 3)
    uint32_t flip32 = 0, flip64 = 0;
    const int p2_group = warp / C::PWS, p2_sub = warp % C::PWS;

    for (uint32_t kc = 0;; ++kc) {
        uint32_t chunk = 0;
        ChunkDesc desc{};
        uint64_t dirty = 0;
        uint32_t tmask = 0;
        hvx_vertex* out_v = nullptr;
        uint32_t* out_i = nullptr;
        const bool debug = p.cells != nullptr;
        const bool do_emit = p.mode == MODE_EXTRACT;
        uint32_t v_base = 0, i_base = 0, active_cells = 0;  // running chunk-local placement / counters
        int released = 0;                    // slabs of this chunk handed back to the producer
        int pend_first = 0, pend_count = 0;  // steps waiting for a collective emission
        const uint32_t sc0 = sc;             // slab counter of slab 0 of this chunk
        const uint32_t stc0 = stc;           // step counter of step 1 of this chunk

        auto release_through = [&](int last_slab) {  // hand back slabs [released, last_slab]
            while (released <= last_slab) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.empty_bar[rel_slot]);
                if (++rel_slot == RS) rel_slot = 0;
                ++released;
            }
        };
        // ---- collective emission of steps [pend_first, pend_first + pend_count) -----------------
        auto flush = [&]() {
            const int z0 = 2 * pend_first - 2;             // first cell layer of the batch (even)
            const int rows = pend_count * C::STEP_ROWS;    // row r: cell layer z0 + r / E, y = r % E
            // sample layer z0 + d lives in slab (released + d/2) -> ring slot rel_slot + d/2; the word
            // offsets go through a tiny smem table (published by the rank scan's barrier below)
            if (tid < 8) {
                int s2 = rel_slot + (tid >> 1);
                if (s2 >= RS) s2 -= RS;
                sm.layer_off[tid] = static_cast<uint32_t>(s2 * C::SLAB_WORDS + (tid & 1) * LW);
            }
            auto layer_words = [&](int d) -> int { return static_cast<int>(sm.layer_off[d]); };
            // Transvoxel case of one cell straight from the bricks: bit i = corner (i&1, i>>1&1, i>>2&1)
            auto case_of = [&](int x, int y, int zl) -> uint32_t {
                const uint32_t* l0 = ring_flat + layer_words(zl + 1) + (y + 1) * S + (x + 1);
                const uint32_t* l1 = ring_flat + layer_words(zl + 2) + (y + 1) * S + (x + 1);
                return (cw_solid(l0[0]) ? 1u : 0u) | (cw_solid(l0[1]) ? 2u : 0u) | (cw_solid(l0[S]) ? 4u : 0u) |
                       (cw_solid(l0[S + 1]) ? 8u : 0u) | (cw_solid(l1[0]) ? 16u : 0u) | (cw_solid(l1[1]) ? 32u : 0u) |
                       (cw_solid(l1[S]) ? 64u : 0u) | (cw_solid(l1[S + 1]) ? 128u : 0u);
            };
            auto active_row = [&](int r) -> uint64_t {
                return sm.active[(stc0 + static_cast<uint32_t>(pend_first - 1 + r / C::STEP_ROWS)) & (C::BR - 1)][r % C::STEP_ROWS];
            };
            // ---- A: ranks.  Each thread owns two adjacent quarter rows (QW cells each); one scan over
            //      their popcounts gives every quarter its first rank in x-fastest cell order. ----------
            const int q0 = 2 * tid;                        // quarter rows q0, q0 + 1 of row q0 >> 2
            uint32_t sub0 = 0, sub1 = 0;
            if (q0 < rows * 4) {
                const uint64_t m = active_row(q0 >> 2);
                const int part = q0 & 3;                   // 0 or 2
                sub0 = static_cast<uint32_t>((m >> (C::QW * part)) & ((1ull << C::QW) - 1ull));
                sub1 = static_cast<uint32_t>((m >> (C::QW * (part + 1))) & ((1ull << C::QW) - 1ull));
            }
            const uint32_t c0 = __popc(sub0), c1 = __popc(sub1);
            uint32_t n_active;
            const uint32_t excl = scan_front_warps<NT>(c0 + c1, (rows * 2 + 31) / 32, sm.scan32, flip32, n_active);

            if (debug) {
                // ---- debug records: per-cell case words, block-relative offsets, scan blocks ----
                // (GpuTransvoxelCell / GpuTransvoxelCellOffset / GpuTransvoxelScanBlock); one thread per row
                uint64_t row_tot = 0, dirty_x = 0;
                const int my_zl = tid / E, my_y = tid % E, z = z0 + my_zl;
                if (tid < rows) {
                    const uint32_t nib = static_cast<uint32_t>(dirty >> (4 * ((my_y / C::QW) + 4 * (z / C::QW)))) & 15u;
                    for (int mx = 0; mx < 4; ++mx)
                        if ((nib >> mx) & 1u) dirty_x |= ((1ull << C::QW) - 1ull) << (mx * C::QW);
                    for (uint64_t m = active_row(tid); m != 0; m &= m - 1) {
                        const uint32_t info = sm.case_info[case_of(__ffsll(static_cast<long long>(m)) - 1, my_y, my_zl)];
                        row_tot += (info & 15u) | (static_cast<uint64_t>(3u * ((info >> 4) & 15u)) << 32);
                    }
                }
                uint64_t step_tot;
                const uint64_t row_pref = scan_front_warps<NT>(row_tot, rows / 32, sm.scan64, flip64, step_tot);
                if (tid < rows) sm.row_pref[tid] = row_pref;
                if (tid == 0) sm.row_pref[rows] = step_tot;
                consumer_sync<NT>();
                if (tid < rows) {
                    const int block_row = (tid / C::RPB) * C::RPB;  // first row of this cell's 256-block
                    const uint64_t bp = sm.row_pref[block_row];
                    uint32_t rv = static_cast<uint32_t>(row_pref) - static_cast<uint32_t>(bp);
                    uint32_t ri = static_cast<uint32_t>(row_pref >> 32) - static_cast<uint32_t>(bp >> 32);
                    const size_t cell0 = static_cast<size_t>(chunk) * E * E * E + static_cast<size_t>(z) * E * E +
                                         static_cast<size_t>(my_y) * E;
                    const uint32_t glo = static_cast<uint32_t>(desc.generation);
                    const uint32_t ghi = static_cast<uint32_t>(desc.generation >> 32);
                    for (int x = 0; x < E; ++x) {
                        if (!((dirty_x >> x) & 1ull)) continue;
                        const uint32_t c = case_of(x, my_y, my_zl);
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
                        const uint64_t nx = sm.row_pref[block_row + C::RPB];
                        hvx_scan_block blk;
                        blk.vertex_count = static_cast<uint32_t>(nx) - static_cast<uint32_t>(bp);
                        blk.index_count = static_cast<uint32_t>(nx >> 32) - static_cast<uint32_t>(bp >> 32);
                        blk.first_vertex = v_base + static_cast<uint32_t>(bp);
                        blk.first_index = i_base + static_cast<uint32_t>(bp >> 32);
                        const size_t b = static_cast<size_t>(chunk) * (E * E * E / 256) +
                                         (static_cast<size_t>(z) * E * E + static_cast<size_t>(my_y) * E) / 256;
                        p.blocks[b] = blk;
                    }
                }
            }

            active_cells += n_active;
            for (uint32_t b0 = 0; b0 < n_active; b0 += C::CB) {
                const uint32_t nb = min(static_cast<uint32_t>(C::CB), n_active - b0);
                if (b0 != 0) consumer_sync<NT>();  // the previous sub-batch's records are consumed
                // ---- B: ordered compaction: scatter (x, row) of the cells whose rank is in this sub-batch
                {
                    uint32_t rank = excl;
                    const int r = q0 >> 2, xbase = C::QW * (q0 & 3);
                    for (uint32_t s = sub0; s != 0; s &= s - 1, ++rank)
                        if (rank - b0 < nb) sm.cell_rec[rank - b0] = static_cast<uint32_t>(xbase + __ffs(s) - 1) | (r << 8);
                    rank = excl + c0;
                    for (uint32_t s = sub1; s != 0; s &= s - 1, ++rank)
                        if (rank - b0 < nb) sm.cell_rec[rank - b0] = static_cast<uint32_t>(xbase + C::QW + __ffs(s) - 1) | (r << 8);
                }
                consumer_sync<NT>();
                // ---- C: one thread per active cell: case, counts, ordered offsets, indices -----------
                uint32_t packed = 0, rec = 0, info = 0;
                if (tid < nb) {
                    rec = sm.cell_rec[tid];
                    const int x = rec & 63, r = rec >> 8;
                    const uint32_t c = case_of(x, r % E, r / E);
                    info = sm.case_info[c];
                    rec |= c << 16;
                    packed = (info & 15u) | ((3u * ((info >> 4) & 15u)) << 16);
                }
                uint32_t batch_tot;
                const uint32_t off = scan_front_warps<NT>(packed, static_cast<int>((nb + 31u) / 32u), sm.scan32, flip32, batch_tot);
                const uint32_t batch_v = batch_tot & 0xffffu, batch_i = batch_tot >> 16;
                if (tid < nb) {
                    const uint32_t vo = off & 0xffffu, io = off >> 16;
                    sm.cell_rec[tid] = rec;
                    sm.cell_vo[tid] = static_cast<uint16_t>(vo);
                    if (do_emit) {
                        const uint32_t nv = info & 15u, ni = 3u * ((info >> 4) & 15u), cls = info >> 8;
                        for (uint32_t k = 0; k < nv; ++k) sm.owner[vo + k] = static_cast<uint8_t>(tid >> 1);
                        const uint32_t first_vertex = v_base + vo, dst = i_base + io;
                        // the class's <= 15 local indices are one aligned 16-byte row
                        const uint4 row = *reinterpret_cast<const uint4*>(&sm.class_index[cls * 16]);
                        const uint32_t words[4] = {row.x, row.y, row.z, row.w};
                        if (dst + ni <= p.max_indices) {
#pragma unroll
                            for (uint32_t j = 0; j < 15; ++j)
                                if (j < ni) out_i[dst + j] = first_vertex + ((words[j >> 2] >> (8 * (j & 3))) & 0xffu);
                        } else {
                            for (uint32_t j = 0; j < ni; ++j)
                                if (dst + j < p.max_indices) out_i[dst + j] = first_vertex + ((words[j >> 2] >> (8 * (j & 3))) & 0xffu);
                        }
                    }
                }
                consumer_sync<NT>();
                // ---- D: one thread per vertex; the owner map names a pair of cells, one compare picks
                if (do_emit) {
                    for (uint32_t v = tid; v < batch_v; v += NT) {
                        uint32_t lo = 2u * sm.owner[v];
                        if (lo + 1 < nb && sm.cell_vo[lo + 1] <= v) ++lo;
                        const uint32_t cr = sm.cell_rec[lo], k = v - sm.cell_vo[lo];
                        const int x = cr & 63, r = (cr >> 8) & 255, c = cr >> 16;
                        const int zl = r / E, y = r % E;
                        const uint32_t code = sm.vertex_edge[c * 12 + k];
                        if (v_base + v < p.max_vertices)
                            emit_regular_vertex<C>(ring_flat, layer_words, x, y, zl, z0 + zl, code, tmask, out_v + v_base + v);
                    }
                }
                v_base += batch_v;
                i_base += batch_i;
            }
            consumer_sync<NT>();  // every ring / active read of this batch is done
            pend_count = 0;
        };

        // ---- P2: the duty group classifies step st = cell layers 2st-2, 2st-1 --------------------
        // (needs the bits of sample layers 2st-1 .. 2st+1, i.e. slabs st-1 and st)
        auto classify_step = [&](int st) {
            if (p2_group != st % C::PG) return;
            const uint32_t s1 = sc0 + static_cast<uint32_t>(st - 1), s2 = s1 + 1;
            mbar_wait(&sm.bits_bar[s1 & (C::BR - 1)], (s1 / C::BR) & 1u);
            mbar_wait(&sm.bits_bar[s2 & (C::BR - 1)], (s2 / C::BR) & 1u);
            const int r = p2_sub * 32 + lane, zl = r / E, y = r % E, z = 2 * st - 2 + zl;
            // sample layers z+1, z+2 = 2st-1+zl, 2st+zl
            const uint32_t* bp0 = zl == 0 ? sm.bits[s1 & (C::BR - 1)] : sm.bits[s2 & (C::BR - 1)];
            const int base0 = zl == 0 ? LW : 0;
            const uint32_t* bp1 = sm.bits[s2 & (C::BR - 1)];
            const int base1 = zl == 0 ? 0 : LW;
            const RowCorners rc = load_row_corners<C>(bp0, base0, bp1, base1, y);
            const uint64_t any = rc.a00 | rc.b00 | rc.a10 | rc.b10 | rc.a01 | rc.b01 | rc.a11 | rc.b11;
            const uint64_t all = rc.a00 & rc.b00 & rc.a10 & rc.b10 & rc.a01 & rc.b01 & rc.a11 & rc.b11;
            uint64_t act = any & ~all & ROWMASK;
            if (dirty != ~0ull) {
                const uint32_t nib = static_cast<uint32_t>(dirty >> (4 * ((y / C::QW) + 4 * (z / C::QW)))) & 15u;
                uint64_t dirty_x = 0;
#pragma unroll
                for (int mx = 0; mx < 4; ++mx)
                    if ((nib >> mx) & 1u) dirty_x |= ((1ull << C::QW) - 1ull) << (mx * C::QW);
                act &= dirty_x;
            }
            const uint32_t vi = stc0 + static_cast<uint32_t>(st - 1);
            sm.active[vi & (C::BR - 1)][r] = act;
            const bool any_row = __any_sync(0xffffffffu, act != 0);
            __syncwarp();  // every lane's active-mask store is ordered before lane 0's releasing arrive
            if (lane == 0) {
                sm.verdict_flag[vi & (C::BR - 1)][p2_sub] = any_row ? 1u : 0u;
                mbar_arrive(&sm.verdict_bar[vi & (C::BR - 1)]);
            }
        };

        // ---- verdict of a step: queue it for emission, or hand the oldest slab back --------------
        // After step st is resolved, slab st - 1 (sample layers 2st-2, 2st-1) is no longer needed.
        auto handle_verdict = [&](int st) {
            const uint32_t vi = stc0 + static_cast<uint32_t>(st - 1);
            mbar_wait(&sm.verdict_bar[vi & (C::BR - 1)], (vi / C::BR) & 1u);
            const uint32_t* f = sm.verdict_flag[vi & (C::BR - 1)];
            uint32_t any_flag = 0;
#pragma unroll
            for (int i = 0; i < C::PWS; ++i) any_flag |= f[i];
            if (any_flag != 0u || debug) {
                if (pend_count == 0) pend_first = st;
                if (++pend_count == C::EBS) {
                    flush();
                    release_through(st - 1);
                }
            } else {
                if (pend_count != 0) flush();
                release_through(st - 1);
            }
        };

        // ---- walk the chunk's slabs as they land ----------------------------------------------
        bool stop = false;
        for (int j = 0; j < C::NSLAB; ++j) {
            mbar_wait(&sm.full_bar[slot], round & 1u);
            if (j == 0) {
                chunk = sm.chunk_ids[kc & 3];
                if (chunk >= p.n_chunks) {
                    stop = true;
                    break;
                }
                desc = p.descs[chunk];
                dirty = desc.dirty_microbricks;
                tmask = desc.transition_mask & 0x3fu;
                out_v = p.vertices + static_cast<size_t>(chunk) * p.max_vertices;
                out_i = p.indices + static_cast<size_t>(chunk) * p.max_indices;
            }
            if (p.mode == MODE_STREAM_ONLY) {  // diagnostics: the bare HBM -> smem pipeline
                release_through(j);
            } else {
                {
                    // ---- P1: this warp's BPW ballot blocks of the slab: LDS, sign test, VOTE, STS ----
                    const short* src = reinterpret_cast<const short*>(&sm.ring[slot][0]) + 2 * (32 * warp + lane);
                    uint32_t* dst = sm.bits[sc & (C::BR - 1)] + warp;
                    // all loads first (the compiler does not hoist shared loads across ballots), then
                    // the votes: BPW independent LDS in flight instead of one
                    short dens[C::BPW];
#pragma unroll
                    for (int k = 0; k < C::BPW; ++k) {
                        const bool in_range = k * NW + NW <= C::FULL || warp < C::FULL - k * NW;
                        dens[k] = in_range ? src[64 * NW * k] : short(1);
                    }
#pragma unroll
                    for (int k = 0; k < C::BPW; ++k) {
                        const uint32_t b = __ballot_sync(0xffffffffu, dens[k] <= 0);
                        if (lane == 0 && (k * NW + NW <= C::FULL || warp < C::FULL - k * NW)) dst[NW * k] = b;
                    }
                    if (C::TAIL != 0 && warp == NW - 1) {
                        const short* tail = reinterpret_cast<const short*>(&sm.ring[slot][0]) + 2 * (32 * C::FULL);
                        const short d = lane < C::TAIL ? tail[2 * lane] : short(1);
                        const uint32_t b = __ballot_sync(0xffffffffu, d <= 0);
                        if (lane == 0) sm.bits[sc & (C::BR - 1)][C::FULL] = b;
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.bits_bar[sc & (C::BR - 1)]);
                }
                if (p.mode == MODE_BITS_ONLY) {
                    release_through(j);
                } else {
                    // P2 runs one slab behind P1 (its bits have been complete for a whole iteration, so
                    // the duty group does not block) and the verdict is consumed one slab later still.
                    if (j >= 2) classify_step(j - 1);
                    if (j >= 3) handle_verdict(j - 2);
                }
            }
            ++sc;
            if (++slot == RS) {
                slot = 0;
                ++round;
            }
        }
        if (stop) break;
        if (p.mode != MODE_STREAM_ONLY && p.mode != MODE_BITS_ONLY) {
            // drain: the last step covers cell layers E-2, E-1 (one-sided gradient on the +z face)
            classify_step(C::NSLAB - 1);
            handle_verdict(C::NSLAB - 2);
            handle_verdict(C::NSLAB - 1);
            if (pend_count != 0) flush();
            release_through(C::NSLAB - 1);
            stc += C::NSLAB - 1;
        }

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
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, regular_extract_kernel<C>, C::NT_ALL, smem);
    if (err != cudaSuccess) return err;
    if (ctas_per_sm < 1) return cudaErrorInvalidConfiguration;
    const uint32_t grid = static_cast<uint32_t>(
        min(static_cast<long long>(p.n_chunks), static_cast<long long>(dev.sm_count) * ctas_per_sm));
    regular_extract_kernel<C><<<grid, C::NT_ALL, smem, stream>>>(p);
    return cudaGetLastError();
}

using Cfg64 = Cfg<64, 2, 6, 16>;
using Cfg32 = Cfg<32, 2, 10, 8>;

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
            case 1: return launch_cfg<Cfg<64, 1, 6, 16>>(p, dev, stream);
            case 2: return launch_cfg<Cfg<64, 1, 6, 15>>(p, dev, stream);
            case 3: return launch_cfg<Cfg<64, 1, 6, 11>>(p, dev, stream);
            case 4: return launch_cfg<Cfg<64, 1, 5, 16>>(p, dev, stream);
            default: return launch_cfg<Cfg64>(p, dev, stream);
        }
    }
    if (edge == 32) {
        switch (variant) {
            case 1: return launch_cfg<Cfg<32, 1, 6, 8>>(p, dev, stream);
            case 2: return launch_cfg<Cfg<32, 2, 8, 8>>(p, dev, stream);
            case 3: return launch_cfg<Cfg<32, 1, 8, 4>>(p, dev, stream);
            default: return launch_cfg<Cfg32>(p, dev, stream);
        }
    }
    return cudaErrorInvalidValue;
}

}  // namespace hvx
