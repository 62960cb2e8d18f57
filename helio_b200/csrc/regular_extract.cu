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
#include "hvx_device.cuh"
#include "hvx_kernels.h"

// edge-32 decoupled kernel: emission warps, ring slots and CTAs per SM (tuning knobs; measured on B200:
// 3 CTAs x (4 front + 6 emission + 2) warps beat 2 x (4 + 8 + 2) by 17-20 % -- the per-slab front-end chain is
// latency bound at this slab size, so a third CTA per SM is a third chain in flight)
// the decoupled kernel's edge parameter: 1 = branch-free division (edge_parameter_int16), 0 = __fdiv_rn with its
// range check and slow-path call; bit-identical (tests/test_gpu_regular.py::test_edge_parameter_division_is_exact)
#ifndef HVX_FAST_EDGE
#define HVX_FAST_EDGE 1
#endif
#ifndef HVX_E32_NW
#define HVX_E32_NW 6
#endif
#ifndef HVX_E32_RS
#define HVX_E32_RS 6
#endif
#ifndef HVX_E32_CTAS
#define HVX_E32_CTAS 3
#endif
// Edge 32: a step whose number of active cells is a multiple of 32 is cut into tiles of 32 cells instead of 30.  That
// is every step of a planar surface (a horizon-plan page of the planet set crosses a step in two full rows: 64 cells =
// two tiles of four full vertex passes each instead of 30 + 30 + 4) and one step in 32 of anything else, where a 32-cell
// tile costs a fifth, nearly empty vertex pass (terrain cells carry five or six vertices here and there).  The rule
// needs nothing from the front end, whose per-slab chain bounds a page: telling planar rows from terrain there was
// tried three ways (a vote, a second count carried by the scan, a same-case-along-the-row test) and cost the terrain
// batch 2-5 %.  Fixed tile sizes 30 / 31 / 32: planet set 2.36 / 2.35 / 2.10 ms, 4096 terrain pages 0.173 / 0.172 /
// 0.179 ms.
// Edge 64 has no shared memory left for the larger owner map.
#ifndef HVX_E32_WIDE_TILES
#define HVX_E32_WIDE_TILES 1
#endif

namespace hvx {

namespace {

// ---- stress-build hooks (tools/repro_race.py; both compile to nothing in the product build) -------------------
// HVX_JITTER: random sleeps at every hand-off of the decoupled kernel, so orderings that the protocol allows
//   but a quiet machine never produces do occur (1 = one draw per warp, 2 = one draw per lane: also splits warps
//   in front of every collective).  HVX_SELFCHECK: invariants of a tile (cell select inside the row mask, a
//   surface case, vertex ownership, prefix fields) recorded into a device log instead of faulting later.
#ifdef HVX_JITTER
__device__ __forceinline__ void jitter(uint32_t salt) {
    uint32_t c = (static_cast<uint32_t>(clock()) ^ (salt * 0x9E3779B9u) ^ (blockIdx.x * 0x85EBCA6Bu)) * 2654435761u;
#if HVX_JITTER >= 2
    c = (c ^ ((threadIdx.x & 31u) * 0xC2B2AE35u)) * 2246822519u;
#else
    c = (c ^ ((threadIdx.x >> 5) * 0xC2B2AE35u)) * 2246822519u;
#endif
    if ((c >> 29) == 0u) __nanosleep((c >> 8) & 8191u);
}
#define HVX_JIT(salt) jitter(salt)
#else
#define HVX_JIT(salt) ((void)0)
#endif
#ifdef HVX_SELFCHECK
__device__ uint32_t g_selfcheck_count;
__device__ uint32_t g_selfcheck_log[64][8];
__device__ __noinline__ void selfcheck_fail(uint32_t code, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e) {
    const uint32_t i = atomicAdd(&g_selfcheck_count, 1u);
    if (i < 64u) {
        uint32_t* r = g_selfcheck_log[i];
        r[0] = code; r[1] = blockIdx.x; r[2] = threadIdx.x; r[3] = a; r[4] = b; r[5] = c; r[6] = d; r[7] = e;
    }
}
#define HVX_CHECK(cond, code, a, b, c, d, e) do { if (!(cond)) selfcheck_fail(code, a, b, c, d, e); } while (0)
#else
#define HVX_CHECK(cond, code, a, b, c, d, e) ((void)0)
#endif

// HVX_TIMELINE (tools/timeline.py; variant build only): a few time stamps per CTA of the decoupled kernel -- when the CTA
// was up, when it drew each ticket, when a walk's first slab had landed, when a chunk's records were written -- to see
// where a short launch spends its time (launch ramp, first-slab latency, tail).  Event = globaltimer ns (44 bits) |
// kind << 44 | id << 48; row 0 of a CTA's record counts its events.
#ifdef HVX_TIMELINE
__device__ unsigned long long g_timeline[1024][64];
__device__ __forceinline__ void timeline_event(uint32_t kind, uint32_t id) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    unsigned long long* row = g_timeline[blockIdx.x & 1023u];
    const unsigned long long n = atomicAdd(&row[0], 1ull);
    if (n < 63ull) row[1ull + n] = (t & ((1ull << 44) - 1ull)) | (static_cast<unsigned long long>(kind) << 44) |
                                   (static_cast<unsigned long long>(id & 0xffffu) << 48);
}
#define HVX_TL(kind, id) timeline_event(kind, id)
#else
#define HVX_TL(kind, id) ((void)0)
#endif
// HVX_WAITTRACE (tools/wait_trace.py; variant build only): every warp role records the wait it is about to enter
// (site | detail << 8) in a per-CTA table in global memory and clears it afterwards, so a launch that does not finish
// can be asked, from the host and while it hangs, what each warp of each CTA is waiting for.
#ifdef HVX_WAITTRACE
__device__ uint32_t g_wait[1024][32];
#define HVX_WAIT_BEGIN(site, detail) (*const_cast<volatile uint32_t*>(&g_wait[blockIdx.x & 1023u][threadIdx.x >> 5]) = (site) | (static_cast<uint32_t>(detail) << 8))
#define HVX_WAIT_END() (*const_cast<volatile uint32_t*>(&g_wait[blockIdx.x & 1023u][threadIdx.x >> 5]) = 0u)
#else
#define HVX_WAIT_BEGIN(site, detail) ((void)0)
#define HVX_WAIT_END() ((void)0)
#endif
enum : uint32_t {
    WS_PRODUCER_EMPTY = 1, WS_FRONT_FIRST_FULL = 2, WS_FRONT_FULL = 3, WS_FRONT_BAR = 4, WS_SCHED_FULL = 5, WS_SCHED_REC = 6,
    WS_SCHED_QFREE = 7, WS_EMIT_QBAR = 8, WS_EMIT_NEXT_SLAB = 9, WS_EMIT_CHAIN = 10, WS_EMIT_CHUNK_TOTAL = 11,
    WS_EMIT_LOOKBACK = 12, WS_EMIT_RECORDS_LOOKBACK = 13, WS_DONE = 15
};
enum : uint32_t { TL_CTA_UP = 0, TL_TICKET = 1, TL_FIRST_SLAB = 2, TL_CHUNK_END = 3, TL_PRODUCER_EXIT = 4 };

// Lengyel's tables live in device global memory (statically initialised); every CTA copies the
// 3.8 KB it needs into shared memory once (the kernel is persistent).
#define HVX_TABLE static __device__ const
#include "transvoxel_tables.inc"

// Nanoseconds an idle role (scheduler lane, emission warp without work) sleeps between two looks at its barrier.
// Measured (variants idle0 / default 100 / idle400): headline 0.754 / 0.755 / 0.755 ms, all-surface 1.791 / 1.797 / 1.798 ms,
// planet shard of 3141 pages 0.326 / 0.351 ms -- the issue slots the polling costs are not what limits these launches,
// and the late wake-up costs the small batch 7 %: off.
#ifndef HVX_IDLE_NS
#define HVX_IDLE_NS 0
#endif

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
                const uint32_t ticket = atomicAdd(p.work_counter, 1u);
                const uint32_t id = ticket < p.n_work ? (p.order != nullptr ? p.order[ticket] : ticket) : 0xffffffffu;
                // chunk_ids[k & 3] was last read for local chunk k - 4, long retired (RS < NSLAB)
                sm.chunk_ids[k & 3] = id;
                if (id >= p.n_chunks) {
                    rearm_work_counter(p.work_counter);
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
            rg.first_vertex = (p.chunk_base + chunk) * p.max_vertices;
            rg.vertex_count = ok ? v_base : 0u;
            rg.first_index = (p.chunk_base + chunk) * p.max_indices;
            rg.index_count = ok ? i_base : 0u;
            p.ranges[chunk] = rg;
        }
    }
}


// Same vertex, restated for the decoupled kernel with fewer instructions and no data-dependent
// branches; every float it produces has the same bits as emit_regular_vertex:
//   * densities are converted with the 2^23 trick: (w & 0xffff) ^ 0x4B008000 is the float
//     8421376 + d exactly, so differences of two biased values are the exact integer differences
//     (what fsub(float(d1), float(d0)) gives), and one exact subtraction recovers d itself;
//   * an edge runs along one axis from corner c0 to c1 = c0 + one bit (checked over the whole table),
//     so mix(P0, P1, t) is P0 + 1*t on that axis and P0 + 0*t = P0 on the others;
//   * the one-sided gradient on the +face becomes "load the centre instead of the +1 neighbour and
//     scale by 1 instead of 0.5".
// wl[d] is the ring word offset of sample layer (first layer of the step + d).
template <class C>
__device__ __forceinline__ void emit_regular_vertex_fast(const uint32_t* __restrict__ ring, const uint32_t* __restrict__ wl,
                                                         int x, int y, int zl, int z, uint32_t code, uint32_t transition_mask,
                                                         hvx_vertex* dst) {
    constexpr int S = C::S;
    constexpr float BIAS = 8421376.0f;  // 2^23 + 2^15
    auto biased = [](uint32_t w) { return __uint_as_float((w & 0xffffu) ^ 0x4B008000u); };
    const int c0 = code >> 4, axis = (code >> 4) ^ (code & 15);  // axis: 1 = x, 2 = y, 4 = z
    const int ax = x + 1 + (c0 & 1), ay = y + 1 + ((c0 >> 1) & 1), ea = zl + 1 + ((c0 >> 2) & 1);
    const int bx = ax + (axis & 1), by = ay + ((axis >> 1) & 1), eb = ea + (axis >> 2);
    const int az = z - zl + ea, bz = z - zl + eb;  // absolute sample-layer indices
    auto endpoint = [&](int xi, int yi, int e, int zi, uint32_t& w, float& m, float g[3]) {
        const bool hx = xi >= C::E + 1, hy = yi >= C::E + 1, hz = zi >= C::E + 1;
        const int off = yi * S + xi;
        const uint32_t* l0 = ring + wl[e] + off;
        const uint32_t* lm = ring + wl[e - 1] + off;
        const uint32_t* lp = ring + wl[hz ? e : e + 1] + off;
        w = l0[0];
        m = biased(w);
        const float xm = biased(l0[-1]), ym = biased(l0[-S]), zm = biased(lm[0]);
        const float xp = biased(l0[hx ? 0 : 1]), yp = biased(l0[hy ? 0 : S]), zp = biased(lp[0]);
        g[0] = fmul(fsub(xp, xm), hx ? 1.0f : 0.5f);
        g[1] = fmul(fsub(yp, ym), hy ? 1.0f : 0.5f);
        g[2] = fmul(fsub(zp, zm), hz ? 1.0f : 0.5f);
    };
    uint32_t wa, wb;
    float ma, mb, ga[3], gb[3];
    endpoint(ax, ay, ea, az, wa, ma, ga);
    endpoint(bx, by, eb, bz, wb, mb, gb);
    const float d0 = fsub(ma, BIAS), d1 = fsub(mb, BIAS);
#if HVX_FAST_EDGE
    const float t = edge_parameter_int16(d0, d1);
#else
    const float t = edge_parameter(d0, d1);
#endif
    float p[3], n[3];
    const float fx = static_cast<float>(ax - 1), fy = static_cast<float>(ay - 1), fz = static_cast<float>(az - 1);
    p[0] = (axis & 1) ? fadd(fx, t) : fx;
    p[1] = (axis & 2) ? fadd(fy, t) : fy;
    p[2] = (axis & 4) ? fadd(fz, t) : fz;
    const float gx = fmix(ga[0], gb[0], t), gy = fmix(ga[1], gb[1], t), gz = fmix(ga[2], gb[2], t);
    const float s = fadd(fadd(fmul(gx, gx), fmul(gy, gy)), fmul(gz, gz));
    const bool sound = s > 1.0e-12f;
    // s is a sum of squares of differences of i16 densities: 1e-12 < s < 2^35, inside the branch-free range
    const float inv = inv_sqrt_rn_normal(sound ? s : 1.0f);
    n[0] = sound ? fmul(gx, inv) : 0.0f;
    n[1] = sound ? fmul(gy, inv) : 1.0f;
    n[2] = sound ? fmul(gz, inv) : 0.0f;
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

// position of the k-th (0-based) set bit of w; k < popc(w)
__device__ __forceinline__ int select_bit32(uint32_t w, uint32_t k) {
    int pos = 0;
    uint32_t c = __popc(w & 0xffffu);
    if (k >= c) { k -= c; w >>= 16; pos += 16; }
    c = __popc(w & 0xffu);
    if (k >= c) { k -= c; w >>= 8; pos += 8; }
    c = __popc(w & 0xfu);
    if (k >= c) { k -= c; w >>= 4; pos += 4; }
    c = __popc(w & 0x3u);
    if (k >= c) { k -= c; w >>= 2; pos += 2; }
    if (k >= (w & 1u)) pos += 1;
    return pos;
}

// =============================================================================================
// Second-generation kernel (the default): DECOUPLED front end / emission with a work queue.
//   * FW FRONT warps do all per-slab work.  Every front warp turns its share of slab j into sign
//     bits; after one named barrier CW of them classify step j (cell layers 2j-2, 2j-1; one lane per
//     cell row) and store the row masks and row ranks in the ring slot of slab j (rec_bar[slot]).
//   * one SCHEDULER lane turns step records into work: a step WITH surface cells becomes an entry of
//     a 16-deep work queue (q_bar[i] signals it); a step WITHOUT is finished on the spot.  It runs
//     beside the front warps, so their per-slab loop has no serial section.
//   * NW EMISSION warps sleep on the queue (a parked mbarrier wait costs no issue slots while the
//     chunk is empty), walk its entries in order and pull TILES of TC consecutive active cells from
//     the entry's counter.  A tile is handled by one warp alone: locate the cells (binary search
//     over row ranks + k-th-set-bit select), cases from the bricks, warp scan of (vertices |
//     indices), indices, then one lane per vertex through a warp-private owner map.  Placement
//     across tiles is a chained prefix in shared memory: tile q polls tile q-1's inclusive totals
//     (one 64-bit word tagged with q) and publishes its own before it starts writing.
//   * slab s is handed back to the producer after THREE arrivals on empty[s]: "step s-1 done",
//     "step s done", "step s+1 done" (the steps whose emission window contains it).  The scheduler
//     makes them for empty and non-existent steps, the warp that finishes the last tile of a step
//     makes them for that step.  The slab ring is therefore the run-ahead window: the front end
//     classifies up to RS-2 slabs ahead of the oldest unfinished step and several steps' tiles are in
//     flight at once.  q_free[i] (one arrival per emission warp) keeps the scheduler from reusing a
//     queue entry somebody has not read yet.
//   * chunk totals travel with the chain: whoever handles the last tile of a step stores the
//     inclusive totals in chunk_total[] before publishing; a CHUNK_END queue entry tells one
//     emission warp to wait for the final value and write the chunk's counter records.
//   * THE RULE every wait obeys: a warp touches shared state that can be recycled -- a queue entry, a step's
//     slabs and their barriers, a prefix word -- only while something it holds keeps that state from being
//     recycled: its missing q_free arrival (the entry), a pulled tile that is not done (the step's slabs), the
//     strictly ordered publication of the chain (prefix words), the queue depth (chunk_total).  An emission warp
//     decides "nothing to do here" from the entry's one state word (kind | tiles << 16 | next tile) and waits for
//     slab s+1 only after its first successful pull of a tile of step s.
// =============================================================================================
// Partially dirty chunks (the incremental-edit path): which steps can hold a dirty cell, and which slabs those
// steps' classification (slabs s-1, s) and emission window (slabs s-1, s, s+1) read.  Dirty bit = mx + 4 my + 16 mz,
// a z-microbrick is QW cell layers = QW / 2 steps.  Slabs outside the set are never fetched: the producer completes
// their full-barrier phase with zero bytes, so a 2x2x2-microbrick edit streams 18 of a chunk's 33 slabs instead of
// all of them.  (The front end still ballots the stale slot -- cheap, and the dirty mask zeroes every cell of a step
// that is not dirty; teaching it to skip cost the headline batch 3 % in registers and per-slab tests.)  A fully
// dirty chunk (the normal case) takes every slab.
template <class C>
__device__ __forceinline__ uint64_t dirty_steps(uint64_t dirty) {
    if (dirty == ~0ull) return ~0ull;
    uint64_t steps = 0;
#pragma unroll
    for (int mz = 0; mz < 4; ++mz)
        if ((dirty >> (16 * mz)) & 0xffffull) steps |= ((1ull << (C::QW / 2)) - 1ull) << (mz * (C::QW / 2) + 1);
    return steps;
}
__device__ __forceinline__ uint64_t slabs_of_steps(uint64_t steps) { return steps | (steps << 1) | (steps >> 1); }

enum : uint32_t { QK_STEP = 0, QK_CHUNK_END = 1, QK_EXIT = 2 };

template <class C>
struct SmemD {
    static constexpr int NQ = 16;
    alignas(128) uint32_t ring[C::RS][C::SLAB_WORDS];
    // queue entry: [0] slot | st<<4 | tmask<<12 | chunk parity<<18 | first-of-chunk<<20 | parity of full[slot+1]<<21 |
    // kind<<30, [1] chunk, [2] active cells, [3] first tile seq, [4..6] cumulative cells of front warps 0..2;
    // CHUNK_END: [2] chunk's active cells, [3] last tile seq, [4],[5] dirty mask, [7] chunk has tiles
    alignas(16) uint32_t queue[NQ][8];
    alignas(8) uint64_t full_bar[C::RS];
    uint64_t empty_bar[C::RS];
    uint64_t rec_bar[C::RS];               // step record of slab j complete (CW arrivals)
    uint64_t q_bar[NQ];
    uint64_t q_free[NQ];
    uint64_t active[C::RS][C::STEP_ROWS];  // per step (ring slot of its newest slab): active-cell bits per row
    uint64_t tile_prefix[32];              // chained tile totals: vertices | indices<<22 | tag<<44
    uint64_t chunk_total[16];              // totals after the last tile of the newest finished step of work item kc & 15
                                           // (the queue is 16 deep, so an item 16 later cannot be published before every
                                           // warp has left this item's CHUNK_END entry)
    uint32_t bits[3][C::BW];               // solid bits of the last three slabs (front warps only)
    uint32_t q_ctr[NQ];                    // state of the entry in one word: next tile (low 16 bits) | tiles << 16 for a step,
                                           // kind << 30 for the others -- a warp that finds no work reads nothing else
    uint32_t q_done[NQ];                   // finished tiles of the entry
    uint32_t wtot[C::RS][C::PWS];          // active cells per classifying warp
    uint32_t wlayer[C::NW][8];             // per emission warp: ring word offset of sample layer z0 + d
    uint32_t chunk_ids[8];                 // work item of the CTA's k-th walk, k & 7 (an item is at least three slabs, the ring six)
    uint16_t rowrank[C::RS][C::STEP_ROWS]; // first cell rank of a row, relative to its classifying warp
    uint16_t case_info[256];
    alignas(16) uint8_t class_index[16 * 16];
    uint16_t vertex_base[256];             // first entry of a case's run in vertex_packed
    uint8_t vertex_packed[1536];           // edge codes (corner pair) of every case's vertices, back to back
    uint8_t owner[C::NW][C::E == 32 ? 384 : 360];  // per emission warp: tile vertex -> lane (cell) that owns it (tile cells * 12)
};

template <class C>
struct DecoupledCfg {
    static constexpr int CW = C::PWS;                          // classifying front warps (one lane per cell row)
    static constexpr int FW = 2 * C::PWS;                      // front warps; all of them turn slabs into sign bits
    static constexpr int NT_ALL = (FW + C::NW + 2) * 32;       // front + emission + producer + scheduler
    static constexpr int GB = CW + 3 * (FW - CW);              // ballot blocks per group: 1 per classifying warp, 3 per other
    static constexpr int NG = C::FULL / GB;                    // groups per slab = loads in flight per batch (17 / 9)
    // cells per tile: a typical surface cell has 4 vertices, so 30 cells fill four 32-lane vertex
    // passes (~120 vertices); 32 cells would spill a handful of vertices into a fifth
    static constexpr int TC = 30;
    static constexpr bool WIDE = C::E == 32 && HVX_E32_WIDE_TILES != 0;  // full-row steps use 32-cell tiles (see HVX_E32_WIDE_TILES)
    static_assert(TC * 12 <= 360, "owner map size");
    static_assert(C::STEP_ROWS == CW * 32, "one classifying lane per cell row");
    static_assert(C::FULL % GB == 0 && NG <= 18, "the slab's ballot blocks split into whole groups");
    static_assert(CW <= 4 && C::RS <= 15 && C::NSLAB <= 255, "queue entry packing");
};

// Records of a chunk that is not walked at all (HVX_CHUNK_UNIFORM, or an empty dirty mask).
template <class C>
__device__ __forceinline__ void write_empty_records(const RegularParams& p, uint32_t chunk) {
    hvx_emission_counters ec{};
    ec.completed = 1u;
    p.counters[chunk] = ec;
    hvx_classify_counters cc{};
    cc.visited_cells = static_cast<uint32_t>(__popcll(p.descs[chunk].dirty_microbricks)) * (C::QW * C::QW * C::QW);
    p.classify[chunk] = cc;
    hvx_range rg{};
    rg.first_vertex = (p.chunk_base + chunk) * p.max_vertices;
    rg.first_index = (p.chunk_base + chunk) * p.max_indices;
    p.ranges[chunk] = rg;
}

// One walk of a CTA: a whole chunk, or (SPLIT) one z-range of a chunk.  Steps [j0 + 1, s_last] are classified and
// emitted; slabs [j0, j1] are streamed (j1 = s_last + 1 only lends its first layer to the +z gradient of step s_last).
struct WorkItem {
    uint32_t chunk;
    int j0, j1, s_last;
    uint32_t tag;  // SPLIT: item index | part << 24 | is-last-part << 28
};
template <class C, bool SPLIT>
__device__ __forceinline__ WorkItem work_item(const RegularParams& p, uint32_t id) {
    if (!SPLIT) return {id, 0, C::NSLAB - 1, C::NSLAB - 1, 0u};
    const SplitItem it = p.items[id];
    const int s_last = it.s_last;
    return {it.chunk, static_cast<int>(it.s_first) - 1, min(s_last + 1, C::NSLAB - 1), s_last,
            id | (static_cast<uint32_t>(it.part) << 24) | ((it.part + 1u == it.parts ? 1u : 0u) << 28)};
}

// part 0 and the last part at once: the item is a whole chunk (a batch whose last wave is split keeps its other chunks whole)
__device__ __forceinline__ bool split_whole(uint32_t tag) { return ((tag >> 24) & 31u) == 16u; }

// Look-back state of the split walk: item_totals[i] = (vertices, indices, active cells, generation).  A part's counting
// walk stores the three totals, fences, then stores the dispatch's generation; readers spin on the generation (parts
// are claimed in ascending order, so a part's predecessors are always running or done: no dead-lock) and fence before
// they read the totals.
__device__ __forceinline__ void publish_part(uint4* totals, uint32_t v, uint32_t i, uint32_t cells, uint32_t generation) {
    volatile uint32_t* w = reinterpret_cast<volatile uint32_t*>(totals);
    w[0] = v;
    w[1] = i;
    w[2] = cells;
    __threadfence();
    w[3] = generation;
}
__device__ __forceinline__ uint4 wait_part(const uint4* totals, uint32_t generation) {
    const volatile uint32_t* w = reinterpret_cast<const volatile uint32_t*>(totals);
    while (w[3] != generation) {
    }
    __threadfence();
    return make_uint4(w[0], w[1], w[2], generation);
}

// PARTIAL: the batch holds partially dirty chunks (incremental edits): slabs no dirty step reads are neither
// fetched nor balloted nor classified.  A separate instantiation, so the fully dirty batches (the headline) keep
// the front-end loop free of the per-slab test and its registers (measured: 0.804 vs 0.814 ms with it compiled in).
// SPLIT: the work list names z-ranges of chunks (SplitItem) instead of chunks, so that a few chunks still fill the
// machine (the latency configurations: one page, an edit frame).  Placement across the parts of a chunk is a decoupled
// look-back over per-part totals in global memory: every part is walked twice by its CTA -- the counting walk
// classifies, counts and publishes (vertices, indices, cells) tagged with the dispatch's generation; the emitting walk
// starts at the sum of the parts before it, spinning on their tags if they are not there yet (parts are claimed in
// ascending order, so predecessors are running or done) -- and the chunk's last part writes the chunk's records.
// Output is byte-identical to the unsplit walk.
template <class C, bool PARTIAL, bool SPLIT>
__global__ void __launch_bounds__(DecoupledCfg<C>::NT_ALL, C::E == 32 ? HVX_E32_CTAS : 1)
regular_extract_decoupled_kernel(const RegularParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    using SM = SmemD<C>;
    using D = DecoupledCfg<C>;
    SM& sm = *reinterpret_cast<SM*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int E = C::E, S = C::S, RS = C::RS, NW = C::NW, LW = C::LAYER_WORDS, FW = D::FW, CW = D::CW, NQ = SM::NQ;
    constexpr uint64_t ROWMASK = E == 64 ? ~0ull : 0xffffffffull;
    constexpr uint64_t FIELD = (1ull << 22) - 1ull;
    const size_t chunk_words = static_cast<size_t>(S) * S * S;

    for (int i = tid; i < 256; i += D::NT_ALL) sm.case_info[i] = HVX_REGULAR_CASE_INFO[i];
    for (int i = tid; i < 256; i += D::NT_ALL) sm.vertex_base[i] = HVX_REGULAR_VERTEX_BASE[i];
    for (int i = tid; i < 1536; i += D::NT_ALL) sm.vertex_packed[i] = HVX_REGULAR_VERTEX_PACKED[i];
    for (int i = tid; i < 256; i += D::NT_ALL) sm.class_index[i] = HVX_REGULAR_CLASS_INDEX[i / 16][i % 16];
    for (int i = tid; i < 3 * C::BW; i += D::NT_ALL) sm.bits[i / C::BW][i % C::BW] = 0u;  // incl. zero padding
    if (tid < 32) sm.tile_prefix[tid] = ~0ull;                                            // no tile has this tag yet
    if (tid < 16) sm.chunk_total[tid] = ~0ull;
    if (tid == 0) {
        for (int i = 0; i < RS; ++i) {
            mbar_init(&sm.full_bar[i], 1);
            mbar_init(&sm.empty_bar[i], 3);
            mbar_init(&sm.rec_bar[i], CW);
        }
        for (int i = 0; i < NQ; ++i) {
            mbar_init(&sm.q_bar[i], 1);
            mbar_init(&sm.q_free[i], NW);
        }
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) HVX_TL(TL_CTA_UP, 0u);

    // ---- PRODUCER warp ------------------------------------------------------------------------
    if (warp == FW + NW) {
        if (lane != 0) {
            // Lanes 1..31 have nothing to stream: they write the records of the chunks the work list leaves out (flagged
            // uniform, or no dirty microbrick): an empty, completed mesh.  Off everybody's critical path, and one launch
            // less for an edit frame that re-submits 256 resident chunks of which two are dirty.
            for (uint32_t i = blockIdx.x * 31u + static_cast<uint32_t>(lane) - 1u; i < p.n_skipped; i += gridDim.x * 31u)
                write_empty_records<C>(p, p.skipped[i]);
            return;
        }
        if (lane == 0) {
            int slot = 0;
            uint32_t round = 0, k = 0;  // k: walks of this CTA so far
            for (;;) {
                const uint32_t ticket = atomicAdd(p.work_counter, 1u);
                // the work list: every chunk in index order, or the caller's list -- heaviest chunks first when the batch
                // carries cost hints, chunks flagged uniform left out (a chunk's slot does not depend on when it runs)
                const uint32_t id = ticket < p.n_work ? (p.order != nullptr ? p.order[ticket] : ticket) : 0xffffffffu;
                HVX_TL(id == 0xffffffffu ? TL_PRODUCER_EXIT : TL_TICKET, ticket);
                if (id == 0xffffffffu) {
                    rearm_work_counter(p.work_counter);
                    sm.chunk_ids[k & 7] = id;
                    HVX_WAIT_BEGIN(WS_PRODUCER_EMPTY, slot | 0x80);
                    mbar_wait_parked(&sm.empty_bar[slot], (round & 1u) ^ 1u);
                    HVX_WAIT_BEGIN(WS_DONE, 0);
                    mbar_arrive(&sm.full_bar[slot]);
                    break;
                }
                const WorkItem it = work_item<C, SPLIT>(p, id);
                const uint32_t* src = p.samples + static_cast<size_t>(it.chunk) * chunk_words;
                const uint64_t need = PARTIAL ? slabs_of_steps(dirty_steps<C>(p.descs[it.chunk].dirty_microbricks)) : ~0ull;
                // SPLIT: a part is walked twice -- the counting walk publishes its totals, the emitting walk looks back over
                // the parts before it (the second stream of its few slabs comes out of L2)
                // (an item that is a whole chunk -- part 0 and last part -- needs no look-back: one walk)
                for (uint32_t pass = SPLIT && !split_whole(it.tag) ? 0u : 1u; pass < 2u; ++pass, ++k) {
                sm.chunk_ids[k & 7] = id | (SPLIT ? pass << 30 : 0u);
                for (int j = it.j0; j <= it.j1; ++j) {
                    HVX_WAIT_BEGIN(WS_PRODUCER_EMPTY, slot | (j << 8));
                    mbar_wait_parked(&sm.empty_bar[slot], (round & 1u) ^ 1u);
                    HVX_WAIT_END();
                    HVX_JIT(30);
                    if ((need >> j) & 1ull) {
                        mbar_arrive_expect_tx(&sm.full_bar[slot], C::SLAB_BYTES);
                        bulk_g2s(&sm.ring[slot][0], src + static_cast<size_t>(j) * C::SLAB_WORDS, C::SLAB_BYTES,
                                 &sm.full_bar[slot]);
                    } else {
                        mbar_arrive(&sm.full_bar[slot]);  // nobody reads this slab: an empty phase keeps the ring in step
                    }
                    if (++slot == RS) {
                        slot = 0;
                        ++round;
                    }
                }
                }
            }
        }
        return;
    }

    // ---- FRONT warps: slab -> sign bits -> step record --------------------------------
    if (warp < FW) {
        int slot = 0;
        uint32_t round = 0;
        int bcur = 0, bprev = 2;  // bits ring: slab j in bits[bcur], slab j-1 in bits[bprev]
        for (uint32_t kc = 0;; ++kc) {
            uint64_t dirty = 0, need = ~0ull;
            // the first wait of a walk is for its first slab (or the sentinel): only then is the work item known
            HVX_WAIT_BEGIN(WS_FRONT_FIRST_FULL, slot | (kc << 8));
            mbar_wait_parked(&sm.full_bar[slot], round & 1u);
            HVX_WAIT_END();
            const uint32_t idw = sm.chunk_ids[kc & 7];
            if (idw == 0xffffffffu) {
                HVX_WAIT_BEGIN(WS_DONE, 0);
                return;
            }
            if (tid == 0) HVX_TL(TL_FIRST_SLAB, idw);
            const WorkItem it = work_item<C, SPLIT>(p, idw & 0x3fffffffu);
            dirty = p.descs[it.chunk].dirty_microbricks;
            if (PARTIAL) need = slabs_of_steps(dirty_steps<C>(dirty));
            for (int j = it.j0; j <= it.j1; ++j) {
                HVX_WAIT_BEGIN(WS_FRONT_FULL, slot | (j << 8));
                if (j != it.j0) mbar_wait_parked(&sm.full_bar[slot], round & 1u);
                HVX_WAIT_END();
                HVX_JIT(20);
                // a slab no dirty step reads was not fetched: no ballots, and its own step (not dirty) has no cells
                const bool live = !PARTIAL || ((need >> j) & 1ull);
                const bool classify = j > it.j0 && j <= it.s_last && live && p.mode != MODE_STREAM_ONLY && p.mode != MODE_BITS_ONLY;
                if (live && p.mode != MODE_STREAM_ONLY) {
                    // P1: LDS, sign test, VOTE, STS for this warp's ballot blocks.  The warps that classify afterwards
                    // take one block of every group of GB, the others three, so the classifying warps (the critical
                    // path of the slab) reach the barrier with a third of the others' ballot work.  NG loads in
                    // flight, then setp+vote pairs kept adjacent (inline PTX) so the compiler does not park 17
                    // predicates in a register and dig them out again.
                    const short* src = reinterpret_cast<const short*>(&sm.ring[slot][0]) + 2 * lane;
                    uint32_t* dst = sm.bits[bcur];
                    const int batches = warp < CW ? 1 : 3;
                    const int first = warp < CW ? warp : CW + 3 * (warp - CW);
#pragma unroll 1
                    for (int g = 0; g < batches; ++g) {
                        const int b0 = first + g;  // block b0 + GB * k
                        int dens[D::NG];
#pragma unroll
                        for (int k = 0; k < D::NG; ++k) dens[k] = src[64 * (b0 + D::GB * k)];
#pragma unroll
                        for (int k = 0; k < D::NG; ++k) {
                            uint32_t v;
                            asm volatile(
                                "{\n\t.reg .pred p;\n\t"
                                "setp.le.s32 p, %1, 0;\n\t"
                                "vote.sync.ballot.b32 %0, p, 0xffffffff;\n\t}"
                                : "=r"(v)
                                : "r"(dens[k]));
                            if (lane == 0) dst[b0 + D::GB * k] = v;
                        }
                    }
                    if (C::TAIL != 0 && warp == FW - 1) {
                        const short* tail = reinterpret_cast<const short*>(&sm.ring[slot][0]) + 2 * (32 * C::FULL);
                        const short d = lane < C::TAIL ? tail[2 * lane] : short(1);
                        const uint32_t v = __ballot_sync(0xffffffffu, d <= 0);
                        if (lane == 0) sm.bits[bcur][C::FULL] = v;
                    }
                }
                HVX_WAIT_BEGIN(WS_FRONT_BAR, slot | (j << 8));
                asm volatile("bar.sync 2, %0;" ::"n"(FW * 32) : "memory");  // the slab's bits are complete
                HVX_WAIT_END();
                HVX_JIT(21);
                if (warp < CW) {
                    uint32_t incl = 0;
                    if (classify) {
                        // step j = cell layers 2j-2, 2j-1: sample layers 2j-1 (slab j-1), 2j, 2j+1 (slab j)
                        const int r = warp * 32 + lane, zl = r / E, y = r % E, z = 2 * j - 2 + zl;
                        const uint32_t* bp0 = zl == 0 ? sm.bits[bprev] : sm.bits[bcur];
                        const int base0 = zl == 0 ? LW : 0;
                        const uint32_t* bp1 = sm.bits[bcur];
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
                        const uint32_t cnt = static_cast<uint32_t>(__popcll(act));
                        incl = cnt;
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
                            if (lane >= d) incl += up;
                        }
                        sm.active[slot][r] = act;
                        sm.rowrank[slot][r] = static_cast<uint16_t>(incl - cnt);
                    }
                    if (lane == 31) sm.wtot[slot][warp] = incl;
                    HVX_JIT(22);
                    __syncwarp();  // every lane's stores are ordered before lane 0's releasing arrive
                    if (lane == 0) mbar_arrive(&sm.rec_bar[slot]);
                }
                bprev = bcur;
                bcur = bcur == 2 ? 0 : bcur + 1;
                if (++slot == RS) {
                    slot = 0;
                    ++round;
                }
            }
        }
    }

    // ---- SCHEDULER warp: step records -> work queue / slab hand-back (one lane, off everybody's critical path) ----
    if (warp == FW + NW + 1) {
        if (lane != 0) return;
        int slot = 0;
        uint32_t round = 0;
        uint32_t q_tail = 0, tile_total = 0;
        auto publish = [&](uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t w4, uint32_t w5, uint32_t w6,
                           uint32_t w7) {
            const uint32_t qi = q_tail & (NQ - 1);
            HVX_WAIT_BEGIN(WS_SCHED_QFREE, qi | (q_tail << 8));
            if (q_tail >= NQ) mbar_wait(&sm.q_free[qi], ((q_tail / NQ) - 1u) & 1u);  // every emission warp has left its last use
            HVX_WAIT_END();
            *reinterpret_cast<uint4*>(&sm.queue[qi][0]) = make_uint4(w0, w1, w2, w3);
            *reinterpret_cast<uint4*>(&sm.queue[qi][4]) = make_uint4(w4, w5, w6, w7);
            const uint32_t tile_cells = (w0 >> 24) & 1u ? 32u : static_cast<uint32_t>(D::TC);
            sm.q_ctr[qi] = (w0 >> 30) == QK_STEP ? ((w2 + tile_cells - 1u) / tile_cells) << 16 : (w0 >> 30) << 30;
            sm.q_done[qi] = 0u;
            HVX_JIT(11);
            mbar_arrive(&sm.q_bar[qi]);  // release: the entry is visible to whoever sees the phase flip
            ++q_tail;
        };
        for (uint32_t kc = 0;; ++kc) {
            HVX_WAIT_BEGIN(WS_SCHED_FULL, slot | (kc << 8));
            mbar_wait_idle<HVX_IDLE_NS>(&sm.full_bar[slot], round & 1u);
            HVX_WAIT_END();
            const uint32_t idw = sm.chunk_ids[kc & 7];
            if (idw == 0xffffffffu) {
                publish(QK_EXIT << 30, 0, 0, 0, 0, 0, 0, 0);
                HVX_WAIT_BEGIN(WS_DONE, 0);
                return;
            }
            WorkItem it = work_item<C, SPLIT>(p, idw & 0x3fffffffu);
            it.tag |= (idw >> 30) << 29;  // SPLIT: bit 29 = the emitting walk
            const uint32_t chunk = it.chunk;
            const ChunkDesc desc = p.descs[chunk];
            const uint64_t dirty = desc.dirty_microbricks;
            const uint32_t tmask = desc.transition_mask & 0x3fu;
            const uint32_t chunk_first_tile = tile_total;
            uint32_t chunk_cells = 0;
            bool prev_empty = false;
            for (int j = it.j0; j <= it.j1; ++j) {
                HVX_WAIT_BEGIN(WS_SCHED_REC, slot | (j << 8));
                mbar_wait_idle<HVX_IDLE_NS>(&sm.rec_bar[slot], round & 1u);
                HVX_WAIT_END();
                HVX_JIT(10);
                const int prev_slot = slot == 0 ? RS - 1 : slot - 1;
                if (j == it.j0) {
                    // the walk's first slab has no steps j0 - 1 and j0: make their arrivals; the next slab gets "step j0 done" next round
                    mbar_arrive(&sm.empty_bar[slot]);
                    mbar_arrive(&sm.empty_bar[slot]);
                    prev_empty = true;
                } else {
                    if (prev_empty) mbar_arrive(&sm.empty_bar[slot]);  // "step j-1 done" for slab j, now that it is current
                    uint32_t cum[4] = {0, 0, 0, 0};
                    uint32_t n = 0;
#pragma unroll
                    for (int i = 0; i < CW; ++i) {
                        n += sm.wtot[slot][i];
                        cum[i] = n;
                    }
                    // A step whose cell count is a multiple of 32 is cut into 32-cell tiles (see HVX_E32_WIDE_TILES)
                    const uint32_t wide = D::WIDE && (n & 31u) == 0u ? 1u : 0u;
                    const uint32_t tc = wide ? 32u : static_cast<uint32_t>(D::TC);
                    if (n == 0) {
                        mbar_arrive(&sm.empty_bar[slot]);       // "step j done" for slab j
                        mbar_arrive(&sm.empty_bar[prev_slot]);  // "step j done" for slab j-1
                        prev_empty = true;
                    } else {
                        const uint32_t next_parity = (slot + 1 == RS ? round + 1u : round) & 1u;
                        const uint32_t w0 = static_cast<uint32_t>(slot) | (static_cast<uint32_t>(j) << 4) | (tmask << 12) |
                                            ((kc & 15u) << 18) | ((tile_total == chunk_first_tile ? 1u : 0u) << 22) |
                                            (next_parity << 23) | (wide << 24) | (QK_STEP << 30);
                        publish(w0, chunk, n, tile_total, cum[0], cum[1], cum[2], it.tag);
                        tile_total += (n + tc - 1u) / tc;
                        chunk_cells += n;
                        prev_empty = false;
                    }
                }
                if (j == it.j1) {
                    mbar_arrive(&sm.empty_bar[slot]);  // the walk's last slab has no step j1 + 1
                    publish((QK_CHUNK_END << 30) | ((kc & 15u) << 18), chunk, chunk_cells, tile_total - 1u,
                            static_cast<uint32_t>(dirty), static_cast<uint32_t>(dirty >> 32), it.tag,
                            tile_total != chunk_first_tile ? 1u : 0u);
                }
                if (++slot == RS) {
                    slot = 0;
                    ++round;
                }
            }
        }
    }

    // ---- EMISSION warps --------------------------------------------------------------------------
    const int ew = warp - FW;
    const uint32_t* const ring_flat = &sm.ring[0][0];
    uint32_t* const wl = sm.wlayer[ew];
    uint8_t* const ow = sm.owner[ew];
    const bool do_emit = p.mode == MODE_EXTRACT;

    for (uint32_t k = 0;; ++k) {
        const uint32_t qi = k & (NQ - 1);
        HVX_WAIT_BEGIN(WS_EMIT_QBAR, qi | (k << 8));
        mbar_wait_idle<HVX_IDLE_NS>(&sm.q_bar[qi], (k / NQ) & 1u);
        HVX_WAIT_END();
        HVX_JIT(1);
        uint32_t state = 0;
#if !defined(HVX_LEGACY_PROTOCOL) && !defined(HVX_LEGACY_LANE_READ)
        {
            // Every emission warp walks every entry, and most visits find nothing to do (a step of a terrain chunk has
            // ~6 tiles for 20 warps): the entry's state word alone decides that.  ONE decision per warp (lane 0 reads,
            // the shuffle makes the warp agree): the counter moves while the lanes look at it.
            if (lane == 0) state = *const_cast<volatile uint32_t*>(&sm.q_ctr[qi]);
            state = __shfl_sync(0xffffffffu, state, 0);
            const bool idle = (state >> 30) == QK_STEP ? (state & 0xffffu) >= (state >> 16)
                                                       : (state >> 30) == QK_CHUNK_END && static_cast<int>(k % NW) != ew;
            if (idle) {
                if (lane == 0) mbar_arrive(&sm.q_free[qi]);
                continue;
            }
        }
#endif
        const uint4 e0 = *reinterpret_cast<const uint4*>(&sm.queue[qi][0]);
        const uint4 e1 = *reinterpret_cast<const uint4*>(&sm.queue[qi][4]);
        const uint32_t kind = e0.x >> 30;
        if (kind == QK_EXIT) {
            HVX_WAIT_BEGIN(WS_DONE, 0);
            return;
        }
        const uint32_t chunk = e0.y, kcpar = (e0.x >> 18) & 15u;
        if (kind == QK_CHUNK_END) {
            // ---- chunk epilogue: one warp waits for the chain's final totals and writes the records ----
            if (static_cast<int>(k % NW) == ew) {
                uint32_t v_tot = 0, i_tot = 0, cells = e0.z;
                if (e1.w != 0u) {
                    uint64_t got = 0;
                    if (lane == 0) {
                        const volatile uint64_t* tot = &sm.chunk_total[kcpar];
                        const uint64_t want = static_cast<uint64_t>(e0.w & 0xfffffu);
                        HVX_WAIT_BEGIN(WS_EMIT_CHUNK_TOTAL, kcpar | (e0.w << 8));
                        do {
                            got = *tot;
                        } while ((got >> 44) != want);
                        HVX_WAIT_END();
                    }
                    got = __shfl_sync(0xffffffffu, got, 0);
                    v_tot = static_cast<uint32_t>(got & FIELD);
                    i_tot = static_cast<uint32_t>((got >> 22) & FIELD);
                }
                bool records = true;
                if (SPLIT) {
                    const uint32_t item = e1.z & 0xffffffu, part = (e1.z >> 24) & 15u;
                    if (!((e1.z >> 29) & 1u)) {
                        // counting walk: what this part adds to the chunk (every part counts from zero)
                        if (lane == 0) publish_part(&p.item_totals[item], v_tot, i_tot, cells, p.split_generation);
                        records = false;
                    } else if (!((e1.z >> 28) & 1u)) {
                        records = false;  // the chunk's last part reports for the chunk
                    } else if (part != 0u) {
                        // one lane per part (this one included: its counting walk published them), then a warp sum
                        uint32_t a = 0, b = 0, c = 0;
                        HVX_WAIT_BEGIN(WS_EMIT_RECORDS_LOOKBACK, part | (item << 8));
                        if (static_cast<uint32_t>(lane) <= part) {
                            const uint4 t = wait_part(&p.item_totals[item - part + static_cast<uint32_t>(lane)], p.split_generation);
                            a = t.x;
                            b = t.y;
                            c = t.z;
                        }
                        __syncwarp();
                        HVX_WAIT_END();
#pragma unroll
                        for (int d = 16; d != 0; d >>= 1) {
                            a += __shfl_xor_sync(0xffffffffu, a, d);
                            b += __shfl_xor_sync(0xffffffffu, b, d);
                            c += __shfl_xor_sync(0xffffffffu, c, d);
                        }
                        v_tot = a;
                        i_tot = b;
                        cells = c;
                    }  // else: a whole chunk walked once, its own totals
                }
                records = records && lane == 0;
                if (lane == 0) HVX_TL(TL_CHUNK_END, chunk);
                if (records) {
                const uint64_t dirty = static_cast<uint64_t>(e1.x) | (static_cast<uint64_t>(e1.y) << 32);
                const uint32_t vo = v_tot > p.max_vertices ? 1u : 0u, io = i_tot > p.max_indices ? 1u : 0u;
                const bool ok = !(vo | io) && do_emit;
                hvx_emission_counters ec;
                ec.required_vertices = v_tot;
                ec.required_indices = i_tot;
                ec.emitted_vertices = ok ? v_tot : 0u;
                ec.emitted_indices = ok ? i_tot : 0u;
                ec.vertex_overflow = vo;
                ec.index_overflow = io;
                ec.completed = 1u;
                ec._pad = 0u;
                p.counters[chunk] = ec;
                hvx_classify_counters cc;
                cc.visited_cells = static_cast<uint32_t>(__popcll(dirty)) * (C::QW * C::QW * C::QW);
                cc.active_cells = cells;
                cc.vertices = v_tot;
                cc.triangles = i_tot / 3u;
                p.classify[chunk] = cc;
                hvx_range rg;
                rg.first_vertex = (p.chunk_base + chunk) * p.max_vertices;
                rg.vertex_count = ok ? v_tot : 0u;
                rg.first_index = (p.chunk_base + chunk) * p.max_indices;
                rg.index_count = ok ? i_tot : 0u;
                p.ranges[chunk] = rg;
                }
            }
        } else {
            const int slot = e0.x & 15u, st = (e0.x >> 4) & 255u;
            const uint32_t tmask = (e0.x >> 12) & 63u, first_of_chunk = (e0.x >> 22) & 1u, next_parity = (e0.x >> 23) & 1u;
            const uint32_t n_cells = e0.z, tile_base = e0.w;
            const uint32_t cum[3] = {e1.x, e1.y, e1.z};
            const bool emit = do_emit && (!SPLIT || ((e1.w >> 29) & 1u));  // SPLIT: the counting walk writes no mesh
            const uint32_t tc = D::WIDE && ((e0.x >> 24) & 1u) ? 32u : static_cast<uint32_t>(D::TC);  // cells per tile of this step
            const uint32_t ntiles = tc == 32u ? (n_cells + 31u) >> 5 : (n_cells + D::TC - 1u) / D::TC;
            HVX_JIT(2);
            // "are there tiles left" must be ONE decision per warp: the counter moves while the lanes look at it, and
            // a warp whose lanes disagreed would meet itself in the collectives below from two different entries
            // (lane 0 reads, the shuffle makes the warp converge and agree)
#if defined(HVX_LEGACY_PROTOCOL) || defined(HVX_LEGACY_LANE_READ)  // stress builds: the round-1 form (every lane reads for itself)
            const uint32_t pulled = *const_cast<volatile uint32_t*>(&sm.q_ctr[qi]) & 0xffffu;
#else
            const uint32_t pulled = state & 0xffffu;  // the warp's one look at the counter (above): tiles were left
#endif
            if (pulled < ntiles) {   // a hint: whether THIS warp gets a tile is decided by the pull below
                const int slot_m = slot == 0 ? RS - 1 : slot - 1, slot_p = slot + 1 == RS ? 0 : slot + 1;
                hvx_vertex* const out_v = p.vertices + static_cast<size_t>(chunk) * p.max_vertices;
                uint32_t* const out_i = p.indices + static_cast<size_t>(chunk) * p.max_indices;
                const int z0 = 2 * st - 2;
                bool window_ready = false;
                for (;;) {
                    uint32_t t = 0;
                    if (lane == 0) t = atomicAdd(&sm.q_ctr[qi], 1u) & 0xffffu;  // at most tiles + NW increments: the low half never carries
                    t = __shfl_sync(0xffffffffu, t, 0);
                    if (t >= ntiles) break;
                    if (!window_ready) {
                        // Only a warp that HOLDS a tile may touch the step's slabs or their barriers: the step cannot
                        // finish -- and its slabs cannot be handed back and refilled -- before this tile is done.  The
                        // wait for slab st+1 (the z gradient of the step's upper layer reads its first layer) used to
                        // sit in front of the pull; a warp delayed between its look at the counter and that wait could
                        // then find the others had finished the step and the slot refilled twice, and wait for a phase
                        // of the slot's barrier that never comes (found by the jittered stress build, which sleeps
                        // exactly there; profiles/r02_wait_trace.txt).
                        HVX_WAIT_BEGIN(WS_EMIT_NEXT_SLAB, slot_p | (st << 8));
                        if (st + 1 < C::NSLAB) mbar_wait(&sm.full_bar[slot_p], next_parity);
                        HVX_WAIT_END();
                        // ring word offsets of sample layers z0 .. z0+4 (slabs st-1, st, st+1)
                        __syncwarp();
                        if (lane < 6) {
                            const int s2 = lane < 2 ? slot_m : lane < 4 ? slot : slot_p;
                            wl[lane] = static_cast<uint32_t>(s2 * C::SLAB_WORDS + (lane & 1) * LW);
                        }
                        __syncwarp();
                        window_ready = true;
                    }
                    HVX_JIT(3);
                    // ---- one tile: TC consecutive active cells of the step ---------------------------
                    const uint32_t seq = tile_base + t;
                    const bool first = first_of_chunk != 0u && t == 0u;
                    HVX_CHECK(slot < RS && st >= 1 && st < C::NSLAB && seq - tile_base < ntiles, 7u, chunk, st | (slot << 8), seq, t, ntiles);
                    const uint32_t r = tc * t + static_cast<uint32_t>(lane);
                    const bool valid = static_cast<uint32_t>(lane) < tc && r < n_cells;
                    uint32_t rec = 0, packed = 0, info = 0, vbase = 0;
                    if (valid) {
                        // locate: classifying warp -> row (largest row whose first rank <= mine) -> k-th set bit
                        uint32_t g = 0, rr = r;
#pragma unroll
                        for (int i = 0; i + 1 < CW; ++i)
                            if (r >= cum[i]) {
                                g = i + 1;
                                rr = r - cum[i];
                            }
                        const uint16_t* rk = sm.rowrank[slot] + 32u * g;
                        uint32_t i = 0;
#pragma unroll
                        for (uint32_t b = 16; b != 0; b >>= 1)
                            if (rk[i + b] <= rr) i += b;
                        uint32_t kk = rr - rk[i];
                        const uint32_t row = 32u * g + i;
                        const uint64_t m = sm.active[slot][row];
                        uint32_t w = static_cast<uint32_t>(m);
                        int x = 0;
                        if (E == 64) {
                            const uint32_t c = __popc(w);
                            if (kk >= c) {
                                kk -= c;
                                w = static_cast<uint32_t>(m >> 32);
                                x = 32;
                            }
                        }
                        HVX_CHECK(kk < static_cast<uint32_t>(__popc(w)), 1u, chunk, st | (slot << 8), seq, row | (kk << 16), w);
                        x += select_bit32(w, kk);
                        const int zl = row / E, y = row % E;
                        const uint32_t* l0 = ring_flat + wl[zl + 1] + (y + 1) * S + (x + 1);
                        const uint32_t* l1 = ring_flat + wl[zl + 2] + (y + 1) * S + (x + 1);
                        const uint32_t c = (cw_solid(l0[0]) ? 1u : 0u) | (cw_solid(l0[1]) ? 2u : 0u) |
                                           (cw_solid(l0[S]) ? 4u : 0u) | (cw_solid(l0[S + 1]) ? 8u : 0u) |
                                           (cw_solid(l1[0]) ? 16u : 0u) | (cw_solid(l1[1]) ? 32u : 0u) |
                                           (cw_solid(l1[S]) ? 64u : 0u) | (cw_solid(l1[S + 1]) ? 128u : 0u);
                        HVX_CHECK(c != 0u && c != 255u, 2u, chunk, st | (slot << 8), seq, row | (x << 16), c);
                        info = sm.case_info[c];
                        vbase = sm.vertex_base[c];
                        rec = static_cast<uint32_t>(x) | (row << 8) | (c << 16);
                        packed = (info & 15u) | ((3u * ((info >> 4) & 15u)) << 16);
                    }
                    __syncwarp();
                    uint32_t incl = packed;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
                        if (lane >= d) incl += up;
                    }
                    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
                    const uint32_t tot_v = total & 0xffffu, tot_i = total >> 16;
                    const uint32_t vo = (incl - packed) & 0xffffu, io = (incl - packed) >> 16;
                    const uint32_t vo_vb = vo | (vbase << 9);  // what a vertex lane needs from its cell: first vertex, code run
                    // chained prefix: wait for the previous tile's inclusive totals, publish ours.  Every
                    // lane polls the same word (a broadcast read), so the warp never splits around the spin.
                    // The first tile of a chunk starts from zero but still waits for its predecessor (the previous chunk's
                    // last tile; for the CTA's very first tile the initial ~0 carries the matching tag): publications
                    // then happen strictly in sequence order, which is what lets 32 prefix words serve any number of
                    // tiles -- tile q can only overwrite word q & 31 after tile q - 31 has read tile q - 32's totals.
                    HVX_JIT(4);
                    uint64_t base = 0;
#if defined(HVX_LEGACY_PROTOCOL) || defined(HVX_LEGACY_FIRST_TILE)  // stress builds: the round-1 form (a chunk's first tile does not wait)
                    if (!first)
#endif
                    {
                        const volatile uint64_t* prev = &sm.tile_prefix[(seq - 1u) & 31u];
                        const uint64_t want = static_cast<uint64_t>((seq - 1u) & 0xfffffu);
                        uint64_t got;
                        HVX_WAIT_BEGIN(WS_EMIT_CHAIN, seq);
                        do {
                            got = *prev;
                        } while ((got >> 44) != want);
                        HVX_WAIT_END();
                        // all lanes poll (a lane-0 spin leaves the warp split, and everything after it is issued twice),
                        // but the word can be overwritten 32 tiles later: the warp takes lane 0's copy, so a lane that
                        // looked late cannot carry a different prefix into the collectives below
                        got = __shfl_sync(0xffffffffu, got, 0);
                        if (!first) base = got & ((1ull << 44) - 1ull);
                    }
                    if (SPLIT && first && emit) {
                        // look-back across the parts of the chunk: this part starts where parts 0 .. q-1 end (their counting
                        // walks publish the totals)
                        const uint32_t item = e1.w & 0xffffffu, part = (e1.w >> 24) & 15u;
                        uint32_t bv = 0, bi = 0;  // one lane per earlier part: the look-back costs one round trip, not `part`
                        HVX_WAIT_BEGIN(WS_EMIT_LOOKBACK, part | (item << 8));
                        if (static_cast<uint32_t>(lane) < part) {
                            const uint4 t = wait_part(&p.item_totals[item - part + static_cast<uint32_t>(lane)], p.split_generation);
                            bv = t.x;
                            bi = t.y;
                        }
                        __syncwarp();
                        HVX_WAIT_END();
#pragma unroll
                        for (int d = 16; d != 0; d >>= 1) {
                            bv += __shfl_xor_sync(0xffffffffu, bv, d);
                            bi += __shfl_xor_sync(0xffffffffu, bi, d);
                        }
                        base = static_cast<uint64_t>(bv) | (static_cast<uint64_t>(bi) << 22);
                    }
                    HVX_CHECK((base & FIELD) + tot_v <= FIELD && ((base >> 22) & FIELD) + tot_i <= FIELD, 5u, chunk, st | (slot << 8), seq,
                              static_cast<uint32_t>(base), static_cast<uint32_t>(base >> 32));
                    HVX_JIT(5);
                    if (lane == 0) {
                        const uint64_t mine = (base + static_cast<uint64_t>(tot_v) + (static_cast<uint64_t>(tot_i) << 22)) |
                                              (static_cast<uint64_t>(seq & 0xfffffu) << 44);
                        if (t + 1 == ntiles) {  // totals so far; later steps overwrite, ordered along the chain
                            *const_cast<volatile uint64_t*>(&sm.chunk_total[kcpar]) = mine;
                            __threadfence_block();
                        }
                        *const_cast<volatile uint64_t*>(&sm.tile_prefix[seq & 31u]) = mine;
                    }
                    __syncwarp();
                    if (emit) {
                        const uint32_t v_base = static_cast<uint32_t>(base & FIELD);
                        const uint32_t i_base = static_cast<uint32_t>((base >> 22) & FIELD);
                        if (valid) {
                            const uint32_t nv = info & 15u, ni = 3u * ((info >> 4) & 15u), cls = info >> 8;
                            for (uint32_t q = 0; q < nv; ++q) ow[vo + q] = static_cast<uint8_t>(lane);
                            const uint32_t first_vertex = v_base + vo, dst = i_base + io;
                            const uint4 row4 = *reinterpret_cast<const uint4*>(&sm.class_index[cls * 16]);
                            const uint32_t words[4] = {row4.x, row4.y, row4.z, row4.w};
                            if (dst + ni <= p.max_indices) {
#pragma unroll
                                for (uint32_t q = 0; q < 15; ++q)
                                    if (q < ni) out_i[dst + q] = first_vertex + ((words[q >> 2] >> (8 * (q & 3))) & 0xffu);
                            } else {
                                for (uint32_t q = 0; q < ni; ++q)
                                    if (dst + q < p.max_indices)
                                        out_i[dst + q] = first_vertex + ((words[q >> 2] >> (8 * (q & 3))) & 0xffu);
                            }
                        }
                        HVX_JIT(6);
                        __syncwarp();
                        for (uint32_t v0 = 0; v0 < tot_v; v0 += 32u) {
                            const uint32_t v = v0 + static_cast<uint32_t>(lane);
                            const bool on = v < tot_v;
                            const uint32_t o = on ? ow[v] : 0u;
                            const uint32_t cr = __shfl_sync(0xffffffffu, rec, o);
                            const uint32_t cvv = __shfl_sync(0xffffffffu, vo_vb, o);
                            HVX_CHECK(!on || (v >= (cvv & 511u) && v - (cvv & 511u) < 12u && (cvv >> 9) + (v - (cvv & 511u)) < 1536u), 4u, chunk,
                                      st | (slot << 8), seq, v | (o << 16), cvv);
                            if (on && v_base + v < p.max_vertices) {
                                const int x = cr & 63, rw = (cr >> 8) & 255;
                                const int zl = rw / E, y = rw % E;
#ifdef HVX_SELFCHECK  // a broken invariant is logged above; stay inside the table so the log can be read
                                const uint32_t code = sm.vertex_packed[min((cvv >> 9) + (v - (cvv & 511u)), 1535u)];
#else
                                const uint32_t code = sm.vertex_packed[(cvv >> 9) + (v - (cvv & 511u))];
#endif
                                emit_regular_vertex_fast<C>(ring_flat, wl, x, y, zl, z0 + zl, code, tmask, out_v + v_base + v);
                            }
                        }
                    }
                    HVX_JIT(7);
                    __syncwarp();  // every ring read of the tile is done; the owner map is free again
                    if (lane == 0 && atomicAdd(&sm.q_done[qi], 1u) + 1u == ntiles) {
                        // last tile of the step: "step st done" for slabs st-1, st, st+1
                        mbar_arrive(&sm.empty_bar[slot_m]);
                        mbar_arrive(&sm.empty_bar[slot]);
                        if (st + 1 < C::NSLAB) mbar_arrive(&sm.empty_bar[slot_p]);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.q_free[qi]);
    }
}

using Cfg64 = Cfg<64, 2, 6, 16>;
using Cfg32 = Cfg<32, 2, 10, 8>;
using Cfg64D = Cfg<64, 1, 6, 20>;  // decoupled kernel at edge 64: 8 front + 20 emission + producer + scheduler warps
using Cfg32D = Cfg<32, 1, HVX_E32_RS, HVX_E32_NW>;  // decoupled kernel at edge 32: 4 front + 6 emission + producer + scheduler warps, 3 CTAs / SM

// One launch.  The shared-memory opt-in and the occupancy query are per device and do not change: they are made
// on a device's first launch of a kernel and remembered (two driver calls less on the single-page latency path).
template <class Kernel>
cudaError_t launch_persistent(Kernel* kernel, int* ctas_cache, int threads, size_t smem, const RegularParams& p,
                              const DeviceInfo& dev, cudaStream_t stream) {
    int ctas_per_sm = dev.ordinal >= 0 && dev.ordinal < 64 ? __atomic_load_n(&ctas_cache[dev.ordinal], __ATOMIC_ACQUIRE) : 0;
    if (ctas_per_sm == 0) {
        cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (err != cudaSuccess) return err;
        err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kernel, threads, smem);
        if (err != cudaSuccess) return err;
        if (ctas_per_sm < 1) return cudaErrorInvalidConfiguration;
        if (dev.ordinal >= 0 && dev.ordinal < 64) __atomic_store_n(&ctas_cache[dev.ordinal], ctas_per_sm, __ATOMIC_RELEASE);
    }
    const uint32_t grid = static_cast<uint32_t>(
        min(static_cast<long long>(p.n_work), static_cast<long long>(dev.sm_count) * ctas_per_sm));
    kernel<<<grid, threads, smem, stream>>>(p);
    return cudaGetLastError();
}

template <class C>
cudaError_t launch_first_generation(const RegularParams& p, const DeviceInfo& dev, cudaStream_t stream) {
    static_assert(sizeof(Smem<C>) <= 232448, "shared memory budget (227 KB per CTA)");
    static int ctas[64];
    return launch_persistent(regular_extract_kernel<C>, ctas, C::NT_ALL, sizeof(Smem<C>), p, dev, stream);
}

template <class C>
cudaError_t launch_decoupled(const RegularParams& p, const DeviceInfo& dev, cudaStream_t stream) {
    static_assert(sizeof(SmemD<C>) <= 232448, "shared memory budget (227 KB per CTA)");
    static int ctas[4][64];
    constexpr int NT = DecoupledCfg<C>::NT_ALL;
    if (p.items != nullptr) {
        if (p.any_partial) return launch_persistent(regular_extract_decoupled_kernel<C, true, true>, ctas[3], NT, sizeof(SmemD<C>), p, dev, stream);
        return launch_persistent(regular_extract_decoupled_kernel<C, false, true>, ctas[2], NT, sizeof(SmemD<C>), p, dev, stream);
    }
    if (p.any_partial) return launch_persistent(regular_extract_decoupled_kernel<C, true, false>, ctas[1], NT, sizeof(SmemD<C>), p, dev, stream);
    return launch_persistent(regular_extract_decoupled_kernel<C, false, false>, ctas[0], NT, sizeof(SmemD<C>), p, dev, stream);
}

}  // namespace

#define HVX_STR2(x) #x
#define HVX_STR(x) HVX_STR2(x)
const char* regular_kernel_name(int edge, bool first_generation, bool partial) {
    if (first_generation) return edge == 64 ? "regular_extract_kernel<Cfg<64,2,6,16>>" : "regular_extract_kernel<Cfg<32,2,10,8>>";
    if (edge == 64) return partial ? "regular_extract_decoupled_kernel<Cfg<64,1,6,20>,true,false>" : "regular_extract_decoupled_kernel<Cfg<64,1,6,20>,false,false>";
    return partial ? "regular_extract_decoupled_kernel<Cfg<32,1," HVX_STR(HVX_E32_RS) "," HVX_STR(HVX_E32_NW) ">,true,false>"
                   : "regular_extract_decoupled_kernel<Cfg<32,1," HVX_STR(HVX_E32_RS) "," HVX_STR(HVX_E32_NW) ">,false,false>";
}

size_t regular_smem_bytes(int edge) {
    return edge == 64 ? max(sizeof(Smem<Cfg64>), sizeof(SmemD<Cfg64D>)) : max(sizeof(Smem<Cfg32>), sizeof(SmemD<Cfg32D>));
}

cudaError_t launch_regular(int edge, const RegularParams& p, const DeviceInfo& dev, cudaStream_t stream) {
    if (p.n_chunks == 0 || p.n_work == 0) return cudaSuccess;
    cudaError_t e = cudaSuccess;  // the work counter rearms itself (rearm_work_counter)
    // The decoupled kernel is the product.  The first-generation kernel (CTA-wide barriers; writes the debug records
    // itself) only runs for contexts created with HVX_CFG_FIRST_GENERATION: the cross-check of the stress tests.
    if (p.first_generation) {
        if (edge == 64) return launch_first_generation<Cfg64>(p, dev, stream);
        if (edge == 32) return launch_first_generation<Cfg32>(p, dev, stream);
        return cudaErrorInvalidValue;
    }
    RegularParams q = p;
    q.cells = nullptr;  // per-cell records come from regular_records.cu
    q.offsets = nullptr;
    q.blocks = nullptr;
    if (edge == 64) e = launch_decoupled<Cfg64D>(q, dev, stream);
    else if (edge == 32) e = launch_decoupled<Cfg32D>(q, dev, stream);
    else return cudaErrorInvalidValue;
    if (e != cudaSuccess) return e;
    return launch_regular_records(edge, p, stream);
}

#ifdef HVX_WAITTRACE
// variant build only (tools/wait_trace.py): what every warp of every CTA is waiting for, readable while a launch hangs
// (the copy runs on its own non-blocking stream)
extern "C" int hvx_debug_wait_read(uint32_t* out /* [1024][32] */) {
    static cudaStream_t side = nullptr;
    if (!side && cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking) != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbolAsync(out, g_wait, sizeof(uint32_t) * 1024 * 32, 0, cudaMemcpyDeviceToHost, side) != cudaSuccess) return -2;
    return cudaStreamSynchronize(side) == cudaSuccess ? 0 : -3;
}
#endif

#ifdef HVX_TIMELINE
// variant build only (tools/timeline.py): the per-CTA time stamps of the decoupled kernel
extern "C" int hvx_debug_timeline_read(unsigned long long* out /* [1024][64] */, int reset) {
    if (cudaMemcpyFromSymbol(out, g_timeline, sizeof(unsigned long long) * 1024 * 64) != cudaSuccess) return -1;
    if (reset) {
        void* ptr = nullptr;
        if (cudaGetSymbolAddress(&ptr, g_timeline) != cudaSuccess || cudaMemset(ptr, 0, sizeof(unsigned long long) * 1024 * 64) != cudaSuccess) return -1;
    }
    return 0;
}
#endif

#ifdef HVX_SELFCHECK
// stress builds only (tools/repro_race.py): the invariant log of the decoupled kernel
extern "C" int hvx_debug_selfcheck_read(uint32_t* count, uint32_t* log /* [64][8] */, int reset) {
    if (cudaMemcpyFromSymbol(count, g_selfcheck_count, sizeof(uint32_t)) != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(log, g_selfcheck_log, sizeof(uint32_t) * 64 * 8) != cudaSuccess) return -1;
    if (reset) {
        const uint32_t zero = 0;
        if (cudaMemcpyToSymbol(g_selfcheck_count, &zero, sizeof zero) != cudaSuccess) return -1;
    }
    return 0;
}
#endif

}  // namespace hvx
