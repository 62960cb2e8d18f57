// transition_extract.cu -- Transvoxel transition (LOD seam) cells for sm_100a.
//
// Replaces classify_transition_cells / scan_transition_cells / scan_transition_blocks /
// emit_transition_cells (PV/src/transvoxel_transition_gpu.wgsl:173-454) with one kernel over a
// batch of chunks.  Arithmetic follows the reference's CPU extractor
// (PV/src/transvoxel_transition.rs:189-270, 412-424, 504-559), including the *0.5 on gradients
// that the WGSL omits, so results are bit-identical to the oracle.
//
// One CTA owns one chunk at a time (atomic queue) and walks its faces in index order.  A whole face
// is classified at once: every thread owns a row segment of CPT consecutive cells (v-major /
// u-fastest, the reference's face-local order), loads its 3 x (2 CPT + 1) layer-1 samples with all
// loads in flight, and one block scan per face gives the ordered placement -- the running prefix
// inside the CTA reproduces the reference's face-major order without any cross-CTA scan.  Vertices
// are then emitted one thread per vertex (owner thread by binary search over the scanned prefix).
// Layer 1 of a slab is read once, coalesced; layers 0 and 2 are touched only around active cells
// for the outward gradient.  Slab rows are 4*(2E+3)
// bytes, not 16-byte multiples, so this path uses plain vector-width-1 loads through L1 rather
// than bulk copies.
#include "hvx_device.cuh"
#include "hvx_kernels.h"

namespace hvx {

namespace {

// How many vertices of a face a thread has in flight (14 sample loads each, one L2 / HBM round trip per trip of the
// loop).  Measured 1 / 2 / 4: 1024 coarse 64^3 pages 0.190 / 0.191 / 0.190 ms, the planet set's faces 0.184 / 0.188 /
// 0.192 ms -- the occupancy already hides the round trips; kept at 1.
#ifndef HVX_T_UNROLL
#define HVX_T_UNROLL 1
#endif
constexpr int T_UNROLL = HVX_T_UNROLL;
#define HVX_TABLE alignas(16) static __device__ const
#include "transvoxel_tables.inc"

// PV/src/transvoxel_transition.rs:463-502 integer face bases: origin, u, v, outward.
struct FaceBasis {
    int origin[3], u[3], v[3], o[3];
};
__constant__ FaceBasis c_face_basis[6] = {
    {{0, 0, 1}, {0, 1, 0}, {0, 0, -1}, {-1, 0, 0}},
    {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 0}},
    {{1, 0, 0}, {0, 0, 1}, {-1, 0, 0}, {0, -1, 0}},
    {{0, 1, 0}, {0, 0, 1}, {1, 0, 0}, {0, 1, 0}},
    {{0, 1, 0}, {1, 0, 0}, {0, -1, 0}, {0, 0, -1}},
    {{0, 0, 1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}},
};

template <int E_, int NT_>
struct TCfg {
    static constexpr int E = E_, NT = NT_;
    static constexpr int W = 2 * E_ + 3;             // slab edge
    static constexpr int FACE_WORDS = W * W * 3;     // one face slab
    static constexpr int FACE_CELLS = E_ * E_;
    static constexpr int CPT = FACE_CELLS / NT_;     // consecutive cells (one row segment) per thread
    static constexpr int RW = 2 * E_ + 1;            // layer-1 samples a cell row touches (columns 1 .. 2E+1)
    static constexpr int RW32 = (RW + 31) / 32;      // ... in 32-sample ballot words
    static_assert(32 * CPT == 4 * E_, "a warp owns exactly four cell rows");
    static_assert(FACE_CELLS % NT_ == 0 && E_ % CPT == 0 && CPT <= 8 && 256 % CPT == 0, "a thread owns a row segment");
};

template <class C>
struct TSmem {
    uint64_t thread_pre[C::NT + 1];  // exclusive prefix per thread: vertices | indices<<16 | active cells<<34
    uint64_t scan_sums[40], scan_prefix[40];
    uint32_t thread_nv[C::NT];       // vertex counts of a thread's CPT cells, 4 bits each
    uint32_t thread_nt[C::NT];       // triangle counts, 4 bits each
    uint32_t rowbits[C::NT / 32][9][C::RW32 + 1];  // per warp: solid bits of its 9 layer-1 sample rows
    uint16_t cases[C::FACE_CELLS];   // 9-bit case of every cell of the face being processed
    uint32_t chunk_id;
    alignas(16) uint16_t case_info[512];     // the three tables are copied 16 bytes at a time
    alignas(16) uint8_t vertex_edge[512 * 12];
    alignas(16) uint8_t class_index[56 * 36];
};

template <class C>
__device__ __forceinline__ void emit_transition_vertex(const uint32_t* __restrict__ slab, int face, int cu, int cv,
                                                       uint32_t code, hvx_vertex* dst) {
    constexpr int W = C::W;
    const FaceBasis& b = c_face_basis[face];
    const int c0 = code >> 4, c1 = code & 15;
    // half-resolution corners 9..C duplicate full-resolution 0, 2, 6, 8 (transvoxel.rs:26)
    const int f0 = c0 < 9 ? c0 : (c0 == 9 ? 0 : c0 == 10 ? 2 : c0 == 11 ? 6 : 8);
    const int f1 = c1 < 9 ? c1 : (c1 == 9 ? 0 : c1 == 10 ? 2 : c1 == 11 ? 6 : 8);
    const int du0 = f0 % 3, dv0 = f0 / 3, du1 = f1 % 3, dv1 = f1 / 3;
    const int su0 = 2 * cu + du0 + 1, sv0 = 2 * cv + dv0 + 1;
    const int su1 = 2 * cu + du1 + 1, sv1 = 2 * cv + dv1 + 1;
    auto word = [&](int su, int sv, int layer) -> uint32_t { return __ldg(slab + su + sv * W + layer * W * W); };
    const uint32_t w0 = word(su0, sv0, 1), w1 = word(su1, sv1, 1);
    const float d0 = cw_density(w0), d1 = cw_density(w1);
    const float t = edge_parameter(d0, d1);
    // xyz gradient: axis a lies along exactly one of u, v, outward with sign +-1
    // (transvoxel_transition.rs:412-424 evaluated on the slab halo)
    float g[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int ddu = b.u[a], ddv = b.v[a], ddl = b.o[a];
        const float g0 = fmul(fsub(cw_density(word(su0 + ddu, sv0 + ddv, 1 + ddl)),
                                   cw_density(word(su0 - ddu, sv0 - ddv, 1 - ddl))), 0.5f);
        const float g1 = fmul(fsub(cw_density(word(su1 + ddu, sv1 + ddv, 1 + ddl)),
                                   cw_density(word(su1 - ddu, sv1 - ddv, 1 - ddl))), 0.5f);
        g[a] = fmix(g0, g1, t);
    }
    const float fu = fadd(static_cast<float>(cu), fmix(fmul(static_cast<float>(du0), 0.5f), fmul(static_cast<float>(du1), 0.5f), t));
    const float fv = fadd(static_cast<float>(cv), fmix(fmul(static_cast<float>(dv0), 0.5f), fmul(static_cast<float>(dv1), 0.5f), t));
    const float depth = fmix(c0 < 9 ? 0.0f : 1.0f, c1 < 9 ? 0.0f : 1.0f, t);
    float n[3];
    const float s = fadd(fadd(fmul(g[0], g[0]), fmul(g[1], g[1])), fmul(g[2], g[2]));
    if (s > 1.0e-12f) {
        const float inv = fdiv(1.0f, fsqrt(s));
#pragma unroll
        for (int a = 0; a < 3; ++a) n[a] = fmul(g[a], inv);
    } else {
#pragma unroll
        for (int a = 0; a < 3; ++a) n[a] = static_cast<float>(b.o[a]);
    }
    float primary[3], inward[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float ua = static_cast<float>(b.u[a]), va = static_cast<float>(b.v[a]), oa = static_cast<float>(b.o[a]);
        const float origin = fmul(static_cast<float>(b.origin[a]), static_cast<float>(C::E));
        primary[a] = fadd(origin, fsub(fadd(fmul(ua, fu), fmul(va, fv)), fmul(oa, 0.0f)));
        inward[a] = fmul(fmul(-oa, 0.25f), depth);
    }
    const float nc = fadd(fadd(fmul(inward[0], n[0]), fmul(inward[1], n[1])), fmul(inward[2], n[2]));
    float p[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) p[a] = fadd(primary[a], fsub(inward[a], fmul(n[a], nc)));
    const uint32_t material = d0 <= 0.0f ? cw_material(w0) : cw_material(w1);
    float4* out = reinterpret_cast<float4*>(dst);
    out[0] = make_float4(p[0], p[1], p[2], __uint_as_float(material));
    out[1] = make_float4(n[0], n[1], n[2], __uint_as_float(1u << face));
}

template <class C>
__global__ void __launch_bounds__(C::NT) transition_extract_kernel(const TransitionParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TSmem<C>& sm = *reinterpret_cast<TSmem<C>*>(smem_raw);
    constexpr int E = C::E, W = C::W, NT = C::NT, CPT = C::CPT, NS = 2 * CPT + 1;
    const int tid = threadIdx.x;
    // Lengyel's transition tables, 9 KB, 16 bytes per load (every CTA of a small batch pays this before its first face)
    static_assert(sizeof(sm.case_info) % 16 == 0 && sizeof(sm.vertex_edge) % 16 == 0 && sizeof(sm.class_index) % 16 == 0, "table copy granularity");
    for (int i = tid; i < static_cast<int>(sizeof(sm.case_info) / 16); i += NT)
        reinterpret_cast<uint4*>(sm.case_info)[i] = reinterpret_cast<const uint4*>(HVX_TRANSITION_CASE_INFO)[i];
    for (int i = tid; i < static_cast<int>(sizeof(sm.vertex_edge) / 16); i += NT)
        reinterpret_cast<uint4*>(sm.vertex_edge)[i] = reinterpret_cast<const uint4*>(&HVX_TRANSITION_VERTEX_EDGE[0][0])[i];
    for (int i = tid; i < static_cast<int>(sizeof(sm.class_index) / 16); i += NT)
        reinterpret_cast<uint4*>(sm.class_index)[i] = reinterpret_cast<const uint4*>(&HVX_TRANSITION_CLASS_INDEX[0][0])[i];
    // this thread's CPT cells of a face: one row segment, u fastest (face-local linear order)
    const int cell0 = tid * CPT;

    for (;;) {
        __syncthreads();
        if (tid == 0) sm.chunk_id = atomicAdd(p.work_counter, 1u);
        __syncthreads();
        const uint32_t chunk = sm.chunk_id;
        if (chunk >= p.n_chunks) {
            if (tid == 0) rearm_work_counter(p.work_counter);  // every CTA draws exactly one ticket past the end
            break;
        }
        const ChunkDesc desc = p.descs[chunk];
        const uint32_t mask = desc.transition_mask & 0x3fu;
        const uint32_t glo = static_cast<uint32_t>(desc.generation), ghi = static_cast<uint32_t>(desc.generation >> 32);
        const uint32_t* chunk_slabs = p.slabs + static_cast<size_t>(chunk) * 6 * C::FACE_WORDS;
        hvx_vertex* const out_v = p.vertices + static_cast<size_t>(chunk) * p.max_vertices;
        uint32_t* const out_i = p.indices + static_cast<size_t>(chunk) * p.max_indices;
        const bool debug = p.cells != nullptr;
        const size_t cell_base = static_cast<size_t>(chunk) * 6 * C::FACE_CELLS;
        const size_t block_base = static_cast<size_t>(chunk) * (6 * C::FACE_CELLS / 256);

        uint32_t v_base = 0, i_base = 0, active_cells = 0;
        for (int face = 0; face < 6; ++face) {
            if (!((mask >> face) & 1u)) {
                if (debug) {
                    // inactive faces publish zero records (transvoxel_transition_gpu.wgsl:183-186)
                    for (int c = tid; c < C::FACE_CELLS; c += NT)
                        *reinterpret_cast<uint4*>(&p.cells[cell_base + face * C::FACE_CELLS + c]) = make_uint4(0, 0, 0, 0);
                    for (int bl = tid; bl < C::FACE_CELLS / 256; bl += NT) {
                        hvx_scan_block blk = {0u, 0u, v_base, i_base};
                        p.blocks[block_base + face * (C::FACE_CELLS / 256) + bl] = blk;
                    }
                }
                continue;
            }
            const uint32_t* slab = chunk_slabs + static_cast<size_t>(face) * C::FACE_WORDS;
            // ---- classify the whole face.  A warp owns four cell rows = nine layer-1 sample rows; it
            //      loads them coalesced (lane = column, all loads in flight), one ballot per 32 samples
            //      turns them into bit rows in shared memory, and a cell's 9 signs are 3-bit windows of
            //      three bit rows ------------------------------------------------------------------
            uint32_t rows[3];
            {
                const int warp = tid >> 5, lane = tid & 31;
                const uint32_t* base = slab + W * W + (2 * (warp * 4) + 1) * W + 1;
                uint32_t w[9][C::RW32];
#pragma unroll
                for (int r = 0; r < 9; ++r)
#pragma unroll
                    for (int c = 0; c < C::RW32; ++c) {
                        const int col = 32 * c + lane;
                        w[r][c] = col < C::RW ? __ldg(base + r * W + col) : 0x7fffu;  // padding reads as air
                    }
#pragma unroll
                for (int r = 0; r < 9; ++r)
#pragma unroll
                    for (int c = 0; c < C::RW32; ++c) {
                        const uint32_t b = __ballot_sync(0xffffffffu, cw_solid(w[r][c]));
                        if (lane == 0) sm.rowbits[warp][r][c] = b;
                    }
                if (lane < 9) sm.rowbits[warp][lane][C::RW32] = 0u;
                __syncwarp();
                const int lv = lane / (E / CPT), bit = 2 * CPT * (lane % (E / CPT));
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const uint32_t* rb = sm.rowbits[warp][2 * lv + r] + (bit >> 5);
                    rows[r] = __funnelshift_r(rb[0], rb[1], bit & 31) & ((1u << NS) - 1u);
                }
                __syncwarp();
            }
            uint32_t cases[CPT], nv_word = 0, nt_word = 0, tot_v = 0, tot_i = 0, tot_a = 0;
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                const uint32_t t0 = (rows[0] >> (2 * j)) & 7u, t1 = (rows[1] >> (2 * j)) & 7u, t2 = (rows[2] >> (2 * j)) & 7u;
                // case weights 1,2,4 / 0x80,0x100,8 / 0x40,0x20,0x10 (PV/src/transvoxel.rs:20-22)
                const uint32_t c = t0 | ((t1 & 3u) << 7) | ((t1 & 4u) << 1) | ((t2 & 1u) << 6) | ((t2 & 2u) << 4) | ((t2 & 4u) << 2);
                cases[j] = c;
                const uint32_t info = sm.case_info[c];
                const uint32_t nv = info & 15u, nt = (info >> 4) & 15u;
                nv_word |= nv << (4 * j);
                nt_word |= nt << (4 * j);
                tot_v += nv;
                tot_i += 3u * nt;
                tot_a += nv != 0u ? 1u : 0u;
                sm.cases[cell0 + j] = static_cast<uint16_t>(c);
            }
            sm.thread_nv[tid] = nv_word;
            sm.thread_nt[tid] = nt_word;
            uint64_t face_tot;
            const uint64_t mine = static_cast<uint64_t>(tot_v) | (static_cast<uint64_t>(tot_i) << 16) | (static_cast<uint64_t>(tot_a) << 34);
            const uint64_t pre = block_exclusive_scan<NT>(mine, sm.scan_sums, sm.scan_prefix, face_tot);
            sm.thread_pre[tid] = pre;
            if (tid == 0) sm.thread_pre[NT] = face_tot;
            const uint32_t face_v = static_cast<uint32_t>(face_tot & 0xffffu), face_i = static_cast<uint32_t>((face_tot >> 16) & 0x3ffffu);
            active_cells += static_cast<uint32_t>(face_tot >> 34);
            __syncthreads();
            // ---- debug records: each thread walks its own cells --------------------------------------
            if (debug) {
                uint32_t vo = static_cast<uint32_t>(pre & 0xffffu), io = static_cast<uint32_t>((pre >> 16) & 0x3ffffu);
                const uint64_t bpre = sm.thread_pre[(tid / (256 / CPT)) * (256 / CPT)];  // start of this cell's 256-cell scan block
                const uint32_t bvo = static_cast<uint32_t>(bpre & 0xffffu), bio = static_cast<uint32_t>((bpre >> 16) & 0x3ffffu);
#pragma unroll
                for (int j = 0; j < CPT; ++j) {
                    const uint32_t info = sm.case_info[cases[j]];
                    const uint32_t nv = info & 15u, nt = (info >> 4) & 15u, class_code = info >> 8;
                    const size_t lin = cell_base + static_cast<size_t>(face) * C::FACE_CELLS + cell0 + j;
                    *reinterpret_cast<uint4*>(&p.cells[lin]) =
                        make_uint4(cases[j] | (class_code << 9) | (nv << 17) | (nt << 21) | 0x80000000u, glo, ghi, 0u);
                    *reinterpret_cast<uint4*>(&p.offsets[lin]) = make_uint4(vo - bvo, io - bio, glo, ghi);
                    vo += nv;
                    io += 3u * nt;
                }
                if ((cell0 % 256) == 0) {
                    const uint64_t nxt = sm.thread_pre[tid + 256 / CPT];
                    hvx_scan_block blk;
                    blk.vertex_count = static_cast<uint32_t>(nxt & 0xffffu) - bvo;
                    blk.index_count = static_cast<uint32_t>((nxt >> 16) & 0x3ffffu) - bio;
                    blk.first_vertex = v_base + bvo;
                    blk.first_index = i_base + bio;
                    p.blocks[block_base + (face * C::FACE_CELLS + cell0) / 256] = blk;
                }
            }
            // ---- indices: one thread per TRIANGLE (owner thread by binary search over the index prefix,
            //      owner cell by walking that thread's 4-bit counts) ------------------------------------
            for (uint32_t tr = tid; 3u * tr < face_i; tr += NT) {
                const uint32_t want = 3u * tr;
                uint32_t lo = 0;
#pragma unroll
                for (uint32_t step = NT / 2; step != 0; step >>= 1)
                    if (static_cast<uint32_t>((sm.thread_pre[lo + step] >> 16) & 0x3ffffu) <= want) lo += step;
                const uint64_t tp = sm.thread_pre[lo];
                uint32_t k = tr - static_cast<uint32_t>((tp >> 16) & 0x3ffffu) / 3u;  // triangle within the thread's cells
                uint32_t first_vertex = v_base + static_cast<uint32_t>(tp & 0xffffu);
                uint32_t ntw = sm.thread_nt[lo], nvw = sm.thread_nv[lo];
                uint32_t j = 0;
                while (k >= (ntw & 15u)) {
                    k -= ntw & 15u;
                    first_vertex += nvw & 15u;
                    ntw >>= 4;
                    nvw >>= 4;
                    ++j;
                }
                const uint32_t class_code = sm.case_info[sm.cases[lo * CPT + j]] >> 8;
                // per-case inverse flip then the global flip => raw order iff inverse
                const uint8_t* tri = &sm.class_index[(class_code & 0x7fu) * 36 + 3u * k];
                const uint32_t a = tri[0], b1 = tri[1], c1 = tri[2];
                const bool inverse = (class_code & 0x80u) != 0;
                const uint32_t d = i_base + want;
                if (d + 2 < p.max_indices) {
                    out_i[d] = first_vertex + a;
                    out_i[d + 1] = first_vertex + (inverse ? b1 : c1);
                    out_i[d + 2] = first_vertex + (inverse ? c1 : b1);
                }
            }
            // ---- vertices: one thread per vertex; owner thread by binary search over the prefix,
            //      owner cell by walking that thread's 4-bit vertex counts ---------------------------
#pragma unroll T_UNROLL
            for (uint32_t v = tid; v < face_v; v += NT) {
                uint32_t lo = 0;
#pragma unroll
                for (uint32_t step = NT / 2; step != 0; step >>= 1)
                    if (static_cast<uint32_t>(sm.thread_pre[lo + step] & 0xffffu) <= v) lo += step;
                uint32_t k = v - static_cast<uint32_t>(sm.thread_pre[lo] & 0xffffu);
                uint32_t nvw = sm.thread_nv[lo];
                uint32_t j = 0;
                while (k >= (nvw & 15u)) {
                    k -= nvw & 15u;
                    nvw >>= 4;
                    ++j;
                }
                const uint32_t cell = lo * CPT + j;
                const uint32_t cc = sm.cases[cell];
                if (v_base + v < p.max_vertices)
                    emit_transition_vertex<C>(slab, face, static_cast<int>(cell % E), static_cast<int>(cell / E),
                                              sm.vertex_edge[cc * 12 + k], out_v + v_base + v);
            }
            v_base += face_v;
            i_base += face_i;
            __syncthreads();  // thread_pre / thread_nv / cases are rewritten by the next face
        }
        if (tid == 0) {
            const uint32_t vo = v_base > p.max_vertices ? 1u : 0u, io = i_base > p.max_indices ? 1u : 0u;
            const bool ok = !(vo | io);
            hvx_transition_counters tc;
            tc.active_cells = active_cells;
            tc.active_faces = __popc(mask);
            tc.required_vertices = v_base;
            tc.required_indices = i_base;
            tc.emitted_vertices = ok ? v_base : 0u;
            tc.emitted_indices = ok ? i_base : 0u;
            tc.vertex_overflow = vo;
            tc.index_overflow = io;
            tc.completed = 1u;
            tc._pad[0] = tc._pad[1] = tc._pad[2] = 0u;
            p.counters[chunk] = tc;
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
cudaError_t launch_tcfg(const TransitionParams& p, const DeviceInfo& dev, cudaStream_t stream) {
    const size_t smem = sizeof(TSmem<C>);
    // the shared-memory opt-in and the occupancy are per device and do not change: asked once (as launch_persistent does)
    static int cache[64];
    int per_sm = dev.ordinal >= 0 && dev.ordinal < 64 ? __atomic_load_n(&cache[dev.ordinal], __ATOMIC_ACQUIRE) : 0;
    if (per_sm == 0) {
        cudaError_t err = cudaFuncSetAttribute(transition_extract_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               static_cast<int>(smem));
        if (err != cudaSuccess) return err;
        err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, transition_extract_kernel<C>, C::NT, smem);
        if (err != cudaSuccess) return err;
        if (per_sm < 1) return cudaErrorInvalidConfiguration;
        if (dev.ordinal >= 0 && dev.ordinal < 64) __atomic_store_n(&cache[dev.ordinal], per_sm, __ATOMIC_RELEASE);
    }
    const uint32_t grid = static_cast<uint32_t>(
        min(static_cast<long long>(p.n_chunks), static_cast<long long>(dev.sm_count) * per_sm));
    transition_extract_kernel<C><<<grid, C::NT, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_transition(int edge, const TransitionParams& p, const DeviceInfo& dev, cudaStream_t stream) {
    if (p.n_chunks == 0) return cudaSuccess;
    // no memset: the work counter rearms itself (work_counter[0] tickets, work_counter[2] CTAs that have left)
    if (edge == 64) return launch_tcfg<TCfg<64, 512>>(p, dev, stream);
    if (edge == 32) return launch_tcfg<TCfg<32, 256>>(p, dev, stream);
    return cudaErrorInvalidValue;
}

}  // namespace hvx
