// transition_extract.cu -- Transvoxel transition (LOD seam) cells for sm_100a.
//
// Replaces classify_transition_cells / scan_transition_cells / scan_transition_blocks /
// emit_transition_cells (PV/src/transvoxel_transition_gpu.wgsl:173-454) with one kernel over a
// batch of chunks.  Arithmetic follows the reference's CPU extractor
// (PV/src/transvoxel_transition.rs:189-270, 412-424, 504-559), including the *0.5 on gradients
// that the WGSL omits, so results are bit-identical to the oracle.
//
// One CTA owns one chunk at a time (atomic queue) and walks its faces in index order, each face
// v-major / u-fastest in steps of NT cells, so the running prefix inside the CTA reproduces the
// reference's face-major order without any cross-CTA scan.  Layer 1 of a slab is read once,
// coalesced (a warp's 9-sample patches cover three contiguous 65-word row segments); layers 0
// and 2 are touched only around active cells for the outward gradient.  Slab rows are 4*(2E+3)
// bytes, not 16-byte multiples, so this path uses plain vector-width-1 loads through L1 rather
// than bulk copies.
#include "hvx_device.cuh"
#include "hvx_kernels.h"

namespace hvx {

namespace {

#define HVX_TABLE static __device__ const
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
    static constexpr int STEPS = FACE_CELLS / NT_;   // NT consecutive cells (v-major) per step
    static_assert(FACE_CELLS % NT_ == 0 && NT_ % 256 == 0, "steps must align with 256-cell scan blocks");
};

template <class C>
struct TSmem {
    uint32_t cell_rec[C::NT];        // case | u<<9 | v<<16   (u, v < 64)
    uint32_t cell_off[C::NT];        // step-local exclusive vertex | index<<16
    uint16_t owner[C::NT * 12];      // vertex -> cell slot | k<<10
    uint32_t scan_sums[40], scan_prefix[40];
    uint32_t chunk_id;
    uint16_t case_info[512];
    uint8_t vertex_edge[512 * 12];
    uint8_t class_index[56 * 36];
};

// 9-bit case weights, PV/src/transvoxel.rs:20-22 (not row-major after sample 2)
__device__ __forceinline__ uint32_t case_weight(int i) {
    const uint32_t w[9] = {0x001, 0x002, 0x004, 0x080, 0x100, 0x008, 0x040, 0x020, 0x010};
    return w[i];
}

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
    constexpr int E = C::E, W = C::W, NT = C::NT;
    const int tid = threadIdx.x;
    for (int i = tid; i < 512; i += NT) sm.case_info[i] = HVX_TRANSITION_CASE_INFO[i];
    for (int i = tid; i < 512 * 12; i += NT) sm.vertex_edge[i] = HVX_TRANSITION_VERTEX_EDGE[i / 12][i % 12];
    for (int i = tid; i < 56 * 36; i += NT) sm.class_index[i] = HVX_TRANSITION_CLASS_INDEX[i / 36][i % 36];

    for (;;) {
        __syncthreads();
        if (tid == 0) sm.chunk_id = atomicAdd(p.work_counter, 1u);
        __syncthreads();
        const uint32_t chunk = sm.chunk_id;
        if (chunk >= p.n_chunks) break;
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
            for (int step = 0; step < C::STEPS; ++step) {
                const int cell = step * NT + tid;  // face-local linear, u fastest
                const int cu = cell % E, cv = cell / E;
                // ---- classify: 9 samples of layer 1 ------------------------------------------
                uint32_t c = 0;
                {
                    const uint32_t* row = slab + W * W + (2 * cv + 1) * W + (2 * cu + 1);
#pragma unroll
                    for (int i = 0; i < 9; ++i)
                        if (cw_solid(__ldg(row + (i / 3) * W + (i % 3)))) c |= case_weight(i);
                }
                const uint32_t info = sm.case_info[c];
                const uint32_t nv = info & 15u, nt = (info >> 4) & 15u, class_code = info >> 8;
                uint32_t tot;
                const uint32_t off = block_exclusive_scan<NT>(nv | ((3u * nt) << 16), sm.scan_sums, sm.scan_prefix, tot);
                const uint32_t step_v = tot & 0xffffu, step_i = tot >> 16;
                const uint32_t vo = off & 0xffffu, io = off >> 16;
                const uint32_t n_act = __syncthreads_count(nv != 0);
                active_cells += n_act;
                sm.cell_rec[tid] = c | (cu << 9) | (cv << 16);
                sm.cell_off[tid] = off;
                for (uint32_t k = 0; k < nv; ++k) sm.owner[vo + k] = static_cast<uint16_t>(tid | (k << 10));
                // ---- indices: per-case inverse flip then the global flip => raw order iff inverse
                {
                    const uint32_t first_vertex = v_base + vo, dst = i_base + io;
                    const uint8_t* tri = &sm.class_index[(class_code & 0x7fu) * 36];
                    const bool inverse = (class_code & 0x80u) != 0;
                    for (uint32_t tr = 0; tr < nt; ++tr) {
                        const uint32_t a = tri[3 * tr], b1 = tri[3 * tr + 1], c1 = tri[3 * tr + 2];
                        const uint32_t second = inverse ? b1 : c1, third = inverse ? c1 : b1;
                        const uint32_t d = dst + 3 * tr;
                        if (d + 2 < p.max_indices) {
                            out_i[d] = first_vertex + a;
                            out_i[d + 1] = first_vertex + second;
                            out_i[d + 2] = first_vertex + third;
                        }
                    }
                }
                __syncthreads();
                if (debug) {
                    const size_t lin = cell_base + static_cast<size_t>(face) * C::FACE_CELLS + cell;
                    *reinterpret_cast<uint4*>(&p.cells[lin]) =
                        make_uint4(c | (class_code << 9) | (nv << 17) | (nt << 21) | 0x80000000u, glo, ghi, 0u);
                    const uint32_t boff = sm.cell_off[(tid / 256) * 256];
                    *reinterpret_cast<uint4*>(&p.offsets[lin]) =
                        make_uint4(vo - (boff & 0xffffu), io - (boff >> 16), glo, ghi);
                    if (tid % 256 == 0) {
                        const uint32_t nxt = tid + 256 < NT ? sm.cell_off[tid + 256] : tot;
                        hvx_scan_block blk;
                        blk.vertex_count = (nxt & 0xffffu) - (boff & 0xffffu);
                        blk.index_count = (nxt >> 16) - (boff >> 16);
                        blk.first_vertex = v_base + (boff & 0xffffu);
                        blk.first_index = i_base + (boff >> 16);
                        p.blocks[block_base + (face * C::FACE_CELLS + cell) / 256] = blk;
                    }
                }
                // ---- vertices: one thread per vertex -------------------------------------------
                for (uint32_t v = tid; v < step_v; v += NT) {
                    const uint32_t o = sm.owner[v];
                    const uint32_t slot = o & 1023u, k = o >> 10;
                    const uint32_t cr = sm.cell_rec[slot];
                    const uint32_t cc = cr & 511u;
                    const int u2 = (cr >> 9) & 127, v2 = cr >> 16;
                    if (v_base + v < p.max_vertices)
                        emit_transition_vertex<C>(slab, face, u2, v2, sm.vertex_edge[cc * 12 + k], out_v + v_base + v);
                }
                v_base += step_v;
                i_base += step_i;
                __syncthreads();
            }
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
    cudaError_t err = cudaFuncSetAttribute(transition_extract_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem));
    if (err != cudaSuccess) return err;
    int per_sm = 1;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, transition_extract_kernel<C>, C::NT, smem);
    if (err != cudaSuccess) return err;
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    const uint32_t grid = static_cast<uint32_t>(
        min(static_cast<long long>(p.n_chunks), static_cast<long long>(dev.sm_count) * per_sm));
    transition_extract_kernel<C><<<grid, C::NT, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_transition(int edge, const TransitionParams& p, const DeviceInfo& dev, cudaStream_t stream) {
    if (p.n_chunks == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(p.work_counter, 0, sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    if (edge == 64) return launch_tcfg<TCfg<64, 512>>(p, dev, stream);
    if (edge == 32) return launch_tcfg<TCfg<32, 256>>(p, dev, stream);
    return cudaErrorInvalidValue;
}

}  // namespace hvx
