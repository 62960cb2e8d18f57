// GPU check of the header-only C++ mirror (include/helio_voxel_cuda.hpp): the reference's sphere
// fixture must give 1,323 vertices / 661 triangles (docs/planetary_voxel_extraction_benchmark.md:61),
// the error enum must map like the Rust one.  Built and run by tests/test_gpu_cpp_mirror.py.
#include <cstdio>
#include <vector>

#include "helio_voxel_cuda.hpp"

using namespace helio_voxel_cuda;

static uint32_t cellword(int density, uint32_t material) { return (uint32_t(density) & 0xffffu) | (material << 16); }

int main() {
    std::vector<uint32_t> samples(34 * 34 * 34);
    size_t i = 0;
    for (int z = -1; z <= 32; ++z)
        for (int y = -1; y <= 32; ++y)
            for (int x = -1; x <= 32; ++x) {
                long d = long(x) * x + long(y) * y + long(z) * z - 144;  // ExtractionFixtureKind::Sphere
                d = d < -32768 ? -32768 : d > 32767 ? 32767 : d;
                samples[i++] = cellword(int(d), d <= 0 ? 1u : 0u);
            }
    TransvoxelGpuExtractor extractor(0);
    extractor.dispatch(samples.data(), samples.size(), 1000, ~0ull, 0);
    const hvx_emission_counters c = extractor.counters_buffer();
    if (c.completed != 1 || c.emitted_vertices != 1323 || c.emitted_indices != 661 * 3) {
        std::printf("FAIL counters %u %u %u\n", c.completed, c.emitted_vertices, c.emitted_indices);
        return 1;
    }
    const auto v = extractor.vertices_buffer(c.emitted_vertices);
    if (v[0].position[0] != 12.0f || v[0].normal[0] != 1.0f || v[0].material != 1) {
        std::printf("FAIL first vertex\n");
        return 1;
    }
    try {
        extractor.dispatch(samples.data(), samples.size() - 1, 1, ~0ull, 0);
        std::printf("FAIL no SampleCount error\n");
        return 1;
    } catch (const Error& e) {
        if (!e.is_sample_count()) return 1;
    }
    try {
        TransvoxelGpuExtractorConfig::create(0, 1);
        return 1;
    } catch (const Error& e) {
        if (!e.is_invalid_extraction_capacity()) return 1;
    }
    TransvoxelGpuExtractor tiny(0, TransvoxelGpuExtractorConfig::create(1, 1));
    tiny.dispatch(samples.data(), samples.size(), 2000, ~0ull, 0);
    const hvx_emission_counters t = tiny.counters_buffer();
    if (!(t.completed == 1 && t.vertex_overflow && t.index_overflow && t.emitted_vertices == 0 && t.required_vertices == 1323)) return 1;
    // the batch form: fill three terrain chunks on the device, pipelined extraction + packed read-back, one sphere
    // edit on the resident samples and the dirty re-extraction it asks for
    {
        ChunkBatchExtractor batch(0, 64, 3);
        const int64_t pages[9] = {0, -1, 0, 1, -1, 0, 0, 4, 0};
        batch.fill(16, pages, nullptr, 3);
        std::vector<ChunkBatchExtractor::Request> requests(3);
        requests[2].uniform = true;  // far above the terrain: the producer knows it is empty
        const auto descs = batch.prepare(requests);
        const ChunkBatchExtractor::Meshes m = batch.extract_to_host(descs, nullptr, 3 * 49152, 3 * 73728);
        if (m.counters[0].emitted_vertices == 0 || m.counters[2].emitted_vertices != 0 || m.counters[2].completed != 1 ||
            m.vertices.size() != size_t(m.counters[0].emitted_vertices) + m.counters[1].emitted_vertices ||
            m.ranges[1].first_vertex != m.counters[0].emitted_vertices) {
            std::printf("FAIL batch read-back\n");
            return 1;
        }
        hvx_voxel_edit edit{0, 2, 0, {3.0f, -4.3f, 3.0f}, 0.8f, 0};  // SubtractSphere around the terrain surface
        uint32_t touched = 0;
        const std::vector<uint64_t> dirty = batch.apply_edit(edit, pages, nullptr, 3, &touched);
        if (touched != 1 || dirty[0] == 0 || dirty[1] != 0 || dirty[2] != 0) {
            std::printf("FAIL edit dirty set\n");
            return 1;
        }
        for (size_t k = 0; k < 3; ++k) requests[k].dirty_microbricks = dirty[k];
        requests[2].uniform = false;
        batch.encode(batch.prepare(requests));
        const auto after = batch.counters_buffer(3);
        if (after[0].completed != 1 || after[0].vertex_overflow || after[1].emitted_vertices != 0 || after[2].emitted_vertices != 0 || after[2].completed != 1) {
            std::printf("FAIL dirty re-extraction\n");
            return 1;
        }
        // optional vertex-reuse output: fewer vertices, the same index count, and a second weld is a no-op
        for (size_t k = 0; k < 3; ++k) requests[k].dirty_microbricks = ~0ull;
        batch.encode(batch.prepare(requests));
        const auto unshared = batch.counters_buffer(3);
        batch.weld_meshes(3);
        const auto shared = batch.counters_buffer(3);
        batch.weld_meshes(3);
        const auto again = batch.counters_buffer(3);
        if (unshared[0].emitted_vertices == 0 || shared[0].emitted_vertices * 2 >= unshared[0].emitted_vertices ||
            shared[0].emitted_indices != unshared[0].emitted_indices || shared[0].required_vertices != unshared[0].required_vertices ||
            again[0].emitted_vertices != shared[0].emitted_vertices || batch.ranges_buffer(3)[0].vertex_count != shared[0].emitted_vertices) {
            std::printf("FAIL weld: %u -> %u -> %u vertices\n", unshared[0].emitted_vertices, shared[0].emitted_vertices, again[0].emitted_vertices);
            return 1;
        }
    }
    std::printf("OK cpp mirror: 1323 vertices / 661 triangles, errors and overflow contract as the reference\n");
    return 0;
}
