"""GPU parity: transition (LOD seam) cells vs the CPU oracle.

Mirrors PV/tests/gpu_transvoxel_transitions.rs:16-313 with bit-exact float comparison.
"""
import numpy as np
import pytest

import helio_b200 as H
from oracle import oracle as O
from hvx_testutil import assert_vertices_equal

pytestmark = pytest.mark.gpu

CASES = [(O.FIELD_PLANE, [0, -1, 0], 0x3F), (O.FIELD_SPHERE, [-1, -1, -1], 0x2A), (O.FIELD_SPHERE, [0, 0, 0], 0x15)]
WEIGHTS = [0x001, 0x002, 0x004, 0x080, 0x100, 0x008, 0x040, 0x020, 0x010]


@pytest.fixture(scope="module")
def extractor():
    ex = H.TransvoxelGpuTransitionExtractor(0)
    yield ex
    ex.close()


def _check(ex, slabs, mask, generation, label, edge=32):
    want = O.extract_transition(slabs, mask, edge=edge, generation=generation)
    ex.dispatch(slabs, mask, generation)
    c = ex.counters_buffer()
    assert c["completed"] == 1 and c["vertex_overflow"] == 0 and c["index_overflow"] == 0, label
    assert c["active_faces"] == bin(mask).count("1"), label
    assert c["active_cells"] == want.counters[0], label
    assert c["required_vertices"] == len(want.vertices) and c["required_indices"] == len(want.indices), label
    assert c["emitted_vertices"] == c["required_vertices"] and c["emitted_indices"] == c["required_indices"]
    got_v = ex.vertices_buffer(len(want.vertices))
    got_i = ex.indices_buffer(len(want.indices))
    assert_vertices_equal(got_v, want.vertices, label)
    assert np.array_equal(got_i, want.indices), f"{label}: indices"
    # per-cell records, offsets and blocks (gpu_transvoxel_transitions.rs:111-179)
    cells, offsets, blocks = ex.cells_buffer(), ex.offsets_buffer(), ex.blocks_buffer()
    per_face = edge * edge
    gen = cells["generation_low"].astype(np.uint64) | (cells["generation_high"].astype(np.uint64) << np.uint64(32))
    valid = ((cells["packed_case_class_counts"] & 0x80000000) != 0) & (gen == generation)
    for face in range(6):
        sl = slice(face * per_face, (face + 1) * per_face)
        if not (mask >> face) & 1:
            assert not valid[sl].any(), f"{label}: inactive face {face}"
            continue
        assert valid[sl].all(), f"{label}: face {face}"
        assert np.array_equal(cells["packed_case_class_counts"][sl], want.cell_words[sl, 0]), f"{label}: face {face} cells"
        block_of = np.arange(face * per_face, (face + 1) * per_face) // 256
        assert np.array_equal(blocks["first_vertex"][block_of] + offsets["first_vertex"][sl], want.cell_ranges[sl, 0])
        assert np.array_equal(blocks["first_index"][block_of] + offsets["first_index"][sl], want.cell_ranges[sl, 1])
    return want, got_v, got_i


@pytest.mark.parametrize("case", range(3))
def test_transition_emission_matches_cpu_on_all_six_faces(extractor, case):
    kind, page, mask = CASES[case]
    slabs = O.slab_fill(kind, page, 1)
    _, got_v, got_i = _check(extractor, slabs, mask, 10_000 + case, f"case {case}")
    extractor.dispatch(slabs, mask, 10_000 + case)
    assert extractor.vertices_buffer(len(got_v)).tobytes() == got_v.tobytes()
    assert np.array_equal(extractor.indices_buffer(len(got_i)), got_i)


def _exhaustive_case_slabs():
    """gpu_transvoxel_transitions.rs:315-342: all 512 cases laid out on faces 0 and 1."""
    air, solid = O.cellword(1, 0, 0), O.cellword(-1, 1, 0)
    w = 67
    slabs = np.full(6 * 3 * w * w, air, dtype=np.uint32)
    for case in range(512):
        face, slot = case // 256, case % 256
        cu, cv = (slot % 16) * 2, (slot // 16) * 2
        for s in range(9):
            su, sv = cu * 2 + s % 3 + 1, cv * 2 + s // 3 + 1
            slabs[face * 3 * w * w + su + sv * w + w * w] = solid if case & WEIGHTS[s] else air
    return slabs


def test_exhaustive_512_case_sweep(extractor):
    slabs = _exhaustive_case_slabs()
    want, _, _ = _check(extractor, slabs, 0x03, 15_000, "sweep")
    cells = extractor.cells_buffer()
    for case in range(512):
        face, slot = case // 256, case % 256
        linear = face * 1024 + (slot % 16) * 2 + (slot // 16) * 2 * 32
        cell = H.GpuTransvoxelTransitionCell(int(cells["packed_case_class_counts"][linear]), 15_000)
        topo = O.case_topology(1, case)
        assert cell.case_index() == case
        assert cell.class_index() == topo["class_index"] and cell.reverse_winding() == topo["reverse"]
        assert cell.vertex_count() == topo["vertex_count"] and cell.triangle_count() == topo["triangle_count"]
    # mask 0: nothing active, every record invalidated (gpu_transvoxel_transitions.rs:259-283)
    extractor.dispatch(slabs, 0, 15_000)
    c = extractor.counters_buffer()
    assert c["completed"] == 1
    for key in ("active_faces", "active_cells", "required_vertices", "required_indices", "emitted_vertices", "emitted_indices"):
        assert c[key] == 0, key
    assert not ((extractor.cells_buffer()["packed_case_class_counts"] & 0x80000000) != 0).any()


def test_overflow_and_errors(extractor):
    slabs = O.slab_fill(O.FIELD_PLANE, [0, -1, 0], 1)
    tiny = H.TransvoxelGpuTransitionExtractor(0, H.TransvoxelGpuTransitionExtractorConfig.new(1, 1))
    tiny.dispatch(slabs, 0x3F, 20_000)
    c = tiny.counters_buffer()
    assert c["completed"] == 1 and c["vertex_overflow"] != 0 and c["index_overflow"] != 0
    assert c["emitted_vertices"] == 0 and c["emitted_indices"] == 0
    assert c["required_vertices"] == 640 and c["required_indices"] == 1152
    tiny.close()
    with pytest.raises(H.TransitionSampleCount):
        extractor.dispatch(slabs[:3], 1, 1)
    with pytest.raises(H.TransitionMask) as info:
        extractor.dispatch(slabs, 0x80, 1)
    assert info.value.mask == 0x80
    with pytest.raises(H.TransitionInvalidExtractionCapacity):
        H.TransvoxelGpuTransitionExtractorConfig.new(0, 1)


@pytest.mark.parametrize("kind,page,lod,mask", [
    (O.FIELD_SPHERE, [0, 0, 0], 1, 0x3F), (O.FIELD_SPHERE, [-1, -1, -1], 1, 0x3F), (O.FIELD_PLANE, [0, -1, 0], 2, 0x33),
    (O.FIELD_TERRAIN_FBM, [0, -1, 0], 1, 0x3F), (O.FIELD_DENSE_RANDOM, [0, 0, 0], 1, 0x3F),
])
@pytest.mark.parametrize("edge", [32, 64])
def test_edges_fields_and_lods(edge, kind, page, lod, mask):
    ex = H.TransvoxelGpuTransitionExtractor(0, edge=edge)
    _check(ex, O.slab_fill(kind, page, lod, edge=edge), mask, 31, f"e{edge} kind {kind} lod {lod}", edge=edge)
    ex.close()


def test_transition_batch():
    """Several coarse pages per dispatch, each with its own face mask."""
    specs = [(O.FIELD_SPHERE, [0, 0, 0], 0x15), (O.FIELD_PLANE, [0, -1, 0], 0x3F), (O.FIELD_SPHERE, [-1, -1, -1], 0x2A),
             (O.FIELD_PLANE, [0, 3, 0], 0x3F), (O.FIELD_CAVE, [0, -1, 0], 0x00)] * 70
    n = len(specs)
    batch = H.ChunkBatchExtractor(0, edge=32, max_chunks=n, max_vertices=8, max_indices=8,
                                  max_transition_vertices=4096, max_transition_indices=8192)
    slabs = np.concatenate([O.slab_fill(k, p, 1) for k, p, _ in specs[:5]] * 70)
    batch.extract_transition(slabs, n, [m for _, _, m in specs])
    counters = batch.transition_counters(n)
    wants = [O.extract_transition(slabs[i * 80802:(i + 1) * 80802], specs[i][2], debug=False) for i in range(5)]
    for i in range(n):
        assert counters["required_vertices"][i] == len(wants[i % 5].vertices), i
        assert counters["active_faces"][i] == bin(specs[i][2]).count("1")
    for i in (0, 1, 2, 173, n - 1):
        v, idx = batch.chunk_mesh(i, kind=1)
        assert_vertices_equal(v, wants[i % 5].vertices, f"chunk {i}")
        assert np.array_equal(idx, wants[i % 5].indices)
    batch.close()
