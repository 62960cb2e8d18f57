"""Cross-LOD seam correctness (BASELINE configs[2]), on the oracle (CPU) and on the GPU output.

1. A three-level horizon plan (PV/src/lod_topology.rs:169-217): every coarse page's transition mesh must put its
   full-resolution side exactly where the finer neighbours' regular meshes put their boundary vertices -- that is
   what makes the seam watertight (PV/src/transvoxel_transition.rs:189-270; positions quantised like :904-925).
2. The reference's randomised all-face neighbourhood test (PV/src/transvoxel_transition.rs:701-776): neighbouring
   transition cells publish identical lateral seam vertices, no triangle is degenerate or duplicated, every vertex
   carries its face bit.  8 seeds x 6 faces, the whole 32 x 32 face instead of the reference's 8 x 8.
"""
import numpy as np
import pytest

import helio_b200 as H
from oracle import oracle as O
import seam_checks as S


@pytest.mark.parametrize("kind", [O.FIELD_PLANE, O.FIELD_TERRAIN_FBM])
def test_oracle_transition_meets_the_finer_neighbours_boundary(kind):
    keys, pages, lods, masks = S.plan_pages(kind)
    regular, transition = {}, {}
    for g, (key, mask) in enumerate(keys):
        page = [int(v) for v in pages[g]]
        regular[g] = O.extract_regular(O.fixture_fill(kind, page, lod=int(lods[g])), transition_mask=mask, debug=False).vertices
        if mask:
            m = O.extract_transition(O.slab_fill(kind, page, int(lods[g])), mask, debug=False)
            transition[g] = (m.vertices, m.indices)
    faces, vertices = S.check_plan_seams(keys, pages, lods, masks, regular, transition)
    assert faces >= 4 and vertices > 100


@pytest.mark.parametrize("seed", [0, 5])
def test_oracle_randomised_neighbourhoods(seed):
    slabs = S.random_slabs(seed)
    want = O.extract_transition(slabs, 0x3F, generation=3)
    S.check_random_faces(want.vertices, want.indices, want.cell_words[:, 0], want.cell_ranges[:, 0], want.cell_ranges[:, 1], seed)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", [O.FIELD_PLANE, O.FIELD_TERRAIN_FBM])
def test_gpu_transition_meets_the_finer_neighbours_boundary(kind):
    keys, pages, lods, masks = S.plan_pages(kind)
    n = len(keys)
    batch = H.ChunkBatchExtractor(0, edge=S.EDGE, max_chunks=n, max_vertices=24_576, max_indices=36_864,
                                  max_transition_vertices=8192, max_transition_indices=24_576)
    batch.fill_density(kind, pages, lods)
    batch.extract_regular(None, n, transition_mask=masks)
    coarse = [g for g in range(n) if masks[g]]
    batch.fill_slabs(kind, pages[coarse], lods[coarse])
    batch.extract_transition(None, len(coarse), [masks[g] for g in coarse])
    assert int(batch.counters(n)["vertex_overflow"].sum()) == 0
    assert int(batch.transition_counters(len(coarse))["vertex_overflow"].sum()) == 0
    regular = {g: batch.chunk_mesh(g, kind=0)[0] for g in range(n)}
    transition = {g: batch.chunk_mesh(j, kind=1) for j, g in enumerate(coarse)}
    S.check_plan_seams(keys, pages, lods, masks, regular, transition)
    batch.close()


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(8))
def test_gpu_randomised_all_face_neighbourhoods_have_exact_seams_and_no_duplicate_triangles(seed):
    ex = H.TransvoxelGpuTransitionExtractor(0)
    slabs = S.random_slabs(seed)
    ex.dispatch(slabs, 0x3F, 77 + seed)
    c = ex.counters_buffer()
    assert c["completed"] == 1 and c["vertex_overflow"] == 0 and c["index_overflow"] == 0 and c["active_faces"] == 6
    verts = ex.vertices_buffer(int(c["emitted_vertices"]))
    idx = ex.indices_buffer(int(c["emitted_indices"]))
    cells, offsets, blocks = ex.cells_buffer(), ex.offsets_buffer(), ex.blocks_buffer()
    block_of = np.arange(6 * S.EDGE * S.EDGE) // 256
    S.check_random_faces(verts, idx, cells["packed_case_class_counts"], blocks["first_vertex"][block_of] + offsets["first_vertex"],
                         blocks["first_index"][block_of] + offsets["first_index"], seed)
    ex.close()
