"""The optional vertex-reuse output (hvx_weld_meshes): the oracle's definition, checked against an independent count.

The reference shares no vertices, so the definition rests on its own, pinned, output: merging the
bit-identical vertex records of the unshared mesh.  These CPU tests check that this IS edge sharing -- the kept
vertices of a page are exactly its crossing cell edges, counted straight from the sample signs."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import weld
from hvx_testutil import ALL, FIXTURE_PAGES


def crossing_edges(samples, edge=32):
    """Distinct cell edges of the page whose end points differ in solidity (corners 1 .. edge+1 of the haloed block)."""
    s = edge + 2
    d = (np.asarray(samples, dtype=np.uint32).reshape(s, s, s) & 0xFFFF).astype(np.uint16).view(np.int16)
    solid = d[1:, 1:, 1:] <= 0            # corners of the page's cells, (edge+1)^3, [z, y, x]
    return int((solid[:, :, 1:] != solid[:, :, :-1]).sum() + (solid[:, 1:, :] != solid[:, :-1, :]).sum() +
               (solid[1:, :, :] != solid[:-1, :, :]).sum())


@pytest.mark.parametrize("name", sorted(FIXTURE_PAGES))
def test_kept_vertices_are_the_crossing_edges(name):
    kind, page = FIXTURE_PAGES[name]
    samples = O.fixture_fill(kind, page)
    mesh = O.extract_regular(samples, edge=32, transition_mask=0, dirty_microbricks=ALL, generation=3, debug=False)
    kept, indices, kept_from = weld.weld_mesh(mesh.vertices, mesh.indices)
    edges = crossing_edges(samples)
    assert len(mesh.vertices) > len(kept) > 0
    # two different edges can only give one record when they meet in a corner of density exactly 0 (t = 0 there)
    d = (np.asarray(samples, dtype=np.uint32) & 0xFFFF).astype(np.uint16).view(np.int16)
    if not (d == 0).any():
        assert len(kept) == edges, (len(kept), edges)
    else:
        assert len(kept) <= edges
    # the triangles are the same triangles
    raw = np.ascontiguousarray(mesh.vertices).view(np.uint32).reshape(-1, 8)
    assert np.array_equal(raw[mesh.indices], np.ascontiguousarray(kept).view(np.uint32).reshape(-1, 8)[indices])
    # first occurrences, in the original order
    assert np.all(np.diff(kept_from) > 0) and kept_from[0] == 0
    assert len(np.unique(np.ascontiguousarray(kept).view(np.dtype((np.void, 32))))) == len(kept)


def test_plane_page_shares_down_to_one_vertex_per_corner():
    """The reference's published plane result is 4,096 vertices for 32 x 32 cells
    (docs/planetary_voxel_extraction_benchmark.md:59); shared, that is the 33 x 33 corners of the layer."""
    kind, page = FIXTURE_PAGES["plane"]
    mesh = O.extract_regular(O.fixture_fill(kind, page), edge=32, transition_mask=0, dirty_microbricks=ALL, generation=1, debug=False)
    kept, indices, _ = weld.weld_mesh(mesh.vertices, mesh.indices)
    assert (len(mesh.vertices), len(kept), len(indices)) == (4096, 1089, 6144)


def test_weld_is_idempotent_and_handles_the_empty_mesh():
    kind, page = FIXTURE_PAGES["sphere"]
    mesh = O.extract_regular(O.fixture_fill(kind, page), edge=32, transition_mask=0x3F, dirty_microbricks=ALL, generation=1, debug=False)
    kept, indices, _ = weld.weld_mesh(mesh.vertices, mesh.indices)
    again, indices2, kept_from = weld.weld_mesh(kept, indices)
    assert again.tobytes() == kept.tobytes() and np.array_equal(indices2, indices) and np.array_equal(kept_from, np.arange(len(kept)))
    empty_v, empty_i, _ = weld.weld_mesh(mesh.vertices[:0], mesh.indices[:0])
    assert len(empty_v) == 0 and len(empty_i) == 0
