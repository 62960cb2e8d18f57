"""GPU parity of hvx_weld_meshes (optional vertex-reuse output) against oracle/weld.py applied to the oracle's mesh."""
import numpy as np
import pytest

import helio_b200 as H
from oracle import oracle as O
from oracle import weld
from hvx_testutil import ALL, FIXTURE_PAGES, assert_vertices_equal

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("edge", [32, 64])
def test_welded_batch_matches_the_oracle(edge):
    """Every fixture plus terrain pages, with transition masks and a partially dirty chunk, an empty chunk and one that
    overflows: vertices, indices, ranges and counters after the weld; a second weld changes nothing."""
    rng = np.random.default_rng(edge)
    specs = [FIXTURE_PAGES[name] for name in sorted(FIXTURE_PAGES)] + [(O.FIELD_TERRAIN_FBM, [x, -1, z]) for x in range(3) for z in range(2)]
    specs.append((O.FIELD_PLANE, [0, 5, 0]))   # all air: an empty mesh
    specs.append((O.FIELD_DENSE_RANDOM, [0, 0, 0]))   # ~6 vertices per cell: overflows the slot, publishes nothing
    n = len(specs)
    masks = [int(m) for m in rng.integers(0, 64, n)]
    dirty = [ALL] * n
    dirty[7] = 0x0000FFFF0000FFFF
    mv, mi = (60_000, 90_000) if edge == 64 else (14_000, 21_000)
    batch = H.ChunkBatchExtractor(0, edge=edge, max_chunks=n, max_vertices=mv, max_indices=mi)
    samples = np.concatenate([O.fixture_fill(k, p, edge=edge) for k, p in specs])
    descs = H.make_descs(n, 9, dirty, masks)
    batch.ctx.extract_regular(samples, descs, n)
    before = batch.counters(n).copy()
    batch.ctx.weld_meshes(n)
    v, i, r = batch.ctx.read_meshes(0, 0, n)
    v, i, r = v.copy(), i.copy(), r.copy()
    c = batch.counters(n).copy()
    words = (edge + 2) ** 3
    total_before = total_after = 0
    for k in range(n):
        want = O.extract_regular(samples[k * words:(k + 1) * words], edge=edge, transition_mask=masks[k], dirty_microbricks=dirty[k],
                                 generation=9, debug=False)
        if len(want.vertices) > mv or len(want.indices) > mi:
            assert c["emitted_vertices"][k] == 0 and r["vertex_count"][k] == 0, k
            continue
        kept, idx, _ = weld.weld_mesh(want.vertices, want.indices)
        assert before["emitted_vertices"][k] == len(want.vertices), k
        assert c["required_vertices"][k] == len(want.vertices) and c["emitted_vertices"][k] == len(kept), k
        assert c["emitted_indices"][k] == len(idx) and r["vertex_count"][k] == len(kept) and r["index_count"][k] == len(idx), k
        assert_vertices_equal(v[r["first_vertex"][k]:r["first_vertex"][k] + r["vertex_count"][k]], kept, f"chunk {k}")
        assert np.array_equal(i[r["first_index"][k]:r["first_index"][k] + r["index_count"][k]], idx), k
        total_before += len(want.vertices)
        total_after += len(kept)
    assert total_after * 2 < total_before, "sharing removes well over half of the vertices"
    batch.ctx.weld_meshes(n)
    v2, i2, r2 = batch.ctx.read_meshes(0, 0, n)
    assert v2.tobytes() == v.tobytes() and i2.tobytes() == i.tobytes() and r2.tobytes() == r.tobytes()
    # the next extraction starts from the unshared mesh again
    batch.ctx.extract_regular(samples, descs, n)
    assert batch.counters(n).tobytes() == before.tobytes()
    batch.close()


def test_welded_transition_meshes_match_the_oracle():
    specs = [(O.FIELD_SPHERE, [0, 0, 0], 0x15), (O.FIELD_PLANE, [0, -1, 0], 0x3F), (O.FIELD_SPHERE, [-1, -1, -1], 0x2A),
             (O.FIELD_PLANE, [0, 3, 0], 0x3F), (O.FIELD_CAVE, [0, -1, 0], 0x00)] * 31
    n = len(specs)
    batch = H.ChunkBatchExtractor(0, edge=32, max_chunks=n, max_vertices=8, max_indices=8,
                                  max_transition_vertices=4096, max_transition_indices=8192)
    slabs = np.concatenate([O.slab_fill(k, p, 1) for k, p, _ in specs[:5]] * 31)
    batch.extract_transition(slabs, n, [m for _, _, m in specs])
    batch.ctx.weld_meshes(n, kind=1)
    counters = batch.transition_counters(n)
    wants = [O.extract_transition(slabs[i * 80802:(i + 1) * 80802], specs[i][2], debug=False) for i in range(5)]
    welded = [weld.weld_mesh(w.vertices, w.indices) for w in wants]
    shared = 0
    for i in range(n):
        kept, idx, _ = welded[i % 5]
        assert counters["required_vertices"][i] == len(wants[i % 5].vertices) and counters["emitted_vertices"][i] == len(kept), i
        if i in (0, 1, 2, 3, 4, 77, n - 1):
            v, got = batch.chunk_mesh(i, kind=1)
            assert_vertices_equal(v, kept, f"chunk {i}")
            assert np.array_equal(got, idx), i
        shared += len(wants[i % 5].vertices) - len(kept)
    assert shared > 0
    batch.close()


def test_weld_errors():
    batch = H.ChunkBatchExtractor(0, edge=32, max_chunks=4, max_vertices=4096, max_indices=6144)
    batch.fill_density(O.FIELD_PLANE, np.array([[0, -1, 0], [1, -1, 0]], dtype=np.int64))
    batch.ctx.extract_regular(None, H.make_descs(2), 2)
    with pytest.raises(H.HvxError, match="weld requested for 3 chunks"):
        batch.ctx.weld_meshes(3)
    with pytest.raises(H.HvxError, match="kind must be"):
        batch.ctx.weld_meshes(2, kind=2)
    with pytest.raises(H.HvxError, match="transition capacities"):
        batch.ctx.weld_meshes(1, kind=1)
    batch.ctx.weld_meshes(2)
    assert list(batch.counters(2)["emitted_vertices"]) == [1089, 1089]
    batch.close()


def test_meshlets_of_a_welded_mesh_match_the_oracle():
    """The weld composes with the next step of the reference's pass: build_terrain_meshlets on the shared-vertex mesh
    (its unique-vertex counts now see the sharing) equals the oracle's builder run on the oracle-welded mesh."""
    edge, specs = 32, [(O.FIELD_SPHERE, [0, 0, 0], 0), (O.FIELD_PLANE, [0, -1, 0], 0x3F), (O.FIELD_TERRAIN_FBM, [0, -1, 0], 0)]
    n, maxv, maxi = len(specs), 16_384, 24_576
    batch = H.ChunkBatchExtractor(0, edge=edge, max_chunks=n, max_vertices=maxv, max_indices=maxi)
    samples = np.concatenate([O.fixture_fill(k, p, edge=edge) for k, p, _ in specs])
    gens = [77 + i for i in range(n)]
    batch.extract_regular(samples, n, generation=gens, transition_mask=[m for _, _, m in specs])
    batch.ctx.weld_meshes(n)
    batch.ctx.build_meshlets(n, 0)
    stride = (maxi + 62) // 63
    for i, (kind, page, mask) in enumerate(specs):
        mesh = O.extract_regular(samples[i * 34 ** 3:(i + 1) * 34 ** 3], edge=edge, transition_mask=mask, debug=False)
        kept, idx, _ = weld.weld_mesh(mesh.vertices, mesh.indices)
        want_m, want_b = O.build_meshlets(kept, idx, i * maxi, i * maxv, i * stride, gens[i], 0)
        got_m, got_b = batch.ctx.read_meshlets(i, 0)
        assert got_m.tobytes() == want_m.tobytes(), f"chunk {i}: descriptors"
        for name in got_b.dtype.names:
            assert np.array_equal(got_b[name], want_b[name]), f"chunk {i}: bounds.{name}"
    batch.close()
