"""Meshlet build.  CPU: the oracle's restatement of build_terrain_meshlets against the
reference's own unit tests (PV/src/terrain_meshlet.rs:325-386).  GPU: the CUDA builder vs the oracle
on real extraction output -- descriptors exact, bounds bit-exact (the reference holds its WGSL to
1e-5 / 1e-4 only, PV/tests/gpu_terrain_meshlet_build.rs:183-194)."""
import numpy as np
import pytest

import helio_b200 as H
from oracle import oracle as O
from hvx_testutil import ALL, assert_vertices_equal  # noqa: F401


def _vertex_array(positions):
    v = np.zeros(len(positions), dtype=O.VERTEX_DTYPE)
    v["position"] = positions
    v["normal"] = [1.0, 0.0, 0.0]
    return v


def test_fixed_partition_preserves_every_triangle_and_generation():
    positions, indices = [], []
    for t in range(50):
        positions += [[t, 0, 0], [t, 1, 0], [t, 0, 1]]
        indices += [3 * t, 3 * t + 1, 3 * t + 2]
    generation = 0x0123_4567_89AB_CDEF
    meshlets, bounds = O.build_meshlets(_vertex_array(positions), indices, 100, 200, 300, generation, 7)
    assert len(meshlets) == 3 and meshlets["index_count"].sum() == 150
    for m, d in enumerate(meshlets):
        assert d["first_index"] == 100 + 63 * m and d["first_vertex"] == 200 and d["bounds_offset"] == 300 + m
        assert d["vertex_count"] <= 64 and d["index_count"] // 3 <= 96 and d["_pad"] == 7
        assert d["generation_low"] == generation & 0xFFFFFFFF and d["generation_high"] == generation >> 32


def test_no_welding_sphere_containment_and_cone():
    meshlets, _ = O.build_meshlets(_vertex_array([[0, 0, 0], [0, 0, 0], [0, 1, 0]]), [0, 1, 2])
    assert meshlets["vertex_count"][0] == 3                      # position-equal seam vertices stay distinct
    pts = np.array([[-4.0, 2.0, 0.5], [8.0, -3.0, 1.5], [1.0, 7.0, -6.0]], dtype=np.float32)
    _, bounds = O.build_meshlets(_vertex_array(pts), [0, 1, 2])
    assert np.all(np.linalg.norm(pts - bounds["center"][0], axis=1) <= bounds["radius"][0] + 1e-5)
    _, bounds = O.build_meshlets(_vertex_array([[0, 0, 0], [0, 1, 0], [0, 0, 1]]), [0, 1, 2])
    b = bounds[0]

    def cone_reject(camera):                                      # perspective_cone_reject, :157-169
        view = b["cone_apex"] - np.array(camera, dtype=np.float32)
        view = view / np.linalg.norm(view)
        return float(np.dot(view, b["cone_axis"])) >= b["cone_cutoff"]
    assert cone_reject([-10.0, 0.25, 0.25]) and not cone_reject([10.0, 0.25, 0.25])
    assert tuple(b["cone_axis"]) == (1.0, 0.0, 0.0) and np.isclose(b["cone_cutoff"], 1e-4)


def test_build_errors():
    v = _vertex_array([[0, 0, 0], [0, 1, 0], [0, 0, 1]])
    for indices, kind in [([0, 1], "IncompleteTriangle"), ([0, 1, 3], "IndexOutOfBounds")]:
        with pytest.raises(ValueError, match=kind):
            O.build_meshlets(v, indices)
    v["position"][1, 1] = np.inf
    with pytest.raises(ValueError, match="NonFinitePosition"):
        O.build_meshlets(v, [0, 1, 2])


def _check_chunk(ctx, chunk, mesh, generation, kind, max_vertices, max_indices):
    got_m, got_b = ctx.read_meshlets(chunk, kind)
    stride = (max_indices + 62) // 63
    want_m, want_b = O.build_meshlets(mesh.vertices, mesh.indices, chunk * max_indices, chunk * max_vertices,
                                      chunk * stride, generation, kind)
    assert len(got_m) == len(want_m) == (len(mesh.indices) + 62) // 63
    assert got_m.tobytes() == want_m.tobytes(), f"chunk {chunk}: descriptors"
    if got_b.tobytes() != want_b.tobytes():
        for name in got_b.dtype.names:
            assert np.array_equal(got_b[name], want_b[name]), f"chunk {chunk}: bounds.{name}"


@pytest.mark.gpu
@pytest.mark.parametrize("edge", [32, 64])
def test_gpu_meshlets_match_the_oracle(edge):
    specs = [(O.FIELD_SPHERE, [0, 0, 0], 0), (O.FIELD_PLANE, [0, -1, 0], 0x3F), (O.FIELD_TERRAIN_FBM, [0, -1, 0], 0),
             (O.FIELD_CAVE, [-1, -1, -1], 0x15), (O.FIELD_PLANE, [0, 2, 0], 0), (O.FIELD_DENSE_RANDOM, [1, 1, 1], 0)]
    n, maxv, maxi = len(specs), edge ** 3 * 12, edge ** 3 * 15
    batch = H.ChunkBatchExtractor(0, edge=edge, max_chunks=n, max_vertices=maxv, max_indices=maxi)
    samples = np.concatenate([O.fixture_fill(k, p, edge=edge) for k, p, _ in specs])
    gens = [0x0123_4567_0000_0000 + i for i in range(n)]
    batch.extract_regular(samples, n, generation=gens, transition_mask=[m for _, _, m in specs])
    batch.ctx.build_meshlets(n, 0)
    for i, (kind, page, mask) in enumerate(specs):
        mesh = O.extract_regular(samples[i * (edge + 2) ** 3:(i + 1) * (edge + 2) ** 3], edge=edge, transition_mask=mask, debug=False)
        _check_chunk(batch.ctx, i, mesh, gens[i], 0, maxv, maxi)
    assert len(batch.ctx.read_meshlets(4, 0)[0]) == 0            # empty chunk: no meshlets
    batch.close()
    # an overflowed chunk publishes no meshlets (terrain_meshlet_build.wgsl:207-210)
    tiny = H.ChunkBatchExtractor(0, edge=edge, max_chunks=1, max_vertices=8, max_indices=8)
    tiny.extract_regular(samples[:(edge + 2) ** 3], 1)
    tiny.ctx.build_meshlets(1, 0)
    assert len(tiny.ctx.read_meshlets(0, 0)[0]) == 0
    tiny.close()


@pytest.mark.gpu
def test_gpu_transition_meshlets_match_the_oracle():
    cases = [(O.FIELD_PLANE, [0, -1, 0], 0x3F), (O.FIELD_SPHERE, [-1, -1, -1], 0x2A), (O.FIELD_SPHERE, [0, 0, 0], 0x15)]
    n = len(cases)
    batch = H.ChunkBatchExtractor(0, edge=32, max_chunks=n, max_vertices=8, max_indices=8,
                                  max_transition_vertices=73_728, max_transition_indices=221_184)
    slabs = np.concatenate([O.slab_fill(k, p, 1) for k, p, _ in cases])
    batch.extract_transition(slabs, n, [m for _, _, m in cases], generation=77)
    batch.ctx.build_meshlets(n, 1)
    for i, (kind, page, mask) in enumerate(cases):
        mesh = O.extract_transition(slabs[i * 80802:(i + 1) * 80802], mask, debug=False)
        _check_chunk(batch.ctx, i, mesh, 77, 1, 73_728, 221_184)
    batch.close()
