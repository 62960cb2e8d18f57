"""Sphere edits on resident samples (BASELINE config 5): oracle properties on the CPU, GPU parity through the C ABI."""
import numpy as np
import pytest

import helio_b200 as H
from oracle import edit as OE
from oracle import oracle as O
from hvx_testutil import assert_vertices_equal

VOXEL_OP_ADD_SPHERE, VOXEL_OP_SUBTRACT_SPHERE = 1, 2   # crates/helio-voxel-core/src/edit.rs:5-10


def _cells_by_microbrick(edge, dirty):
    q = edge // 4
    z, y, x = np.indices((edge, edge, edge))
    bit = (x // q) + 4 * (y // q) + 16 * (z // q)
    return ((np.uint64(dirty) >> bit.astype(np.uint64)) & np.uint64(1)).astype(bool).ravel()


@pytest.mark.parametrize("op", [VOXEL_OP_SUBTRACT_SPHERE, VOXEL_OP_ADD_SPHERE])
@pytest.mark.parametrize("xz,radius", [((1.7, 1.1), 0.45), ((0.05, 3.15), 0.7), ((3.21, 0.02), 0.33)])
def test_dirty_microbricks_cover_every_cell_whose_mesh_changed(op, xz, radius):
    """The octree rule with r + 2 cells must mark every microbrick in which a cell's vertices or triangles change
    (a vertex reads the gradient neighbours of its cell's corners), on the edited chunk and on its neighbours."""
    # the page layer in which this column's terrain surface lies, and the surface height there
    for py in (-3, -2, -1, 0):
        column = O.fixture_fill(O.FIELD_TERRAIN_FBM, [0, py, 0]).reshape(34, 34, 34)[min(int(xz[1] * 10) + 1, 33), 1:33, min(int(xz[0] * 10) + 1, 33)]
        solid = (column & 0xFFFF).astype(np.uint16).view(np.int16) <= 0
        if solid.any() and not solid.all():
            break
    else:
        pytest.fail("no surface in this column")
    center = (xz[0], (py * 32 + int(np.nonzero(solid)[0].max())) * 0.1 + 0.03, xz[1])
    pages = np.array([[0, py, 0], [1, py, 0], [0, py, 1], [0, py + 1, 0], [0, py - 1, 0]], dtype=np.int64)
    before = np.concatenate([O.fixture_fill(O.FIELD_TERRAIN_FBM, [int(v) for v in p]) for p in pages])
    after, dirty, touched = OE.apply_edit(before, pages, None, 32, op, center, radius, material=3)
    assert len(touched) >= 1 and np.any(after != before)
    words = 34 ** 3
    for i in range(len(pages)):
        a = O.extract_regular(before[i * words:(i + 1) * words], generation=1)
        b = O.extract_regular(after[i * words:(i + 1) * words], generation=1)
        changed_cells = a.cell_words[:, 0] != b.cell_words[:, 0]
        # per-cell vertex bytes, for cells whose case word is unchanged
        for cell in np.nonzero(~changed_cells)[0]:
            nv = (int(a.cell_words[cell, 0]) >> 16) & 0xFF
            if nv and a.vertices[a.cell_ranges[cell, 0]:a.cell_ranges[cell, 0] + nv].tobytes() != \
                    b.vertices[b.cell_ranges[cell, 0]:b.cell_ranges[cell, 0] + nv].tobytes():
                changed_cells[cell] = True
        covered = _cells_by_microbrick(32, dirty[i])
        assert not np.any(changed_cells & ~covered), f"chunk {i}: {np.count_nonzero(changed_cells & ~covered)} changed cells outside the dirty set"
        if i not in touched:
            assert dirty[i] == 0 and not changed_cells.any()


@pytest.mark.gpu
@pytest.mark.parametrize("edge", [32, 64])
@pytest.mark.parametrize("op", [VOXEL_OP_SUBTRACT_SPHERE, VOXEL_OP_ADD_SPHERE])
def test_gpu_edit_matches_the_oracle_and_reextracts(edge, op):
    rng = np.random.default_rng(edge + op)
    pages = np.array([[x, y, z] for z in (0, 1) for y in (-2, -1) for x in (-1, 0, 1)], dtype=np.int64)
    n = len(pages)
    batch = H.ChunkBatchExtractor(0, edge=edge, max_chunks=n, max_vertices=60_000, max_indices=90_000)
    batch.fill_density(O.FIELD_TERRAIN_FBM, pages)
    words = (edge + 2) ** 3
    host = batch.ctx.read(H._ffi.BUF_SAMPLES, 0, n * words)
    span = edge * 0.1
    before, changed = host.copy(), 0
    for k in range(6):
        center = (float(rng.uniform(-0.6 * span, 1.6 * span)), float(rng.uniform(-5.0, -3.6)), float(rng.uniform(0.1 * span, 1.9 * span)))
        radius = float(rng.uniform(0.25, 1.5))
        dirty, touched = batch.ctx.apply_edit(op, center, radius, pages, material=7)
        host, want_dirty, want_touched = OE.apply_edit(host, pages, None, edge, op, center, radius, material=7)
        assert touched == len(want_touched) and np.array_equal(dirty, want_dirty), k
        got = batch.ctx.read(H._ffi.BUF_SAMPLES, 0, n * words)
        assert np.array_equal(got, host), f"edit {k}: {np.count_nonzero(got != host)} samples differ from the oracle"
        changed += int(np.count_nonzero(got != before))
        before = got
        # re-extract only what the edit made dirty; every chunk's partial mesh equals the oracle's for the same mask
        todo = np.nonzero(dirty)[0]
        if len(todo) == 0:
            continue
        descs = H.make_descs(n, 100 + k, [int(d) for d in dirty])
        batch.ctx.extract_regular(None, descs, n)
        counters = batch.counters(n)
        for i in todo[:3]:
            want = O.extract_regular(host[i * words:(i + 1) * words], edge=edge, generation=100 + k, dirty_microbricks=int(dirty[i]), debug=False)
            v, idx = batch.chunk_mesh(int(i))
            assert counters["required_vertices"][i] == len(want.vertices)
            assert_vertices_equal(v, want.vertices, f"edit {k} chunk {i}")
            assert np.array_equal(idx, want.indices)
    assert changed > 1000, "the edits of this test must actually carve / fill the terrain"
    with pytest.raises(H.HvxError, match="not a sphere edit"):
        batch.ctx.apply_edit(0, (0, 0, 0), 1.0, pages)
    batch.close()
