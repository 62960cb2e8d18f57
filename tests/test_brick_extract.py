"""Legacy 8^3-brick marching cubes.

The reference has no golden vectors for voxel_surface_extract.wgsl (parity unpinned), so the oracle
restatement is pinned on answers that follow from the shader's text alone; the CUDA kernel must then equal
the oracle bit for bit (positions, normals, indices, descriptors, indirect draws), brick by brick."""
import numpy as np
import pytest

import helio_b200 as H
from oracle import oracle as O


def brick(fill=None):
    v = np.zeros((9, 9, 9), dtype=np.uint8)      # [z][y][x]
    if fill is not None:
        fill(v)
    return v


def run_oracle(v, origin=(0.0, 0.0, 0.0), size=1.0):
    return O.brick_extract(H.pack_brick(v), 0, origin, size)


# ---- known answers (CPU) -------------------------------------------------------------------------------------
def test_pods_and_word_packing():
    assert H.BRICK_META_DTYPE.itemsize == 8 and H.DIRTY_BRICK_DTYPE.itemsize == 32
    assert H.BRICK_MESHLET_DTYPE.itemsize == 32 and H.BRICK_DRAW_DTYPE.itemsize == 20
    v = brick()
    v[0, 0, 1], v[8, 8, 8] = 7, 9
    words = H.pack_brick(v)
    assert words.size == H.VOXEL_MESH_BRICK_VOXEL_WORDS == 183
    assert words[0] == 7 << 8 and words[182] == 9            # linear 1 -> byte 1 of word 0; linear 728 -> byte 0 of word 182


def test_empty_and_full_bricks_emit_nothing():
    for v in (brick(), brick(lambda b: b.fill(3))):
        verts, normals, idx, raw = run_oracle(v)
        assert raw == 0 and len(verts) == 0


def test_single_voxel_is_an_octahedron_of_eight_triangles():
    verts, normals, idx, raw = run_oracle(brick(lambda b: b.__setitem__((4, 4, 4), 5)), origin=(10.0, 20.0, 30.0), size=0.5)
    assert raw == 24 and np.array_equal(idx, np.arange(24))   # eight corner cells, one triangle each
    assert np.all(verts[:, 3] == 5.0)                          # material = first non-zero corner
    # every vertex is the midpoint of one of the six edges leaving the voxel: |p - centre|_1 == half a voxel
    centre = np.array([10.0, 20.0, 30.0]) + 4 * 0.5
    d = np.abs(verts[:, :3] - centre)
    assert np.allclose(d.sum(axis=1), 0.25) and np.allclose(d.max(axis=1), 0.25)
    assert len({tuple(p) for p in verts[:, :3]}) == 6
    assert np.allclose(np.linalg.norm(normals[:, :3], axis=1), 1.0, atol=1e-6) and np.all(normals[:, 3] == 0.0)


def test_half_filled_brick_is_a_flat_sheet():
    verts, normals, idx, raw = run_oracle(brick(lambda b: b.__setitem__((slice(None), slice(0, 4), slice(None)), 2)))
    assert raw == 8 * 8 * 6                                    # 64 surface cells, two triangles each
    assert np.all(verts[:, 1] == 3.5) and np.all(verts[:, 3] == 2.0)
    # occupancy decreases with y, and the normal is the (unnormalised) occupancy gradient: (0, -1, 0)
    assert np.all(normals[:, 0] == 0.0) and np.all(normals[:, 1] == -1.0) and np.all(normals[:, 2] == 0.0)
    # cells come in linear order, x fastest: the first cell's six entries sit in cell (0, 3, 0)
    assert np.all((verts[:6, 0] >= 0.0) & (verts[:6, 0] <= 1.0) & (verts[:6, 2] >= 0.0) & (verts[:6, 2] <= 1.0))


def test_every_cube_case_emits_its_table_row():
    edge_mid = np.array([[.5, 0, 0], [1, .5, 0], [.5, 1, 0], [0, .5, 0], [.5, 0, 1], [1, .5, 1], [.5, 1, 1], [0, .5, 1],
                         [0, 0, .5], [1, 0, .5], [1, 1, .5], [0, 1, .5]], dtype=np.float32)
    corners = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
    import re
    from pathlib import Path
    text = (Path(O.__file__).parent / "mc_tables.inc").read_text().split("HVXO_MC_EDGES", 1)[1].split("};", 1)[0]
    table = [int(w, 16) for w in re.findall(r"0x[0-9A-Fa-f]+", text.split("=", 1)[1])]
    assert len(table) == 256 * 16
    for case in range(1, 255):
        v = brick()
        for bit, (x, y, z) in enumerate(corners):              # isolate the case in cell (3, 3, 3)
            if case >> bit & 1:
                v[3 + z, 3 + y, 3 + x] = 10 + bit
        verts, normals, idx, raw = run_oracle(v)
        # find this cell's entries: vertices inside [3,4]^3 whose cell is exactly (3,3,3) come after the cells before it
        row = table[16 * case:16 * case + 16]
        row = row[:row.index(255)]
        want = np.array([edge_mid[e] + 3 for e in row], dtype=np.float32)
        # neighbouring cells see a subset of the corners; locate the run that matches the row
        pos = verts[:, :3]
        hits = [s for s in range(len(pos) - len(want) + 1) if np.array_equal(pos[s:s + len(want)], want)]
        assert hits, case
        first_material = next(10 + b for b in range(8) if case >> b & 1)
        assert any(np.all(verts[s:s + len(want), 3] == first_material) for s in hits), case


def test_overflow_drops_whole_cells_and_clamps_the_published_counts():
    def checker(b):
        z, y, x = np.indices((9, 9, 9))
        b[:] = ((x + y + z) & 1) * 4
    verts, normals, idx, raw = run_oracle(brick(checker))
    assert raw == 512 * 12 and len(verts) == 2048              # every cell is the 4-triangle checker case
    written = np.any(verts != 0, axis=1)
    assert written[:2040].all() and not written[2040:].any()   # 170 cells x 12 fit; cell 171 would end at 2052
    assert np.array_equal(idx[:2040], np.arange(2040))


def test_active_brick_range_follows_the_reference_tests():
    r = H.ActiveBrickRange(8)                                    # lib.rs:793-823
    assert r.draw_count() == 0
    assert r.set(2, True) and r.set(6, True) and r.draw_count() == 7
    assert r.set(6, False) and r.draw_count() == 3
    r = H.ActiveBrickRange(8)
    assert r.set(4, True) and r.set(4, False) and r.draw_count() == 0
    assert not r.set(8, True) and r.draw_count() == 0


# ---- GPU parity ------------------------------------------------------------------------------------------------
def _scene(rng, n):
    """n bricks: noise, blobs, planes, checkers, empty, full -- at shuffled data offsets."""
    bricks = []
    for k in range(n):
        kind = k % 6
        z, y, x = np.indices((9, 9, 9))
        if kind == 0:
            b = (rng.random((9, 9, 9)) < rng.uniform(0.05, 0.95)) * rng.integers(1, 255, (9, 9, 9))
        elif kind == 1:
            c = rng.uniform(1, 7, 3)
            b = (((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2) < rng.uniform(2, 30)) * int(rng.integers(1, 255))
        elif kind == 2:
            nrm = rng.normal(size=3)
            b = ((nrm[0] * x + nrm[1] * y + nrm[2] * z) < nrm.sum() * 4) * int(rng.integers(1, 255))
        elif kind == 3:
            b = ((x + y + z) & 1) * 4
        elif kind == 4:
            b = np.zeros((9, 9, 9))
        else:
            b = np.full((9, 9, 9), 200)
        bricks.append(b.astype(np.uint8))
    return bricks


@pytest.mark.gpu
def test_gpu_bricks_equal_the_oracle_bit_for_bit():
    rng = np.random.default_rng(11)
    n = 96
    bricks = _scene(rng, n)
    order = rng.permutation(n)                                  # slot k's data lives at a shuffled offset
    words = np.zeros(n * 183 + 5, dtype=np.uint32)
    meta = np.zeros(n, dtype=H.BRICK_META_DTYPE)
    for k in range(n):
        meta[k]["data_offset"] = 5 + int(order[k]) * 183
        words[meta[k]["data_offset"]:meta[k]["data_offset"] + 183] = H.pack_brick(bricks[k])
    ex = H.VoxelMeshExtractor(0, max_bricks=n)
    ex.write_brick_meta(meta)
    ex.write_voxel_data(words)
    origins = rng.uniform(-100, 100, (n, 3)).astype(np.float32)
    sizes = rng.choice([0.1, 0.25, 1.0, 0.37], n).astype(np.float32)
    for k in rng.permutation(n):
        assert ex.mark_dirty(int(k), 1000 + int(k), origins[k], sizes[k], bool(bricks[k].any()))
    assert not ex.mark_dirty(n, 0, (0, 0, 0), 1.0, True)        # out of range: dropped
    assert ex.execute() == ex.active_bricks.draw_count() and not ex.dirty_bricks
    desc, draws = ex.descriptors(), ex.indirect_draws()
    total = 0
    for k in range(n):
        wv, wn, wi, raw = O.brick_extract(words, int(meta[k]["data_offset"]), origins[k], sizes[k])
        count = min(raw, 2048)
        d = desc[k]
        assert (d["vertex_offset"], d["index_offset"], d["vertex_count"], d["index_count"], d["brick_index"], d["volume_id"]) == \
            (k * 2048, k * 2048, count, count, k, 1000 + k), k
        assert tuple(int(x) for x in draws[k]) == (count, 1 if count else 0, k * 2048, k * 2048, 0), k
        gv, gn, gi = ex.brick_mesh(k)
        written = np.any(wv != 0, axis=1) | np.any(wn != 0, axis=1)   # entries of dropped cells are never written
        assert gv[written].tobytes() == wv[written].tobytes(), k
        assert gn[written].tobytes() == wn[written].tobytes(), k
        assert np.array_equal(gi[written], wi[written]), k
        total += count
    assert total > 20000 and (desc["vertex_count"] == 2048).any() and (desc["vertex_count"] == 0).any()

    # re-extraction of one edited brick leaves every other slot alone; clearing a slot zeroes its draw
    bricks[7][:] = 0
    bricks[7][2:5, 2:5, 2:5] = 9
    words[meta[7]["data_offset"]:meta[7]["data_offset"] + 183] = H.pack_brick(bricks[7])
    ex.write_voxel_data(words)
    ex.mark_dirty(7, 42, origins[7], sizes[7], True)
    ex.execute()
    after = ex.descriptors()
    wv, wn, wi, raw = O.brick_extract(words, int(meta[7]["data_offset"]), origins[7], sizes[7])
    assert after[7]["vertex_count"] == raw and after[7]["volume_id"] == 42
    assert ex.brick_mesh(7)[0].tobytes() == wv.tobytes()
    keep = np.arange(n) != 7
    assert np.array_equal(after[keep], desc[keep])
    assert ex.clear_brick_slot(3) and tuple(int(x) for x in ex.indirect_draws(3, 1)[0]) == (0, 0, 0, 0, 0)
    assert not ex.clear_brick_slot(n)
    ex.close()


@pytest.mark.gpu
def test_gpu_brick_errors_are_loud():
    ex = H.VoxelMeshExtractor(0, max_bricks=4)
    ex.write_voxel_data(np.zeros(183, dtype=np.uint32))
    meta = np.zeros(4, dtype=H.BRICK_META_DTYPE)
    meta[1]["data_offset"] = 1                                  # words [1, 184) of 183
    ex.write_brick_meta(meta)
    dirty = np.zeros(1, dtype=H.DIRTY_BRICK_DTYPE)
    dirty[0]["brick_slot"] = 1
    with pytest.raises(H.SampleCount):
        ex.extract(dirty)
    dirty[0]["brick_slot"] = 4
    with pytest.raises(H.BatchCapacity):
        ex.extract(dirty)
    ex.close()
