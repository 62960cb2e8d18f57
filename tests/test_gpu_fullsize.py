"""Bit-exact parity at BASELINE's full size (configs[1]: 4096 fBm chunks of 64^3) through hashes.

tests/golden/terrain_4096x64.json holds, for every chunk of the bench's 16x16x16 grid that has a
mesh, the oracle's vertex / index counts and CRC-32s (tools/gen_golden_terrain.py, CPU only).  The GPU
leg fills the same grid on the device, extracts it in ONE dispatch and compares every entry.
"""
import json
import zlib
from pathlib import Path

import numpy as np
import pytest

import helio_b200 as H
from oracle import oracle as O

GOLDEN = json.loads((Path(__file__).parent / "golden" / "terrain_4096x64.json").read_text())


def _pages():
    g = GOLDEN["grid"]
    xs = np.arange(-g // 2, g // 2, dtype=np.int64)
    z, y, x = np.meshgrid(xs, xs, xs, indexing="ij")
    return np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)


def test_golden_file_is_what_the_oracle_produces():
    """CPU: three chunks of the golden file recomputed by the oracle (the generator is committed)."""
    pages = _pages()
    keys = sorted(GOLDEN["chunks"], key=int)
    assert GOLDEN["chunks_with_mesh"] == len(keys) == 298
    assert GOLDEN["vertices_total"] == sum(GOLDEN["chunks"][k][0] for k in keys) == 6_217_668
    assert GOLDEN["indices_total"] == sum(GOLDEN["chunks"][k][1] for k in keys) == 9_326_502
    for k in (keys[0], keys[len(keys) // 2], keys[-1]):
        s = O.fixture_fill(O.FIELD_TERRAIN_FBM, [int(v) for v in pages[int(k)]], edge=GOLDEN["edge"])
        m = O.extract_regular(s, edge=GOLDEN["edge"], debug=False)
        assert [len(m.vertices), len(m.indices), zlib.crc32(m.vertices.tobytes()), zlib.crc32(m.indices.tobytes())] == \
            GOLDEN["chunks"][k]


@pytest.mark.gpu
def test_headline_batch_matches_the_oracle_chunk_by_chunk():
    pages = _pages()
    n = len(pages)
    batch = H.ChunkBatchExtractor(0, edge=GOLDEN["edge"], max_chunks=n, max_vertices=49_152, max_indices=73_728)
    batch.fill_density(O.FIELD_TERRAIN_FBM, pages)
    batch.extract_regular(None, n)
    counters = batch.counters(n)
    assert int(counters["completed"].sum()) == n
    assert int(counters["vertex_overflow"].sum()) == 0 and int(counters["index_overflow"].sum()) == 0
    assert int(counters["emitted_vertices"].astype(np.int64).sum()) == GOLDEN["vertices_total"]
    assert int(counters["emitted_indices"].astype(np.int64).sum()) == GOLDEN["indices_total"]
    verts, idx, packed = batch.ctx.read_meshes(0, 0, n)
    with_mesh = 0
    for i in range(n):
        r = packed[i]
        want = GOLDEN["chunks"].get(str(i))
        if want is None:
            assert r["vertex_count"] == 0 and r["index_count"] == 0, i
            continue
        with_mesh += 1
        v = verts[r["first_vertex"]:r["first_vertex"] + r["vertex_count"]]
        t = idx[r["first_index"]:r["first_index"] + r["index_count"]]
        assert [len(v), len(t), zlib.crc32(v.tobytes()), zlib.crc32(t.tobytes())] == want, f"chunk {i} page {pages[i]}"
    assert with_mesh == GOLDEN["chunks_with_mesh"]
    batch.close()
