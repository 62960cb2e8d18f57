"""GPU parity: K1 density / SDF fill (device twin of ExtractionFixture::new) vs the CPU oracle."""
import numpy as np
import pytest

import helio_b200 as H
from helio_b200 import _ffi
from oracle import oracle as O

pytestmark = pytest.mark.gpu

KINDS = [0, 1, 2, 3, 4, 5, 16, 17]
PAGES = [([0, 0, 0], 0), ([0, -1, 0], 0), ([-1, -1, -1], 0), ([7, -3, 2], 1), ([-2, 0, 5], 3),
         ([1_990_937, -1, -1], 0), ([-3, -1, 2], 6)]


@pytest.mark.parametrize("edge", [32, 64])
def test_sample_fill_is_bit_exact(edge):
    n = len(PAGES)
    batch = H.ChunkBatchExtractor(0, edge=edge, max_chunks=n, max_vertices=8, max_indices=8)
    words = (edge + 2) ** 3
    for kind in KINDS:
        batch.fill_density(kind, [p for p, _ in PAGES], [l for _, l in PAGES])
        got = batch.ctx.read(_ffi.BUF_SAMPLES, 0, n * words)
        for i, (page, lod) in enumerate(PAGES):
            want = O.fixture_fill(kind, page, lod=lod, edge=edge)
            diff = np.flatnonzero(got[i * words:(i + 1) * words] != want)
            assert diff.size == 0, f"kind {kind} page {page} lod {lod}: {diff.size} words differ, first {diff[:4]}"
    batch.close()


@pytest.mark.parametrize("edge", [32, 64])
def test_slab_fill_is_bit_exact(edge):
    pages = [([0, 0, 0], 1), ([0, -1, 0], 1), ([-1, -1, -1], 1), ([3, -1, -2], 2), ([-5, 0, 1], 4)]
    n = len(pages)
    batch = H.ChunkBatchExtractor(0, edge=edge, max_chunks=n, max_vertices=8, max_indices=8,
                                  max_transition_vertices=64, max_transition_indices=64)
    words = 18 * (2 * edge + 3) ** 2
    for kind in KINDS:
        batch.fill_slabs(kind, [p for p, _ in pages], [l for _, l in pages])
        got = batch.ctx.read(_ffi.BUF_SLABS, 0, n * words)
        for i, (page, lod) in enumerate(pages):
            want = O.slab_fill(kind, page, lod, edge=edge)
            diff = np.flatnonzero(got[i * words:(i + 1) * words] != want)
            assert diff.size == 0, f"kind {kind} page {page} lod {lod}: {diff.size} words differ"
    with pytest.raises(H.FinestLodHasNoFinerNeighbor):
        batch.fill_slabs(0, [[0, 0, 0]], [0])
    batch.close()


def test_fill_then_extract_without_leaving_the_device():
    """fill -> extract on the ctx sample arena (samples = NULL), the bench's device-resident path."""
    edge, pages = 64, [[x, -1, z] for x in range(-2, 2) for z in range(-2, 2)]
    batch = H.ChunkBatchExtractor(0, edge=edge, max_chunks=len(pages))
    batch.fill_density(16, pages)
    batch.extract_regular(None, len(pages))
    counters = batch.counters(len(pages))
    for i, page in enumerate(pages):
        want = O.extract_regular(O.fixture_fill(16, page, edge=edge), edge=edge, debug=False)
        assert counters["required_vertices"][i] == len(want.vertices) and counters["emitted_indices"][i] == len(want.indices)
        v, idx = batch.chunk_mesh(i)
        assert v.tobytes() == want.vertices.tobytes() and np.array_equal(idx, want.indices), page
    assert counters["required_vertices"].min() > 0
    batch.close()


def test_address_errors():
    batch = H.ChunkBatchExtractor(0, edge=32, max_chunks=2, max_vertices=8, max_indices=8)
    with pytest.raises(H.AddressError):
        batch.fill_density(0, [[2 ** 62, 0, 0]])
    with pytest.raises(H.AddressError):
        batch.fill_density(0, [[0, 0, 0]], [58])
    with pytest.raises(H.BatchCapacity):
        batch.fill_density(0, [[0, 0, 0]] * 3)
    batch.close()
