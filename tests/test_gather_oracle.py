"""CPU: the gather oracle against the reference's own known answers.

PV/src/surface_sampling.rs:474-752: the 27-page dependency set, the signed fine-page dependencies, the
LOD0 rule, the 480,424-byte payload, and the GPU gather test itself -- target PageKey(2, [-1,-2,1]),
all six faces, a frame origin ([32,0,0]) that is NOT aligned to the coarse page, every gathered word
equal to canonical_cell(position), counters (39304, 80802, 0 misses, not stale, completed) and the
eight indirect dispatches."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import residency as R

PLANET = R.planet_words(bytes([0x5A] * 16))
ORIGIN = [32, 0, 0]


def reference_scene(lod=2, page=(-1, -2, 1), mask=0x3F, drop=None, tiles=(8, 8, 4)):
    deps = sorted(R.required_pages(lod, page, mask))
    atlas = R.Atlas(tiles, 256, 64, epoch=1)
    slots = {}
    for i, (l, xyz) in enumerate(deps):
        if drop is not None and (l, xyz) == drop:
            continue
        gen = (1 << 32 | 1) if (l, xyz) == (lod, tuple(page)) else (1 << 32 | (i + 2))
        slots[(l, xyz)] = atlas.upload(PLANET, l, xyz, ORIGIN, R.canonical_page_cells(l, xyz), gen)
    job = R.make_job(PLANET, lod, page, ORIGIN, 1 << 32 | 1, mask, slots.get((lod, tuple(page)), 0), atlas.epoch)
    return atlas, job, deps


def test_dependency_sets_match_the_reference():
    pages = R.required_pages(4, (-2, 5, -7), 0)
    assert len(pages) == 27 and (4, (-2, 5, -7)) in pages and (4, (-3, 4, -8)) in pages and (4, (-1, 6, -6)) in pages
    pages = R.required_pages(3, (-1, 0, -1), 0b100001)   # NegativeX | PositiveZ
    assert any(l == 2 for l, _ in pages) and any(p[0] < 0 for _, p in pages) and any(p[2] >= 0 for _, p in pages)
    with pytest.raises(ValueError):
        R.required_pages(0, (0, 0, 0), 1)
    assert (34 ** 3 + 6 * 3 * 67 * 67) * 4 == 480_424


def test_hash_and_table_semantics():
    rel = R.relative_lod0_cell_min(2, (-1, -2, 1), ORIGIN)
    assert R.key_hash(PLANET, rel, 2) == O.page_hash(PLANET, rel, 2)
    assert R.relative_lod0_cell_min(0, (3, -1, 1), [64, -64, 0]) == [32, 32, 32]
    t = R.PageTable(8, 8)
    keys = [(PLANET, [32 * i, 0, 0], 0) for i in range(6)]
    for i, k in enumerate(keys):
        t.insert(*k, slot=i, generation=i + 1)
    assert t.occupied == 6 and all(t.lookup(*k)[1]["slot"] == i for i, k in enumerate(keys))
    assert t.remove(*keys[2]) and t.lookup(*keys[2]) is None and t.tombstones == 1
    assert t.lookup(*keys[5])[1]["slot"] == 5            # lookups walk over the tombstone
    t.insert(*keys[2], slot=9, generation=7)             # and inserts reuse it
    assert t.tombstones == 0 and t.lookup(*keys[2])[1]["slot"] == 9
    with pytest.raises(ValueError):
        R.PageTable(6, 2)


def test_gather_matches_canonical_pages_across_signed_regular_and_transition_boundaries():
    atlas, job, deps = reference_scene()
    assert R.relative_lod0_cell_min(2, (-1, -2, 1), ORIGIN)[0] % R.lod0_cell_span(2) != 0   # the regression's premise
    regular, transition, c, indirect = O.gather_surface(atlas.residency(), atlas.table.entries, atlas.words, job)
    assert np.array_equal(regular, R.expected_regular(2, (-1, -2, 1)))
    assert np.array_equal(transition, R.expected_transition(2, (-1, -2, 1)))
    assert (c["regular_samples"], c["transition_samples"], c["page_misses"], c["stale_targets"], c["completed"]) == \
        (39304, 80802, 0, 0, 1)
    assert c["table_probes"] >= 39304 + 80802 + 1
    assert list(indirect) == [512, 1, 1, 128, 1, 1, 1, 1, 1, 512, 1, 1, 96, 1, 1, 24, 1, 1, 1, 1, 1, 96, 1, 1]


def test_misses_stale_targets_and_epochs():
    atlas, job, deps = reference_scene(drop=(2, (-2, -2, 1)))
    regular, _, c, indirect = O.gather_surface(atlas.residency(), atlas.table.entries, atlas.words, job)
    want = R.expected_regular(2, (-1, -2, 1)).reshape(34, 34, 34).copy()
    want[1:33, 1:33, 0] = 0x7FFF                         # the -x neighbour's face of the halo reads as AIR
    assert np.array_equal(regular.reshape(34, 34, 34)[1:33, 1:33, :], want[1:33, 1:33, :])
    assert c["page_misses"] > 0 and c["completed"] == 0 and c["stale_targets"] == 0 and not indirect.any()
    atlas, job, _ = reference_scene(mask=0)
    job["generation_low"] += 1                           # the resident page is newer than the request
    _, transition, c, _ = O.gather_surface(atlas.residency(), atlas.table.entries, atlas.words, job)
    assert c["stale_targets"] == 1 and c["completed"] == 0 and c["regular_samples"] == 39304 and c["transition_samples"] == 0
    assert (transition == 0xFFFFFFFF).all()
    atlas, job, _ = reference_scene(mask=0)
    job["residency_epoch_low"] += 1                      # residency moved on: nothing is gathered
    regular, _, c, _ = O.gather_surface(atlas.residency(), atlas.table.entries, atlas.words, job)
    assert c["regular_samples"] == 0 and c["stale_targets"] == 1 and (regular == 0xFFFFFFFF).all()
