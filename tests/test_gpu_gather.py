"""GPU parity: surface gather through the C ABI vs the CPU oracle and the reference's golden field.

Mirrors PV/src/surface_sampling.rs:569-752 (gpu_gather_matches_canonical_pages_across_signed_regular_and_
transition_boundaries) and adds misses, stale targets, epochs, tombstones, batches and the fused
gather -> extract call stack."""
import numpy as np
import pytest

import helio_b200 as H
from helio_b200 import _ffi
from oracle import oracle as O
from oracle import residency as R
from hvx_testutil import assert_vertices_equal
from test_gather_oracle import ORIGIN, PLANET, reference_scene

pytestmark = pytest.mark.gpu


def make_ctx(n=1, transition=True):
    return H.Context(0, edge=32, max_chunks=n, max_vertices=393_216, max_indices=491_520,
                     max_transition_vertices=73_728 if transition else 0, max_transition_indices=221_184 if transition else 0)


def check_job(sampler, k, atlas, job, mask):
    regular, transition, c, indirect = O.gather_surface(atlas.residency(), atlas.table.entries, atlas.words, job)
    got_c = sampler.counters_buffer()[k]
    for name in ("regular_samples", "transition_samples", "table_probes", "page_misses", "stale_targets", "completed"):
        assert int(got_c[name]) == int(c[name]), (k, name, int(got_c[name]), int(c[name]))
    assert np.array_equal(sampler.indirect_buffer()[k].reshape(-1), indirect), k
    if c["regular_samples"]:
        assert np.array_equal(sampler.regular_samples(k), regular), k
    got_t = sampler.transition_samples(k) if mask else None
    for face in range(6):
        if (mask >> face) & 1 and c["transition_samples"]:
            assert np.array_equal(got_t[face * 13467:(face + 1) * 13467], transition[face * 13467:(face + 1) * 13467]), (k, face)
    return c


def test_gather_matches_canonical_pages_across_signed_regular_and_transition_boundaries():
    atlas, job, _ = reference_scene()
    ctx = make_ctx()
    sampler = H.GpuSurfaceSampler(ctx)
    sampler.dispatch(atlas.residency(), atlas.table.entries, atlas.words, job)
    c = check_job(sampler, 0, atlas, job[0:1], 0x3F)
    assert (c["regular_samples"], c["transition_samples"], c["page_misses"], c["stale_targets"], c["completed"]) == (39304, 80802, 0, 0, 1)
    assert np.array_equal(sampler.regular_samples(0), R.expected_regular(2, (-1, -2, 1)))
    assert np.array_equal(sampler.transition_samples(0), R.expected_transition(2, (-1, -2, 1)))
    assert sampler.indirect_buffer()[0].reshape(-1).tolist() == [512, 1, 1, 128, 1, 1, 1, 1, 1, 512, 1, 1, 96, 1, 1, 24, 1, 1, 1, 1, 1, 96, 1, 1]
    ctx.close()


def test_misses_stale_targets_and_epochs():
    ctx = make_ctx()
    sampler = H.GpuSurfaceSampler(ctx)
    atlas, job, _ = reference_scene(drop=(2, (-2, -2, 1)))
    sampler.dispatch(atlas.residency(), atlas.table.entries, atlas.words, job)
    c = check_job(sampler, 0, atlas, job, 0x3F)
    assert c["page_misses"] > 0 and c["completed"] == 0
    atlas, job, _ = reference_scene(mask=0)
    job["generation_low"] += 1
    sampler.dispatch(atlas.residency(), atlas.table.entries, atlas.words, job)
    assert check_job(sampler, 0, atlas, job, 0)["stale_targets"] == 1
    atlas, job, _ = reference_scene(mask=0)
    job["residency_epoch_low"] += 1
    ctx.write(_ffi.BUF_SAMPLES, np.full(34 ** 3, 0x12345678, dtype=np.uint32))
    sampler.dispatch(atlas.residency(), atlas.table.entries, atlas.words, job)
    c = check_job(sampler, 0, atlas, job, 0)
    assert c["regular_samples"] == 0 and c["stale_targets"] == 1
    assert (sampler.regular_samples(0) == 0x12345678).all()      # a stale epoch gathers nothing
    ctx.close()


def test_batch_with_tombstones_mixed_lods_and_a_device_resident_atlas():
    torch = pytest.importorskip("torch")
    targets = [(2, (-1, -2, 1), 0x3F), (0, (3, -1, 0), 0), (1, (0, 0, -1), 0x15), (3, (-1, 0, 0), 0x2A), (1, (-2, -1, 1), 0x01)]
    atlas = R.Atlas((12, 12, 4), 1024, 48, epoch=(7 << 32) | 3)
    slots, gens = {}, {}
    order = []
    for lod, page, mask in targets:
        for dep in sorted(R.required_pages(lod, page, mask)):
            if dep not in slots:
                order.append(dep)
                slots[dep] = None
    rng = np.random.default_rng(5)
    rng.shuffle(order)
    for i, (l, xyz) in enumerate(order):
        gens[(l, xyz)] = (3 << 32) | (i + 1)
        slots[(l, xyz)] = atlas.upload(PLANET, l, xyz, ORIGIN, R.canonical_page_cells(l, xyz), gens[(l, xyz)])
    # churn: remove and re-insert a few unrelated keys so probe sequences cross tombstones
    for i in range(40):
        atlas.table.insert(PLANET, [1 << 20, 32 * i, 0], 0, 1000 + i, 1)
    for i in range(0, 40, 2):
        atlas.table.remove(PLANET, [1 << 20, 32 * i, 0], 0)
    jobs = np.concatenate([R.make_job(PLANET, l, p, ORIGIN, gens[(l, tuple(p))], m, slots[(l, tuple(p))], atlas.epoch) for l, p, m in targets])
    ctx = make_ctx(len(targets))
    sampler = H.GpuSurfaceSampler(ctx)
    sampler.dispatch(atlas.residency(), atlas.table.entries, torch.from_numpy(atlas.words.reshape(-1).view(np.int32)).cuda(), jobs)
    for k, (lod, page, mask) in enumerate(targets):
        c = check_job(sampler, k, atlas, jobs[k:k + 1], mask)
        assert c["completed"] == 1 and c["page_misses"] == 0, k
        assert np.array_equal(sampler.regular_samples(k), R.expected_regular(lod, page)), k
    ctx.close()


def test_gather_then_extract_equals_extracting_the_canonical_block():
    atlas, job, _ = reference_scene()
    ctx = make_ctx()
    sampler = H.GpuSurfaceSampler(ctx)
    sampler.dispatch(atlas.residency(), atlas.table.entries, atlas.words, job)
    sampler.extract()
    want = O.extract_regular(R.expected_regular(2, (-1, -2, 1)), edge=32, generation=(1 << 32) | 1, transition_mask=0x3F, debug=False)
    counters = ctx.read(_ffi.BUF_REGULAR_COUNTERS, 0, 1)[0]
    assert counters["completed"] == 1 and counters["emitted_vertices"] == len(want.vertices) > 0
    verts, idx, _ = ctx.read_meshes(0, 0, 1)
    assert_vertices_equal(verts, want.vertices, "gathered block")
    assert np.array_equal(idx, want.indices)
    want_t = O.extract_transition(R.expected_transition(2, (-1, -2, 1)), 0x3F, edge=32, generation=(1 << 32) | 1, debug=False)
    tv, ti, _ = ctx.read_meshes(1, 0, 1)
    assert_vertices_equal(tv, want_t.vertices, "gathered slabs")
    assert np.array_equal(ti, want_t.indices)
    ctx.close()


def test_bound_page_table_stays_resident():
    """hvx_gather_bind_table: the table is uploaded once, dispatches with table = NULL use it; rebinding replaces it."""
    atlas, job, _ = reference_scene()
    ctx = make_ctx()
    sampler = H.GpuSurfaceSampler(ctx)
    with pytest.raises(H.HvxError, match="no table is bound"):
        sampler.dispatch(atlas.residency(), None, atlas.words, job)
    sampler.bind_table(atlas.table.entries)
    for _ in range(2):
        ctx.write(_ffi.BUF_SAMPLES, np.zeros(34 ** 3, dtype=np.uint32))
        sampler.dispatch(atlas.residency(), None, atlas.words, job)
        assert np.array_equal(sampler.regular_samples(0), R.expected_regular(2, (-1, -2, 1)))
        assert check_job(sampler, 0, atlas, job[0:1], 0x3F)["completed"] == 1
    dropped, job2, _ = reference_scene(drop=(2, (-2, -2, 1)))
    sampler.bind_table(dropped.table.entries)       # a new publication: one more upload
    sampler.dispatch(dropped.residency(), None, dropped.words, job2)
    assert check_job(sampler, 0, dropped, job2, 0x3F)["page_misses"] > 0
    sampler.bind_table(None)
    with pytest.raises(H.HvxError, match="no table is bound"):
        sampler.dispatch(atlas.residency(), None, atlas.words, job)
    ctx.close()


def test_errors():
    atlas, job, _ = reference_scene(mask=0)
    with pytest.raises(ValueError):
        H.GpuSurfaceSampler(H.Context(0, edge=64, max_chunks=1, max_vertices=8, max_indices=8))
    ctx = make_ctx(1, transition=False)
    sampler = H.GpuSurfaceSampler(ctx)
    bad = job.copy()
    bad["transition_mask"] = 0x40
    with pytest.raises(H.TransitionMask):
        sampler.dispatch(atlas.residency(), atlas.table.entries, atlas.words, bad)
    bad = job.copy()
    bad["lod"], bad["transition_mask"] = 0, 1
    with pytest.raises(H.FinestLodHasNoFinerNeighbor):
        sampler.dispatch(atlas.residency(), atlas.table.entries, atlas.words, bad)
    with pytest.raises(H.SampleCount):
        sampler.dispatch(atlas.residency(), atlas.table.entries, atlas.words.reshape(-1)[:-32768], job)
    with pytest.raises(H.BatchCapacity):
        sampler.dispatch(atlas.residency(), atlas.table.entries, atlas.words, np.concatenate([job, job]))
    ctx.close()
