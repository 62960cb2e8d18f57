"""GPU parity: surface publication through the C ABI vs the oracle and the reference's known answers.

Mirrors PV/tests/gpu_surface_publication.rs:182-520 and adds the mesh copies into the inactive bank, a
batch of jobs, and the whole device-resident call stack gather -> extract -> publish."""
import numpy as np
import pytest

import helio_b200 as H
from helio_b200 import _ffi
from oracle import oracle as O
from oracle import publication as P
from oracle import residency as R
from test_gather_oracle import ORIGIN, PLANET, reference_scene
from test_publish_oracle import reference_fixture, run_reference_sequence

pytestmark = pytest.mark.gpu


def test_publication_is_atomic_generation_safe_and_visibility_gated():
    ctx = H.Context(0, edge=32, max_chunks=1, max_vertices=32, max_indices=64, max_transition_vertices=16, max_transition_indices=48)
    pub = H.SurfacePublisher(ctx, 1)
    ref, _, _, _, _ = reference_fixture()
    pub.write(_ffi.PUB_STATES, ref.states.view(H.SURFACE_STATE_DTYPE))
    pub.write(_ffi.PUB_REGULAR_DRAWS, ref.regular_draws.view(H.DRAW_INDEXED_INDIRECT_DTYPE))
    pub.write(_ffi.PUB_TRANSITION_DRAWS, ref.transition_draws.view(H.DRAW_INDEXED_INDIRECT_DTYPE))
    job = pub.surface_job(0, 43)

    def publish(meta, regular, transition):
        ctx.write(_ffi.BUF_REGULAR_COUNTERS, regular.view(H.EMISSION_COUNTERS_DTYPE))
        ctx.write(_ffi.BUF_TRANSITION_COUNTERS, transition.view(H.TRANSITION_COUNTERS_DTYPE))
        pub.publish(job, [0], meta.view(H.PAGE_META_DTYPE))

    def refresh(visible):
        pages = np.zeros(1, dtype=H.DRAW_PAGE_DTYPE)
        pages["lod0_cell_size_m"], pages["visible"] = 0.1, visible
        pub.refresh_visibility(pages)

    def read():
        return (pub.surface_states()[0].view(P.STATE_DTYPE), pub.regular_draws()[0].view(P.DRAW_DTYPE),
                pub.transition_draws()[0].view(P.DRAW_DTYPE), pub.feedback())

    run_reference_sequence(publish, refresh, read)
    pub.close()
    ctx.close()


def test_gather_extract_publish_call_stack_matches_the_oracle():
    """Three slots, two jobs (one current, one whose page metadata moved on), twice in a row: banks flip."""
    atlas, gjob, _ = reference_scene()
    n_slots = 3
    ctx = H.Context(0, edge=32, max_chunks=2, max_vertices=40_000, max_indices=60_000, max_transition_vertices=8_192,
                    max_transition_indices=24_576)
    sampler = H.GpuSurfaceSampler(ctx)
    pub = H.SurfacePublisher(ctx, n_slots)
    want = P.Publisher(n_slots, 40_000, 60_000, 8_192, 24_576)
    jobs2 = np.concatenate([gjob, gjob])
    generation = (1 << 32) | 1
    meta = np.zeros(n_slots, dtype=P.PAGE_META_DTYPE)
    meta[2] = (tuple(gjob[0]["relative_lod0_cell_min"]), 2, 2, generation & 0xFFFFFFFF, generation >> 32, 0x3F)
    meta[0] = ((0, 0, 0), 2, 0, 99, 0, 0)                                      # slot 0 holds a newer page: stale for our job
    pages = np.zeros(n_slots, dtype=H.DRAW_PAGE_DTYPE)
    pages["visible"] = [1, 0, 1]
    mesh = O.extract_regular(R.expected_regular(2, (-1, -2, 1)), edge=32, generation=generation, transition_mask=0x3F, debug=False)
    tmesh = O.extract_transition(R.expected_transition(2, (-1, -2, 1)), 0x3F, edge=32, generation=generation, debug=False)
    for round_ in range(2):
        sampler.dispatch(atlas.residency(), atlas.table.entries, atlas.words, jobs2)
        sampler.extract()
        sjobs = np.concatenate([pub.surface_job(2, generation, 0x3F), pub.surface_job(0, generation, 0x3F)])
        pub.publish(sjobs, [0, 1], meta.view(H.PAGE_META_DTYPE))
        pub.refresh_visibility(pages)
        rc = ctx.read(_ffi.BUF_REGULAR_COUNTERS, 0, 2)
        tc = ctx.read(_ffi.BUF_TRANSITION_COUNTERS, 0, 2)
        assert rc["emitted_vertices"][0] == len(mesh.vertices) and tc["emitted_vertices"][0] == len(tmesh.vertices)
        for k, slot in enumerate((2, 0)):
            want.publish(want.job(slot, generation, 0x3F), meta, rc[k:k + 1], tc[k:k + 1], mesh.vertices, mesh.indices,
                         tmesh.vertices, tmesh.indices)
        want.refresh_visibility(pages["visible"])
        assert pub.surface_states().tobytes() == want.states.tobytes()
        assert pub.regular_draws().tobytes() == want.regular_draws.tobytes()
        assert pub.transition_draws().tobytes() == want.transition_draws.tobytes()
        assert pub.feedback().tobytes() == want.feedback[0].tobytes()
        state = pub.surface_states()[2]
        assert state["valid"] == 1 and state["active_bank"] == 1 - round_ and pub.surface_states()[0]["valid"] == 0
        bank = 2 * 2 + int(state["active_bank"])
        got_v = pub.read(_ffi.PUB_REGULAR_VERTICES, bank * 40_000, len(mesh.vertices))
        got_i = pub.read(_ffi.PUB_REGULAR_INDICES, bank * 60_000, len(mesh.indices))
        assert got_v.tobytes() == want.vertices[bank * 40_000:bank * 40_000 + len(mesh.vertices)].tobytes() == mesh.vertices.tobytes()
        assert np.array_equal(got_i, mesh.indices)
        got_tv = pub.read(_ffi.PUB_TRANSITION_VERTICES, bank * 8_192, len(tmesh.vertices))
        got_ti = pub.read(_ffi.PUB_TRANSITION_INDICES, bank * 24_576, len(tmesh.indices))
        assert got_tv.tobytes() == tmesh.vertices.tobytes() and np.array_equal(got_ti, tmesh.indices)
        draw = pub.regular_draws()[2]
        assert (draw["index_count"], draw["instance_count"], draw["first_index"], draw["base_vertex"], draw["first_instance"]) == \
            (len(mesh.indices), 1, bank * 60_000, bank * 40_000, 2)
    assert pub.feedback()["published_jobs"] == 2 and pub.feedback()["stale_rejections"] == 2
    pub.close()
    ctx.close()


def test_errors():
    ctx = H.Context(0, edge=32, max_chunks=2, max_vertices=64, max_indices=96)
    pub = H.SurfacePublisher(ctx, 2)
    meta = np.zeros(2, dtype=H.PAGE_META_DTYPE)
    twice = np.concatenate([pub.surface_job(1, 5), pub.surface_job(1, 5)])
    with pytest.raises(H.HvxError):
        pub.publish(twice, [0, 1], meta)                     # one slot twice in a batch
    with pytest.raises(H.HvxError):
        pub.publish(pub.surface_job(2, 5), [0], meta)        # slot out of range
    bad = pub.surface_job(0, 5)
    bad["regular_max_vertices"] = 63
    with pytest.raises(H.InvalidExtractionCapacity):
        pub.publish(bad, [0], meta)                          # job built for other bank capacities
    with pytest.raises(H.BatchCapacity):
        pub.publish(pub.surface_job(0, 5), [7], meta)        # chunk outside the extraction batch
    # a ctx without transition capacity publishes with empty, completed transition counters
    ctx.write(_ffi.BUF_REGULAR_COUNTERS, np.array([(3, 3, 3, 3, 0, 0, 1, 0)], dtype=H.EMISSION_COUNTERS_DTYPE))
    meta[0] = ((0, 0, 0), 0, 0, 5, 0, 0)
    pub.publish(pub.surface_job(0, 5), [0], meta)
    s = pub.surface_states()[0]
    assert s["valid"] == 1 and s["regular_vertex_count"] == 3 and s["transition_index_count"] == 0 and pub.feedback()["published_jobs"] == 1
    pub.close()
    ctx.close()
