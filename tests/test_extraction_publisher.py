"""Bounded extraction publisher: the reference's own unit tests
(PV/src/extraction.rs:731-868) replayed through the C ABI and through the oracle restatement, then
random operation sequences C++ vs oracle; on the GPU, commit into the bounded arenas."""
import random

import numpy as np
import pytest

import helio_b200 as H
from helio_b200 import _ffi
from oracle import extraction_publisher as XO
from oracle import oracle as O

PLANET = bytes([1] * 16)


def key(index):
    return H.PlanetPageKey.new(PLANET, H.PageKey(0, (index, 0, 0)))


def counts(vertices, triangles, meshlets):
    return H.SurfaceCounts(vertices, triangles * 3, meshlets)


def reservation(outcome):
    assert outcome.kind == "Reserved", outcome
    return outcome.reservation


# ---- the reference's tests through the C ABI ---------------------------------------------------------------
def test_gpu_requests_preserve_full_generations_and_dirty_masks():
    request = H.GpuExtractionRequest.new(17, 0xFEDC_BA98_7654_3210, H.TRANSITION_FACE_MASK, 0x8000_0000_0000_0001)
    assert request.generation() == 0xFEDC_BA98_7654_3210
    assert request.dirty_microbricks() == 0x8000_0000_0000_0001
    assert request.page_slot == 17 and request.transition_mask == 0x3F
    with pytest.raises(H.TransitionMask) as err:
        H.GpuExtractionRequest.new(0, 0, 0x80, 0)
    assert err.value.mask == 0x80
    assert H.EXTRACTION_REQUEST_DTYPE.itemsize == 32 and H.EXTRACTION_RANGE_DTYPE.itemsize == 32
    assert H.EXTRACTION_COUNTERS_DTYPE.itemsize == 48


def test_byte_plan_is_exact_and_checked_against_device_limits():
    limits = H.ExtractionLimits.new(4, 2, 100, 300, 10)
    plan = limits.allocation_plan()
    assert plan.request_bytes == 2 * 32 and plan.page_range_bytes == 4 * 32
    assert plan.vertex_bytes == 100 * 32 and plan.index_bytes == 300 * 4
    assert plan.meshlet_bytes == 10 * 32 and plan.counter_bytes == 48
    assert plan.total_bytes == 64 + 128 + 3200 + 1200 + 320 + 48
    with pytest.raises(H.ExtractionError) as err:
        limits.validate_device(3_199, 3_199)
    assert err.value.kind == "DeviceBufferLimit" and err.value.name == "terrain vertices" and err.value.requested == 3200
    limits.validate_device(3_200, 3_200)
    for bad in [(0, 1, 1, 1, 1), (2, 3, 1, 1, 1), (2, 0, 1, 1, 1), (2, 1, 0, 1, 1), (2, 1, 1, 0, 1), (2, 1, 1, 1, 0)]:
        with pytest.raises(H.ExtractionError) as err:
            H.ExtractionLimits.new(*bad)
        assert err.value.kind == "InvalidLimits"
    assert H.ExtractionLimits() == H.ExtractionLimits.new(256, 32, 1_048_576, 3_145_728, 32_768)   # Default


def test_replacement_keeps_current_ranges_until_atomic_publication():
    publisher = H.BoundedExtractionPublisher(H.ExtractionLimits.new(2, 1, 20, 60, 4))
    first = reservation(publisher.reserve(key(0), 1, counts(8, 8, 1)))
    publisher.publish(first)
    before = publisher.current(key(0))
    replacement = reservation(publisher.reserve(key(0), 2, counts(10, 10, 1)))
    assert publisher.current(key(0)) == before
    assert publisher.pending(key(0)) == replacement
    assert publisher.counters().used_vertices == 18
    outcome = publisher.publish(replacement)
    assert outcome.kind == "Published" and outcome.replaced == before
    assert publisher.current(key(0)).generation == 2
    assert publisher.counters().used_vertices == 10
    assert publisher.counters().replacements == 1 and publisher.counters().vertex_high_water == 18


def test_stale_publications_and_evictions_cannot_replace_newer_surfaces():
    publisher = H.BoundedExtractionPublisher(H.ExtractionLimits.new(2, 2, 32, 96, 4))
    first = reservation(publisher.reserve(key(-1), 4, counts(8, 4, 1)))
    publisher.publish(first)
    assert publisher.reserve(key(-1), 3, counts(8, 4, 1)) == H.ReservationOutcome("Stale", newest_generation=4)
    assert publisher.evict(key(-1), 3) == H.ExtractionEvictOutcome("Stale", newest_generation=4)
    assert publisher.current(key(-1)).generation == 4
    fifth = reservation(publisher.reserve(key(-1), 5, counts(8, 4, 1)))
    sixth = reservation(publisher.reserve(key(-1), 6, counts(8, 4, 1)))
    assert publisher.publish(fifth) == H.PublicationOutcome("Stale", newest_generation=6)
    publisher.publish(sixth)
    assert publisher.current(key(-1)).generation == 6
    assert publisher.counters().stale_rejected == 3 and publisher.counters().cancellations == 1


def test_capacity_failure_rolls_back_partial_ranges_and_reuses_freed_space():
    publisher = H.BoundedExtractionPublisher(H.ExtractionLimits.new(2, 2, 10, 12, 1))
    with pytest.raises(H.ExtractionError) as err:
        publisher.reserve(key(0), 1, counts(8, 5, 1))
    assert err.value.kind == "ArenaCapacity" and err.value.capacity == "Indices"
    assert publisher.counters().used_vertices == 0
    valid = reservation(publisher.reserve(key(0), 1, counts(8, 4, 1)))
    publisher.publish(valid)
    assert publisher.evict(key(0), 1) == H.ExtractionEvictOutcome("Evicted")
    c = publisher.counters()
    assert (c.used_vertices, c.used_indices, c.used_meshlets) == (0, 0, 0)
    reused = reservation(publisher.reserve(key(1), 1, counts(10, 4, 1)))
    assert reused.allocation.vertices.first == 0


def test_pending_capacity_and_duplicate_generations_are_explicit():
    publisher = H.BoundedExtractionPublisher(H.ExtractionLimits.new(2, 1, 32, 96, 4))
    first = reservation(publisher.reserve(key(0), 1, counts(8, 4, 1)))
    assert publisher.reserve(key(0), 1, counts(8, 4, 1)) == H.ReservationOutcome("DuplicatePending", reservation=first)
    with pytest.raises(H.ExtractionError) as err:
        publisher.reserve(key(1), 1, counts(8, 4, 1))
    assert err.value.kind == "PendingCapacity" and err.value.maximum == 1
    with pytest.raises(H.ExtractionError) as err:
        publisher.reserve(key(0), 1, counts(9, 4, 1))
    assert err.value.kind == "GenerationConflict"


def test_count_validation_and_reservation_errors():
    publisher = H.BoundedExtractionPublisher(H.ExtractionLimits.new(2, 2, 32, 96, 4))
    for bad, kind in [(H.SurfaceCounts(3, 4, 1), "NonTriangleIndexCount"), (H.SurfaceCounts(0, 3, 1), "IncompleteSurfaceCounts"),
                      (H.SurfaceCounts(3, 0, 0), "IncompleteSurfaceCounts"), (H.SurfaceCounts(3, 3, 0), "IncompleteSurfaceCounts")]:
        with pytest.raises(H.ExtractionError) as err:
            publisher.reserve(key(0), 1, bad)
        assert err.value.kind == kind
    empty = reservation(publisher.reserve(key(0), 1, H.SurfaceCounts()))       # an empty surface reserves nothing
    assert empty.allocation == H.SurfaceAllocation()
    ghost = H.ExtractionReservation(key(5), 1, H.SurfaceAllocation())
    with pytest.raises(H.ExtractionError) as err:
        publisher.publish(ghost)
    assert err.value.kind == "ReservationMissing"
    forged = H.ExtractionReservation(key(0), 1, H.SurfaceAllocation(H.ArenaSlice(1, 1)))
    with pytest.raises(H.ExtractionError) as err:
        publisher.publish(forged)
    assert err.value.kind == "ReservationMismatch"
    with pytest.raises(H.ExtractionError) as err:
        publisher.cancel_pending(key(0), 2)
    assert err.value.kind == "ReservationMismatch"
    assert publisher.cancel_pending(key(0), 1) is True and publisher.cancel_pending(key(0), 1) is False
    assert publisher.evict(key(0), 9) == H.ExtractionEvictOutcome("Missing")
    rng = H.SurfaceAllocation(H.ArenaSlice(3, 4), H.ArenaSlice(6, 9), H.ArenaSlice(1, 2)).gpu_range(0x1_0000_0002)
    assert tuple(int(v) for v in rng) == (3, 4, 6, 9, 1, 2, 2, 1)


# ---- the oracle replays the same reference tests (pins the restatement) ---------------------------------------
def test_oracle_restatement_reproduces_the_reference_tests():
    p = XO.Publisher(2, 1, 20, 60, 4)
    kind, first = p.reserve("a", 1, (8, 24, 1))
    p.publish(first)
    before = p.current("a")
    kind, replacement = p.reserve("a", 2, (10, 30, 1))
    assert p.current("a") == before and p.pending("a") == replacement and p.counters()["used_vertices"] == 18
    assert p.publish(replacement) == ("Published", (2, replacement[2]), before)
    assert p.counters()["used_vertices"] == 10

    p = XO.Publisher(2, 2, 32, 96, 4)
    p.publish(p.reserve("k", 4, (8, 12, 1))[1])
    assert p.reserve("k", 3, (8, 12, 1)) == ("Stale", 4) and p.evict("k", 3) == ("Stale", 4)
    fifth, sixth = p.reserve("k", 5, (8, 12, 1))[1], p.reserve("k", 6, (8, 12, 1))[1]
    assert p.publish(fifth) == ("Stale", 6)
    p.publish(sixth)
    assert p.current("k")[0] == 6

    p = XO.Publisher(2, 2, 10, 12, 1)
    with pytest.raises(XO.Error) as err:
        p.reserve("a", 1, (8, 15, 1))
    assert err.value.kind == "ArenaCapacity" and err.value.fields["capacity"] == "Indices" and p.counters()["used_vertices"] == 0
    p.publish(p.reserve("a", 1, (8, 12, 1))[1])
    assert p.evict("a", 1) == ("Evicted", None) and p.counters()["used_indices"] == 0
    assert p.reserve("b", 1, (10, 12, 1))[1][2][0] == (0, 10)

    p = XO.Publisher(2, 1, 32, 96, 4)
    first = p.reserve("a", 1, (8, 12, 1))[1]
    assert p.reserve("a", 1, (8, 12, 1)) == ("DuplicatePending", first)
    with pytest.raises(XO.Error) as err:
        p.reserve("b", 1, (8, 12, 1))
    assert err.value.kind == "PendingCapacity" and err.value.fields["maximum"] == 1


# ---- random sequences: C++ == oracle after every operation --------------------------------------------------
def _alloc_tuple(a):
    return ((a.vertices.first, a.vertices.count), (a.indices.first, a.indices.count), (a.meshlets.first, a.meshlets.count))


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_random_operation_sequences_match_the_oracle(seed):
    rnd = random.Random(seed)
    dims = (6, 3, 600, 1500, 24)
    pub, ora = H.BoundedExtractionPublisher(H.ExtractionLimits.new(*dims)), XO.Publisher(*dims)
    keys = list(range(-3, 5))
    held = []                                      # reservations handed out: (ours, oracle's)
    for step in range(1500):
        k = rnd.choice(keys)
        op = rnd.random()
        generation = rnd.randint(1, 12)
        if op < 0.45:
            tri = rnd.randint(0, 90)
            c = (0, 0, 0) if tri == 0 else (rnd.randint(1, 140), 3 * tri, rnd.randint(1, 6))
            try:
                want = ora.reserve(k, generation, c)
            except XO.Error as e:
                with pytest.raises(H.ExtractionError) as err:
                    pub.reserve(key(k), generation, H.SurfaceCounts(*c))
                assert err.value.kind == e.kind, step
                continue
            got = pub.reserve(key(k), generation, H.SurfaceCounts(*c))
            assert got.kind == want[0], step
            if want[0] in ("Reserved", "DuplicatePending"):
                assert (got.reservation.generation, _alloc_tuple(got.reservation.allocation)) == (want[1][1], want[1][2]), step
                held.append((got.reservation, want[1]))
            elif want[0] == "Current":
                assert (got.current.generation, _alloc_tuple(got.current.allocation)) == want[1], step
            else:
                assert got.newest_generation == want[1], step
        elif op < 0.75 and held:
            ours, theirs = held.pop(rnd.randrange(len(held)))
            try:
                want = ora.publish(theirs)
            except XO.Error as e:
                with pytest.raises(H.ExtractionError) as err:
                    pub.publish(ours)
                assert err.value.kind == e.kind, step
                continue
            got = pub.publish(ours)
            assert got.kind == want[0], step
            if want[0] == "Published":
                assert (got.replaced is None) == (want[2] is None), step
                if want[2] is not None:
                    assert (got.replaced.generation, _alloc_tuple(got.replaced.allocation)) == want[2], step
            else:
                assert got.newest_generation == want[1], step
        elif op < 0.87:
            try:
                want = ora.cancel_pending(k, generation)
            except XO.Error as e:
                with pytest.raises(H.ExtractionError) as err:
                    pub.cancel_pending(key(k), generation)
                assert err.value.kind == e.kind, step
                continue
            assert pub.cancel_pending(key(k), generation) == want, step
        else:
            want = ora.evict(k, generation)
            got = pub.evict(key(k), generation)
            assert (got.kind, got.newest_generation) == want, step
        oc, gc = ora.counters(), pub.counters()
        assert {name: getattr(gc, name) for name in oc} == oc, step
    assert pub.counters().publications > 50 and pub.counters().backpressured > 0 and pub.counters().stale_rejected > 0


# ---- device: commit into the bounded arenas ------------------------------------------------------------------
@pytest.mark.gpu
def test_commit_places_meshes_at_the_reserved_ranges_and_survives_replacement():
    pages = [(0, 0, 0), (0, -1, 0), (1, 0, 0), (-1, 0, 0)]         # sphere: three surface pages and an empty one
    n = len(pages)
    batch = H.ChunkBatchExtractor(0, edge=32, max_chunks=n, max_vertices=4096, max_indices=6144)
    sphere = int(H.ExtractionFixtureKind.Sphere)
    batch.fill_density(sphere, pages)
    batch.extract_regular(None, n)
    ec = batch.counters(n)
    meshes = [O.extract_regular(O.fixture_fill(O.FIELD_SPHERE, p), debug=False) for p in pages]
    assert [int(v) for v in ec["emitted_vertices"]] == [len(m.vertices) for m in meshes]

    limits = H.ExtractionLimits.new(8, 4, 6000, 9000, 200)
    pub = H.BoundedExtractionPublisher(limits)
    pub.attach(batch.ctx)

    def surface_counts(k):
        ni = int(ec["emitted_indices"][k])
        return H.SurfaceCounts(int(ec["emitted_vertices"][k]), ni, H.max_meshlets_for_indices(ni) if ni else 0)

    keys = [H.PlanetPageKey.new(PLANET, H.PageKey(0, p)) for p in pages]
    res = [reservation(pub.reserve(keys[k], 7, surface_counts(k))) for k in range(n)]
    pub.commit(range(n), [5, 1, 2, 0], res)
    for r in res:
        assert pub.publish(r).kind == "Published"
    ranges = pub.read(_ffi.XPUB_PAGE_RANGES, H.EXTRACTION_RANGE_DTYPE)
    verts = pub.read(_ffi.XPUB_VERTICES, H.VERTEX_DTYPE)
    idx = pub.read(_ffi.XPUB_INDICES, np.uint32)
    first_v = 0
    for k, slot in enumerate([5, 1, 2, 0]):
        r = ranges[slot]
        nv = len(meshes[k].vertices)                                # an empty surface reserves ArenaSlice::default()
        assert (r["first_vertex"], r["vertex_count"], r["index_count"]) == (first_v if nv else 0, nv, len(meshes[k].indices))
        assert (int(r["generation_low"]), int(r["generation_high"])) == (7, 0)
        assert verts[r["first_vertex"]:r["first_vertex"] + r["vertex_count"]].tobytes() == meshes[k].vertices.tobytes()
        assert np.array_equal(idx[r["first_index"]:r["first_index"] + r["index_count"]], meshes[k].indices)
        first_v += len(meshes[k].vertices)
    dc = pub.read(_ffi.XPUB_COUNTERS, H.EXTRACTION_COUNTERS_DTYPE)[0]
    assert (dc["requests"], dc["completed"], dc["overflowed"]) == (n, n, 0)
    assert dc["vertices"] == sum(len(m.vertices) for m in meshes)

    # replacement: generation 8 of page 0 lands in NEW ranges while generation 7 stays readable
    newer = reservation(pub.reserve(keys[0], 8, surface_counts(0)))
    assert newer.allocation.vertices.first == first_v
    pub.commit([0], [5], [newer])
    verts = pub.read(_ffi.XPUB_VERTICES, H.VERTEX_DTYPE)
    old = res[0].allocation.vertices
    assert verts[old.first:old.first + old.count].tobytes() == meshes[0].vertices.tobytes()
    assert verts[first_v:first_v + old.count].tobytes() == meshes[0].vertices.tobytes()
    assert pub.publish(newer).replaced.generation == 7
    assert pub.counters().used_vertices == first_v          # the old ranges were recycled only now

    # a reservation that does not match what the chunk emitted is counted, not copied
    wrong = reservation(pub.reserve(keys[1], 9, H.SurfaceCounts(3, 3, 1)))
    pub.commit([1], [1], [wrong])
    dc = pub.read(_ffi.XPUB_COUNTERS, H.EXTRACTION_COUNTERS_DTYPE)[0]
    assert dc["overflowed"] == 1 and dc["requests"] == n + 2
    assert int(pub.read(_ffi.XPUB_PAGE_RANGES, H.EXTRACTION_RANGE_DTYPE)[1]["generation_low"]) == 7
    pub.close()
    batch.close()
