"""GPU parity: regular-cell extraction through the C ABI vs the CPU oracle.

Mirrors PV/tests/gpu_transvoxel.rs:10-136 and PV/tests/gpu_transvoxel_emission.rs:12-264, with the
float tolerance tightened from the reference's 1e-5 / 2e-4 to bit equality.
"""
import numpy as np
import pytest

import helio_b200 as H
from oracle import oracle as O
from hvx_testutil import ALL, FIXTURE_PAGES, assert_vertices_equal, check_offsets

pytestmark = pytest.mark.gpu


# Every pinned case runs three ways: the decoupled kernel alone (what the bench times and production runs), the
# decoupled kernel plus the per-cell record kernels (HVX_CFG_DEBUG_RECORDS), and the first-generation kernel
# (HVX_CFG_FIRST_GENERATION, an independent second implementation that writes the records itself).
VARIANTS = {"decoupled": dict(debug_records=False), "records": dict(debug_records=True),
            "first_generation": dict(debug_records=True, first_generation=True)}


@pytest.fixture(scope="module", params=list(VARIANTS))
def extractor(request):
    ex = H.TransvoxelGpuExtractor(0, H.TransvoxelGpuExtractorConfig(), **VARIANTS[request.param])
    ex.has_records = VARIANTS[request.param]["debug_records"]
    yield ex
    ex.close()


def _check_dispatch(ex, samples, generation, dirty, mask, label, edge=32):
    want = O.extract_regular(samples, edge=edge, generation=generation, dirty_microbricks=dirty, transition_mask=mask)
    ex.dispatch(samples, generation, dirty, mask)
    counters = ex.counters_buffer()
    assert counters["completed"] == 1, label
    assert counters["vertex_overflow"] == 0 and counters["index_overflow"] == 0, label
    assert counters["required_vertices"] == len(want.vertices), label
    assert counters["required_indices"] == len(want.indices), label
    assert counters["emitted_vertices"] == counters["required_vertices"]
    assert counters["emitted_indices"] == counters["required_indices"]
    got_v = ex.vertices_buffer(len(want.vertices))
    got_i = ex.indices_buffer(len(want.indices))
    assert_vertices_equal(got_v, want.vertices, label)
    assert np.array_equal(got_i, want.indices), f"{label}: indices"
    cls = ex.classify_counters_buffer()
    assert [int(cls[k]) for k in ("visited_cells", "active_cells", "vertices", "triangles")] == list(want.classify), label
    return want, got_v, got_i


@pytest.mark.parametrize("name", list(FIXTURE_PAGES))
def test_emission_matches_cpu_geometry_on_every_fixture(extractor, name):
    kind, page = FIXTURE_PAGES[name]
    generation = 1000 + kind
    samples = O.fixture_fill(kind, page)
    want, got_v, got_i = _check_dispatch(extractor, samples, generation, ALL, 0, name)
    # per-cell records and ranges (debug outputs a6 / a11)
    if extractor.has_records:
        cells = extractor.cells_buffer()
        assert np.array_equal(cells["packed_case_class_counts"], want.cell_words[:, 0]), name
        assert np.array_equal(cells["generation_low"], want.cell_words[:, 1]), name
        check_offsets(extractor.offsets_buffer(), extractor.blocks_buffer(), want.cell_ranges, generation, name)
    # byte-identical on repeat (gpu_transvoxel_emission.rs:125-151)
    extractor.dispatch(samples, generation, ALL, 0)
    assert extractor.vertices_buffer(len(got_v)).tobytes() == got_v.tobytes()
    assert np.array_equal(extractor.indices_buffer(len(got_i)), got_i)


def test_published_fixture_counts(extractor):
    """docs/planetary_voxel_extraction_benchmark.md:59-70 vertex / triangle counts."""
    published = {"plane": (4096, 2048), "sphere": (1323, 661), "cave": (1311, 655), "sharp_corner": (3, 1),
                 "thin_slab": (4096, 2048), "material_seam": (4096, 2048)}
    for name, (kind, page) in FIXTURE_PAGES.items():
        extractor.dispatch(O.fixture_fill(kind, page), 5, ALL, 0)
        c = extractor.counters_buffer()
        assert (int(c["emitted_vertices"]), int(c["emitted_indices"]) // 3) == published[name], name


def test_dirty_microbrick_subset(extractor):
    """gpu_transvoxel_emission.rs:154-214: only microbrick 12 is revisited; other cells keep stale records."""
    samples = O.fixture_fill(O.FIELD_PLANE, [0, -1, 0])
    extractor.dispatch(samples, 1400, ALL, 0)
    want, _, _ = _check_dispatch(extractor, samples, 1500, 1 << 12, 0, "dirty")
    assert len(want.vertices) > 0
    if extractor.has_records:
        gen = check_offsets(extractor.offsets_buffer(), extractor.blocks_buffer(), want.cell_ranges, 1500, "dirty")
        visited = want.cell_ranges[:, 0] != 0xFFFFFFFF
        assert visited.sum() == 512 and np.all(gen[~visited] != 1500)
        cells = extractor.cells_buffer()
        assert np.all(cells["generation_low"][visited] == 1500) and np.all(cells["generation_low"][~visited] == 1400)
    cls = extractor.classify_counters_buffer()
    assert cls["visited_cells"] == 8 * 8 * 8 and cls["active_cells"] > 0


def test_secondary_position_on_transition_face(extractor):
    """gpu_transvoxel_emission.rs:216-247: +X boundary vertices move to x = 31.75."""
    samples = O.fixture_fill(O.FIELD_PLANE, [0, -1, 0])
    mask = H.TransitionFace.PositiveX.bit()
    _, got_v, _ = _check_dispatch(extractor, samples, 1750, ALL, mask, "secondary")
    x, z = got_v["position"][:, 0], got_v["position"][:, 2]
    interior = (z >= 1.0) & (z < 31.0)
    assert np.any((np.abs(x - 31.75) <= 1e-5) & interior)
    assert not np.any((np.abs(x - 32.0) <= 1e-5) & interior)


@pytest.mark.parametrize("mask", [0x3F, 0x15, 0x2A, 0x01, 0x24])
def test_secondary_positions_all_faces(extractor, mask):
    for kind, page in [(O.FIELD_SPHERE, [0, 0, 0]), (O.FIELD_SPHERE, [-1, -1, -1]), (O.FIELD_CAVE, [-1, 0, -1])]:
        _check_dispatch(extractor, O.fixture_fill(kind, page), 3, ALL, mask, f"mask {mask:#x} {page}")


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("capacity", [(1, 1), (4095, 6144), (4096, 6143), (100, 1_000_000), (1_000_000, 100)])
def test_overflow_contract(variant, capacity):
    """gpu_transvoxel_emission.rs:249-262: capacity (1,1) suppresses the emission, reports the need; one element
    short in either arena does the same (and a neighbouring guard word stays untouched)."""
    tiny = H.TransvoxelGpuExtractor(0, H.TransvoxelGpuExtractorConfig.new(*capacity), **VARIANTS[variant])
    tiny.dispatch(O.fixture_fill(O.FIELD_PLANE, [0, -1, 0]), 2000, ALL, 0)
    c = tiny.counters_buffer()
    assert c["completed"] == 1
    assert (c["vertex_overflow"] != 0) == (capacity[0] < 4096) and (c["index_overflow"] != 0) == (capacity[1] < 6144)
    assert c["emitted_vertices"] == 0 and c["emitted_indices"] == 0
    assert c["required_vertices"] == 4096 and c["required_indices"] == 6144
    r = tiny.context.read(H._ffi.BUF_REGULAR_RANGES, 0, 1)[0]
    assert r["vertex_count"] == 0 and r["index_count"] == 0
    tiny.close()


@pytest.mark.parametrize("variant", list(VARIANTS))
def test_exact_capacity_is_not_an_overflow(variant):
    ex = H.TransvoxelGpuExtractor(0, H.TransvoxelGpuExtractorConfig.new(4096, 6144), **VARIANTS[variant])
    samples = O.fixture_fill(O.FIELD_PLANE, [0, -1, 0])
    want = O.extract_regular(samples, generation=3)
    ex.dispatch(samples, 3, ALL, 0)
    c = ex.counters_buffer()
    assert c["vertex_overflow"] == 0 and c["index_overflow"] == 0 and c["emitted_vertices"] == 4096
    assert_vertices_equal(ex.vertices_buffer(4096), want.vertices, "exact capacity")
    assert np.array_equal(ex.indices_buffer(6144), want.indices)
    ex.close()


def test_errors_mirror_the_reference(extractor):
    samples = O.fixture_fill(O.FIELD_PLANE, [0, -1, 0])
    with pytest.raises(H.SampleCount) as info:
        extractor.dispatch(samples[:-1], 1, ALL, 0)
    assert info.value.actual == 34 ** 3 - 1 and info.value.expected == 34 ** 3
    with pytest.raises(H.InvalidExtractionCapacity):
        H.TransvoxelGpuExtractorConfig.new(0, 5)
    with pytest.raises(H.InvalidExtractionCapacity):
        H.Context(0, max_vertices=0, max_indices=4)
    stats = extractor.resource_stats()
    extractor.resize(3840, 2160)
    assert extractor.resource_stats() == stats


@pytest.mark.parametrize("first_generation", [False, True])
def test_classifier_matches_cpu(first_generation):
    """gpu_transvoxel.rs:10-136: every GpuTransvoxelCell + counters, determinism, dirty subset."""
    clf = H.TransvoxelGpuClassifier(0, first_generation=first_generation)
    for name, (kind, page) in FIXTURE_PAGES.items():
        samples = O.fixture_fill(kind, page)
        want = O.extract_regular(samples, generation=10 + kind)
        clf.dispatch(samples, 10 + kind, ALL)
        cells = clf.output_buffer()
        assert np.array_equal(cells["packed_case_class_counts"], want.cell_words[:, 0]), name
        assert np.array_equal(cells["generation_low"], want.cell_words[:, 1]), name
        c = clf.counters_buffer()
        assert [int(c[k]) for k in c.dtype.names] == list(want.classify), name
        clf.dispatch(samples, 10 + kind, ALL)
        assert clf.output_buffer().tobytes() == cells.tobytes()
    samples = O.fixture_fill(O.FIELD_PLANE, [0, -1, 0])
    want = O.extract_regular(samples, generation=100, dirty_microbricks=1 << 12)
    clf.dispatch(samples, 100, 1 << 12)
    c = clf.counters_buffer()
    assert [int(c[k]) for k in c.dtype.names] == list(want.classify) and c["visited_cells"] == 512
    with pytest.raises(H.SampleCount):
        clf.dispatch(samples[:-1], 100, ALL)
    clf.close()


def test_device_resident_samples(extractor):
    """Device pointers are used in place (no staging copy)."""
    torch = pytest.importorskip("torch")
    samples = O.fixture_fill(O.FIELD_CAVE, [0, 0, 0])
    want = O.extract_regular(samples, generation=77)
    dev = torch.from_numpy(samples.view(np.int32)).cuda()
    extractor.dispatch(dev, 77, ALL, 0)
    assert_vertices_equal(extractor.vertices_buffer(len(want.vertices)), want.vertices, "device input")
    assert np.array_equal(extractor.indices_buffer(len(want.indices)), want.indices)


# ---- edge 64 (no reference counterpart: pinned by the oracle's generalisation over the edge) ----

@pytest.mark.parametrize("kind,page,lod", [
    (O.FIELD_SPHERE, [0, 0, 0], 0), (O.FIELD_SPHERE, [-1, -1, -1], 0), (O.FIELD_PLANE, [0, -1, 0], 0),
    (O.FIELD_CAVE, [-1, 0, 0], 0), (O.FIELD_SHARP_CORNER, [0, 0, 0], 0), (O.FIELD_MATERIAL_SEAM, [-1, -1, 0], 1),
    (O.FIELD_TERRAIN_FBM, [0, -1, 0], 0), (O.FIELD_TERRAIN_FBM, [3, -1, -7], 0), (O.FIELD_TERRAIN_FBM, [0, -1, 0], 2),
])
@pytest.mark.parametrize("variant", list(VARIANTS))
def test_edge64_matches_oracle(kind, page, lod, variant):
    ex = H.TransvoxelGpuExtractor(0, H.TransvoxelGpuExtractorConfig(262_144, 393_216), edge=64, **VARIANTS[variant])
    samples = O.fixture_fill(kind, page, lod=lod, edge=64)
    for mask, dirty in [(0, ALL), (0x3F, ALL), (0x12, 0x0F0F_0000_FFFF_00F0)]:
        want, _, _ = _check_dispatch(ex, samples, 42, dirty, mask, f"e64 kind {kind} {page} mask {mask:#x}", edge=64)
        if VARIANTS[variant]["debug_records"]:
            cells = ex.cells_buffer()
            visited = want.cell_ranges[:, 0] != 0xFFFFFFFF
            assert np.array_equal(cells["packed_case_class_counts"][visited], want.cell_words[visited, 0])
            check_offsets(ex.offsets_buffer(), ex.blocks_buffer(), want.cell_ranges, 42, "e64")
    ex.close()


@pytest.mark.parametrize("edge", [32, 64])
def test_dense_random_worst_case(edge):
    """Dense adversarial field: ~6 vertices per cell, every emission batch full."""
    cells = edge ** 3
    ex = H.TransvoxelGpuExtractor(0, H.TransvoxelGpuExtractorConfig(cells * 12, cells * 15), edge=edge, debug_records=False)
    samples = O.fixture_fill(O.FIELD_DENSE_RANDOM, [1, 2, 3], edge=edge)
    want, _, _ = _check_dispatch(ex, samples, 9, ALL, 0x3F, f"dense e{edge}", edge=edge)
    assert len(want.vertices) > 4 * cells
    ex.close()


@pytest.mark.parametrize("edge", [32, 64])
def test_batch_of_mixed_chunks(edge):
    """N chunks per dispatch: every chunk's slot equals its single-chunk oracle result."""
    specs = [(O.FIELD_SPHERE, [0, 0, 0]), (O.FIELD_PLANE, [0, 1, 0]), (O.FIELD_PLANE, [0, -1, 0]),
             (O.FIELD_TERRAIN_FBM, [0, -1, 0]), (O.FIELD_CAVE, [-1, -1, -1]), (O.FIELD_PLANE, [5, -3, 2]),
             (O.FIELD_TERRAIN_FBM, [2, -1, 1]), (O.FIELD_SHARP_CORNER, [0, 0, 0]), (O.FIELD_THIN_SLAB, [0, -1, 0]),
             (O.FIELD_MATERIAL_SEAM, [0, -1, 0])] * 40  # 400 chunks > 148 SMs
    n = len(specs)
    batch = H.ChunkBatchExtractor(0, edge=edge, max_chunks=n, max_vertices=40_000, max_indices=60_000)
    samples = np.concatenate([O.fixture_fill(k, p, edge=edge) for k, p in specs[:10]] * 40)
    masks = [(i * 7) % 64 for i in range(n)]
    gens = [100 + i for i in range(n)]
    batch.extract_regular(samples, n, generation=gens, transition_mask=masks)
    counters = batch.counters(n)
    ranges = batch.ranges(n)
    wants = {}
    for i in range(n):
        key = (i % 10, masks[i])
        if key not in wants:
            wants[key] = O.extract_regular(samples[(i % 10) * (edge + 2) ** 3:(i % 10 + 1) * (edge + 2) ** 3], edge=edge,
                                           transition_mask=masks[i], debug=False)
        want = wants[key]
        assert counters["completed"][i] == 1 and counters["required_vertices"][i] == len(want.vertices), i
        assert ranges["first_vertex"][i] == i * 40_000 and ranges["vertex_count"][i] == len(want.vertices)
        assert ranges["first_index"][i] == i * 60_000 and ranges["index_count"][i] == len(want.indices)
    # spot-check full meshes, and the packed readback of the whole batch
    verts, idx, packed = batch.ctx.read_meshes(0, 0, n)
    for i in list(range(0, n, 37)) + [n - 1]:
        want = wants[(i % 10, masks[i])]
        r = packed[i]
        assert_vertices_equal(verts[r["first_vertex"]:r["first_vertex"] + r["vertex_count"]], want.vertices, f"chunk {i}")
        assert np.array_equal(idx[r["first_index"]:r["first_index"] + r["index_count"]], want.indices), i
    assert batch.ctx.launch_count >= 2
    with pytest.raises(H.BatchCapacity):
        batch.extract_regular(np.zeros((n + 1) * (edge + 2) ** 3, dtype=np.uint32), n + 1)
    batch.close()


@pytest.mark.parametrize("edge,n,all_surface", [(64, 592, False), (32, 1184, False), (64, 1184, True), (32, 2368, True)])
def test_large_batch_is_deterministic_and_equal_across_kernel_generations(edge, n, all_surface):
    """Race detector for the asynchronous (decoupled) kernel at a size where every SM walks several chunks.

    A terrain batch (surface and empty chunks mixed, random transition masks, a few partially dirty
    chunks) is extracted three times with the default kernel and once with the first-generation
    kernel (CTA-wide barriers, a second context with HVX_CFG_FIRST_GENERATION): counters, ranges and every mesh byte must be
    identical, and a sample of chunks must equal the oracle.  ``all_surface`` puts every chunk on the
    surface layer: the emission warps are the bottleneck throughout, the work queue stays full and the
    front end is throttled by the slab ring -- the opposite regime of the headline batch.
    """
    rng = np.random.default_rng(7)
    side = int(round(n ** 0.5)) + 1
    pages = np.array([[x - side // 2, -1 if all_surface or (x + z) % 3 else 0, z - side // 2]
                      for z in range(side) for x in range(side)][:n], dtype=np.int64)
    batch = H.ChunkBatchExtractor(0, edge=edge, max_chunks=n, max_vertices=49_152 if edge == 64 else 12_288,
                                  max_indices=73_728 if edge == 64 else 18_432)
    batch.fill_density(O.FIELD_TERRAIN_FBM, pages)
    masks = [int(m) for m in rng.integers(0, 64, n)]
    dirty = [ALL] * n
    for i in range(0, n, 17):
        dirty[i] = int(rng.integers(1, 1 << 62))
    gens = [1000 + i for i in range(n)]

    def run(b=batch):
        b.extract_regular(None, n, generation=gens, transition_mask=masks, dirty_microbricks=dirty)
        c, r = b.counters(n).copy(), b.ranges(n).copy()
        v, i, packed = b.ctx.read_meshes(0, 0, n)
        return c, r, v.copy(), i.copy(), packed.copy()

    first = run()
    assert int(first[0]["vertex_overflow"].sum()) == 0 and int(first[0]["completed"].sum()) == n
    assert int(first[0]["emitted_vertices"].astype(np.int64).sum()) > 100 * n
    for _ in range(2):
        again = run()
        for a, b in zip(first, again):
            assert a.tobytes() == b.tobytes()
    gen1 = H.ChunkBatchExtractor(0, edge=edge, max_chunks=n, max_vertices=batch.ctx.max_vertices,
                                 max_indices=batch.ctx.max_indices, first_generation=True)
    gen1.fill_density(O.FIELD_TERRAIN_FBM, pages)   # the same (deterministic) samples in its own arena
    old = run(gen1)
    gen1.close()
    for a, b in zip(first, old):
        assert a.tobytes() == b.tobytes()
    # a sample of chunks against the oracle (surface chunks included)
    picks = [int(k) for k in np.argsort(-first[0]["emitted_vertices"].astype(np.int64))[:3]] + [0, n - 1]
    for k in picks:
        s = O.fixture_fill(O.FIELD_TERRAIN_FBM, [int(t) for t in pages[k]], edge=edge)
        want = O.extract_regular(s, edge=edge, transition_mask=masks[k], dirty_microbricks=dirty[k], debug=False)
        r = first[4][k]
        assert r["vertex_count"] == len(want.vertices) and r["index_count"] == len(want.indices), k
        assert_vertices_equal(first[2][r["first_vertex"]:r["first_vertex"] + r["vertex_count"]], want.vertices, f"chunk {k}")
        assert np.array_equal(first[3][r["first_index"]:r["first_index"] + r["index_count"]], want.indices), k
    batch.close()


def test_cost_hints_change_the_start_order_and_nothing_else():
    """hvx_chunk_desc.cost_hint: heaviest-first scheduling must leave counters, ranges and every mesh byte alone."""
    rng = np.random.default_rng(3)
    n = 300
    pages = np.array([[x - 10, -1 if (x + z) % 2 else 1, z - 8] for z in range(15) for x in range(20)][:n], dtype=np.int64)
    batch = H.ChunkBatchExtractor(0, edge=64, max_chunks=n, max_vertices=49_152, max_indices=73_728)
    batch.fill_density(O.FIELD_TERRAIN_FBM, pages)
    masks = [int(m) for m in rng.integers(0, 64, n)]

    def run(descs):
        batch.ctx.extract_regular(None, descs, n)
        c, r = batch.counters(n).copy(), batch.ranges(n).copy()
        v, i, packed = batch.ctx.read_meshes(0, 0, n)
        return c, r, v.copy(), i.copy(), packed.copy()

    plain = run(H.make_descs(n, transition_mask=masks))
    hints = [int(v) for v in plain[0]["required_vertices"]]
    assert max(hints) > 0 and min(hints) == 0
    for h in (hints, [int(v) for v in rng.integers(0, 1000, n)], [7] * n):
        again = run(H.make_descs(n, transition_mask=masks, cost_hint=h))
        for a, b in zip(plain, again):
            assert a.tobytes() == b.tobytes()
    batch.close()


def test_edge_parameter_division_is_exact():
    """The decoupled kernel's branch-free d0 / (d0 - d1): all 2^32 pairs of i16 densities, bit for bit against
    the reference formula with IEEE division (hvx_selftest_edge_parameter)."""
    import ctypes as C
    bad, witness = C.c_uint64(123), C.c_uint32()
    rc = H._ffi.load().hvx_selftest_edge_parameter(0, C.byref(bad), C.byref(witness))
    assert rc == 0
    assert bad.value == 0, f"{bad.value} operand pairs differ, e.g. d0={(witness.value >> 16) - 32768} d1={(witness.value & 0xffff) - 32768}"


def test_inverse_square_root_is_exact():
    """The normal's branch-free 1 / sqrt(s): every float from the kernel's guard 1e-12 to 2^40, bit for bit against
    IEEE sqrt followed by IEEE division (hvx_selftest_inv_sqrt)."""
    import ctypes as C
    bad, witness = C.c_uint64(123), C.c_uint32()
    assert H._ffi.load().hvx_selftest_inv_sqrt(0, C.byref(bad), C.byref(witness)) == 0
    assert bad.value == 0, f"{bad.value} operands differ, e.g. s = float bits {witness.value:#x}"


@pytest.mark.parametrize("edge,n", [(64, 520), (32, 3600)])
def test_host_sample_pipeline_and_packed_readback(edge, n):
    """Host samples go up in ~256 MiB sub-batches beside the kernels (hvx_extract_regular), and
    hvx_extract_regular_to_host also brings the packed meshes back behind them: both must equal the
    single-launch device-resident dispatch byte for byte -- with cost hints, partially dirty chunks and
    chunks flagged HVX_CHUNK_UNIFORM (whose samples are never uploaded: the host array holds garbage there)."""
    rng = np.random.default_rng(5)
    side = int(round(n ** 0.5)) + 1
    pages = np.array([[x - side // 2, (-1, 0, 1)[(x + 2 * z) % 3], z - side // 2] for z in range(side) for x in range(side)][:n],
                     dtype=np.int64)
    mv, mi = (49_152, 73_728) if edge == 64 else (12_288, 18_432)
    words = (edge + 2) ** 3
    assert n * words * 4 > 2 * (256 << 20)          # at least three sub-batches
    batch = H.ChunkBatchExtractor(0, edge=edge, max_chunks=n, max_vertices=mv, max_indices=mi)
    batch.fill_density(O.FIELD_TERRAIN_FBM, pages)
    masks = [int(m) for m in rng.integers(0, 64, n)]
    dirty = [ALL] * n
    for i in range(3, n, 29):
        dirty[i] = int(rng.integers(1, 1 << 62))

    def snapshot():
        c, r, k = batch.counters(n).copy(), batch.ranges(n).copy(), batch.classify_counters(n).copy()
        v, i, packed = batch.ctx.read_meshes(0, 0, n)
        return c, r, k, v.copy(), i.copy(), packed.copy()

    batch.ctx.extract_regular(None, H.make_descs(n, 7, dirty, masks), n)
    want = snapshot()
    surface = want[0]["required_vertices"] > 0
    assert 0 < surface.sum() < n
    host = batch.ctx.read(H._ffi.BUF_SAMPLES, 0, n * words)
    hints = [int(v) for v in want[0]["required_vertices"]]
    # chunks without a surface are flagged uniform (every third of them, so runs of both kinds occur) and scribbled over
    flags = np.zeros(n, dtype=np.uint32)
    flags[np.nonzero(~surface)[0][::3]] = H._ffi.HVX_CHUNK_UNIFORM
    for i in np.nonzero(flags)[0]:
        host[i * words:(i + 1) * words] = 0xDEAD0001
    descs = H.make_descs(n, 7, dirty, masks, cost_hint=hints, flags=[int(f) for f in flags])
    # 1. sub-batched upload, outputs stay on the device
    batch.ctx.write(H._ffi.BUF_SAMPLES, np.zeros(words, dtype=np.uint32), 0)   # the arena really is overwritten by the upload
    batch.ctx.extract_regular(host, descs, n)
    got = snapshot()
    for a, b in zip(want, got):
        assert a.tobytes() == b.tobytes()
    # 2. upload + extraction + packed read-back in one call
    v, i, ranges, counters = batch.ctx.extract_regular_to_host(host, descs, n, vertex_cap=len(want[3]) + 7, index_cap=len(want[4]))
    assert v.tobytes() == want[3].tobytes() and i.tobytes() == want[4].tobytes()
    assert ranges.tobytes() == want[5].tobytes() and counters.tobytes() == want[0].tobytes()
    # 3. too small a destination: totals and placement are still reported
    with pytest.raises(H.InvalidExtractionCapacity):
        batch.ctx.extract_regular_to_host(host, descs, n, vertex_cap=len(want[3]) - 1, index_cap=len(want[4]))
    # 4. device-resident input through the same entry point (one launch, then the read-back)
    batch.ctx.extract_regular(host, H.make_descs(n, 7, dirty, masks), n)      # refill the arena (uniform chunks hold the scribble)
    v, i, ranges, counters = batch.ctx.extract_regular_to_host(None, descs, n, vertex_cap=len(want[3]), index_cap=len(want[4]))
    assert v.tobytes() == want[3].tobytes() and i.tobytes() == want[4].tobytes() and counters.tobytes() == want[0].tobytes()
    batch.close()


def test_uniform_flag_on_a_surface_chunk_empties_it():
    """HVX_CHUNK_UNIFORM is the caller's promise: a flagged chunk reports an empty, completed mesh whatever it holds."""
    pages = np.array([[0, -1, 0], [1, -1, 0], [0, 5, 0]], dtype=np.int64)
    batch = H.ChunkBatchExtractor(0, edge=64, max_chunks=3)
    batch.fill_density(O.FIELD_TERRAIN_FBM, pages)
    batch.ctx.extract_regular(None, H.make_descs(3, flags=[0, H._ffi.HVX_CHUNK_UNIFORM, H._ffi.HVX_CHUNK_UNIFORM]), 3)
    c, r, k = batch.counters(3), batch.ranges(3), batch.classify_counters(3)
    assert c["required_vertices"][0] > 0 and list(c["required_vertices"][1:]) == [0, 0] and list(c["completed"]) == [1, 1, 1]
    assert list(r["first_vertex"]) == [0, 49_152, 98_304] and list(r["vertex_count"][1:]) == [0, 0]
    assert list(k["visited_cells"]) == [64 ** 3] * 3 and list(k["active_cells"][1:]) == [0, 0]
    with pytest.raises(H.HvxError, match="unknown descriptor flags"):
        batch.ctx.extract_regular(None, H.make_descs(3, flags=[0, 2, 0]), 3)
    batch.close()


@pytest.mark.parametrize("edge,n", [(32, 1), (32, 7), (32, 60), (64, 1), (64, 5), (64, 40), (64, 147), (64, 175), (32, 500)])
def test_split_walk_equals_the_whole_chunk_walk(edge, n):
    """A dispatch with fewer chunks than resident CTAs walks z-ranges of chunks (counting walk + look-back over
    per-part totals); one with a few waves of chunks and a thin last wave (175 and 500 here: 159 on 148 CTAs, 454 on
    444) can split only the chunks of that wave and walk the others whole, in the same launch (debug bit 0x200).  Counters, classify counters, ranges and every mesh byte must equal the unsplit walk
    (hvx_debug_set_mode 0x100) -- with transition masks, partially dirty and empty-dirty chunks, cost hints, uniform
    flags -- and the oracle."""
    rng = np.random.default_rng(100 * edge + n)
    side = int(np.ceil(n ** 0.5))
    pages = np.array([[x - 1, (-1, -1, 0, -2)[(x + z) % 4] if edge == 64 else (-2, -1, -2, -3)[(x + z) % 4], z - 2]
                      for z in range(side) for x in range(side)][:n], dtype=np.int64)
    mv, mi = (49_152, 73_728) if edge == 64 else (12_288, 18_432)
    batch = H.ChunkBatchExtractor(0, edge=edge, max_chunks=n, max_vertices=mv, max_indices=mi)
    batch.fill_density(O.FIELD_TERRAIN_FBM, pages)
    masks = [int(m) for m in rng.integers(0, 64, n)]
    dirty = [ALL] * n
    for i in range(1, n, 3):
        dirty[i] = int(rng.integers(1, 1 << 62)) if i % 2 else 0xFFFF << (16 * int(rng.integers(0, 4)))
    if n > 4:
        dirty[4] = 0
    hints = [int(h) for h in rng.integers(0, 1000, n)]
    flags = [H._ffi.HVX_CHUNK_UNIFORM if (i % 11 == 10) else 0 for i in range(n)]
    descs = H.make_descs(n, 5, dirty, masks, cost_hint=hints, flags=flags)

    def run():
        batch.ctx.extract_regular(None, descs, n)
        c, r, k = batch.counters(n).copy(), batch.ranges(n).copy(), batch.classify_counters(n).copy()
        v, i, packed = batch.ctx.read_meshes(0, 0, n)
        return c, r, k, v.copy(), i.copy(), packed.copy()

    if n > 150:
        batch.ctx.debug_set_mode(0x200)   # more chunks than resident CTAs: ask for the split last wave
    launches0 = batch.ctx.launch_count
    split = run()
    split_launches = batch.ctx.launch_count - launches0
    batch.ctx.debug_set_mode(0x100)
    launches0 = batch.ctx.launch_count
    whole = run()
    whole_launches = batch.ctx.launch_count - launches0
    batch.ctx.debug_set_mode(0x200 if n > 150 else 0)
    assert split_launches == whole_launches, "the split walk is still one launch"
    for a, b in zip(split, whole):
        assert a.tobytes() == b.tobytes()
    again = run()
    batch.ctx.debug_set_mode(0)
    for a, b in zip(split, again):
        assert a.tobytes() == b.tobytes()
    assert int(split[0]["emitted_vertices"].astype(np.int64).sum()) > 1000
    words = (edge + 2) ** 3
    for k in sorted({0, n // 2, n - 1}):
        if flags[k]:
            continue
        s = batch.ctx.read(H._ffi.BUF_SAMPLES, k * words, words)
        want = O.extract_regular(s, edge=edge, transition_mask=masks[k], dirty_microbricks=dirty[k], generation=5, debug=False)
        r = split[5][k]
        assert r["vertex_count"] == len(want.vertices) and r["index_count"] == len(want.indices), k
        assert_vertices_equal(split[3][r["first_vertex"]:r["first_vertex"] + r["vertex_count"]], want.vertices, f"chunk {k}")
        assert np.array_equal(split[4][r["first_index"]:r["first_index"] + r["index_count"]], want.indices), k
    batch.close()


def test_split_walk_overflow_is_reported_for_the_chunk():
    """Capacity is a property of the chunk, not of a part: one element short anywhere suppresses the whole emission."""
    samples = O.fixture_fill(O.FIELD_PLANE, [0, -1, 0])
    for capacity in [(4095, 6144), (4096, 6143), (2000, 100_000)]:
        ex = H.TransvoxelGpuExtractor(0, H.TransvoxelGpuExtractorConfig.new(*capacity), debug_records=False)
        ex.dispatch(samples, 9, ALL, 0)
        c = ex.counters_buffer()
        assert c["completed"] == 1 and c["required_vertices"] == 4096 and c["required_indices"] == 6144
        assert c["emitted_vertices"] == 0 and c["emitted_indices"] == 0
        assert (c["vertex_overflow"] != 0) == (capacity[0] < 4096) and (c["index_overflow"] != 0) == (capacity[1] < 6144)
        ex.close()
