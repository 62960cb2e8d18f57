"""CPU: the C-ABI library loads and exports every symbol include/hvx.h declares, the host-side LOD
scheduler input matches the Python oracle and the reference's own unit tests
(PV/src/lod_topology.rs:403-619), and the Python mirror validates like the Rust constructors."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import helio_b200 as H
from helio_b200 import _ffi
from oracle import lod_topology as LT

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "hvx.h").read_text()
    declared = set(re.findall(r"\b(hvx_[a-z_0-9]+)\s*\(", header))
    declared -= {"hvx_ctx"}
    lib = _ffi.load()
    missing = [name for name in sorted(declared) if not hasattr(lib, name)]
    assert not missing, missing
    assert set(_ffi.EXPORTS) == declared
    assert lib.hvx_abi_version() == 2
    assert lib.hvx_status_name(-1).decode() == "HVX_E_SAMPLE_COUNT" and lib.hvx_status_name(0).decode() == "HVX_OK"


def test_struct_layouts_match_the_reference_pods():
    """PV/tests/wgsl_layout.rs style: sizes / offsets of every POD crossing the boundary."""
    assert H.VERTEX_DTYPE.itemsize == 32 and H.VERTEX_DTYPE.fields["normal"][1] == 16 and H.VERTEX_DTYPE.fields["flags"][1] == 28
    assert H.EMISSION_COUNTERS_DTYPE.itemsize == 32 and H.EMISSION_COUNTERS_DTYPE.fields["completed"][1] == 24
    assert H.CLASSIFY_COUNTERS_DTYPE.itemsize == 16
    assert H.TRANSITION_COUNTERS_DTYPE.itemsize == 48 and H.TRANSITION_COUNTERS_DTYPE.fields["completed"][1] == 32
    assert H.CELL_RECORD_DTYPE.itemsize == H.CELL_OFFSET_DTYPE.itemsize == H.SCAN_BLOCK_DTYPE.itemsize == H.RANGE_DTYPE.itemsize == 16
    assert C.sizeof(_ffi.ChunkDesc) == 32 and C.sizeof(_ffi.Config) == 32 and C.sizeof(_ffi.Page) == 32


def test_no_gpu_means_a_loud_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(H.CudaError, match="no CPU fallback"):
        H.TransvoxelGpuExtractor(0)
    with pytest.raises(H.CudaError):
        H.ChunkBatchExtractor(0, edge=64, max_chunks=4)


def test_config_validation_happens_before_any_device_work():
    """PV/src/transvoxel_emit.rs:63-75, PV/src/transvoxel_transition_gpu.rs:160-171."""
    with pytest.raises(H.InvalidExtractionCapacity) as info:
        H.TransvoxelGpuExtractorConfig.new(0, 7)
    assert info.value.max_vertices == 0 and info.value.max_indices == 7 and isinstance(info.value, H.TransvoxelGpuError)
    with pytest.raises(H.TransitionInvalidExtractionCapacity):
        H.TransvoxelGpuTransitionExtractorConfig.new(3, 0)
    assert H.TransvoxelGpuExtractorConfig() == H.TransvoxelGpuExtractorConfig(393_216, 491_520)
    assert H.TransvoxelGpuTransitionExtractorConfig() == H.TransvoxelGpuTransitionExtractorConfig(73_728, 221_184)
    with pytest.raises(H.InvalidExtractionCapacity):   # the C ABI applies the same rule
        H.Context(0, max_vertices=0, max_indices=1)
    with pytest.raises(H.HvxError):
        H.Context(0, edge=48)


def test_value_types():
    word = H.CellWord(-123, 17, 9)
    assert int(word) == 0x0911FF85 and word.density() == -123 and word.material() == 17 and word.flags() == 9
    assert word.is_solid() and not H.CellWord(raw=H.CellWord.AIR).is_solid()
    mask = 0
    for i, face in enumerate(H.TransitionFace):
        assert face.index() == i and face.axis() == i // 2 and face.is_positive() == bool(i & 1)
        mask |= face.bit()
    assert mask == H.TRANSITION_FACE_MASK
    assert H.PageKey(3, [5, -3, 0]).parent() == H.PageKey(4, [2, -2, 0])
    assert H.PageKey(1, [-3, 2, 0]).lod0_cell_min() == (-192, 128, 0)
    assert sorted([H.PageKey(1, [0, 0, 0]), H.PageKey(0, [9, 9, 9])])[0].lod == 0
    cell = H.GpuTransvoxelTransitionCell(0x1FF | (0xAB << 9) | (12 << 17) | (12 << 21) | 0x80000000, 7)
    assert cell.case_index() == 0x1FF and cell.class_index() == 0x2B and cell.reverse_winding()
    assert cell.vertex_count() == 12 and cell.triangle_count() == 12 and cell.is_valid_for(7) and not cell.is_valid_for(8)


# ---- LOD topology: the reference's unit tests, restated (PV/src/lod_topology.rs:407-538) -------------

def test_coarse_page_owns_transition_bits_on_every_face_and_quadrant():
    for face in H.TransitionFace:
        axis = face.axis()
        tangential = [a for a in range(3) if a != axis]
        for quadrant in range(4):
            fine = [-6, -6, -6]
            fine[axis] += 2 if face.is_positive() else -1
            fine[tangential[0]] += quadrant & 1
            fine[tangential[1]] += (quadrant >> 1) & 1
            coarse, fine_key = H.PageKey(1, [-3, -3, -3]), H.PageKey(0, fine)
            topo = H.TerrainLodTopology([coarse, fine_key])
            assert topo.transition_mask(coarse) == face.bit() and topo.transition_mask(fine_key) == 0


def test_edge_and_corner_contacts_do_not_create_false_transition_faces():
    coarse = H.PageKey(1, [-1, -1, -1])
    topo = H.TerrainLodTopology([coarse, H.PageKey(0, [0, 0, -2]), H.PageKey(0, [0, 0, 0])])
    assert topo.transition_mask(coarse) == 0 and topo.stats().transition_faces == 0


def test_topology_errors_are_explicit():
    with pytest.raises(H.TerrainLodTopologyError) as info:
        H.TerrainLodTopology([H.PageKey(2, [-1, -1, -1]), H.PageKey(0, [-4, -4, -4])])
    assert info.value.kind == "OverlappingPages"
    with pytest.raises(H.TerrainLodTopologyError) as info:
        H.TerrainLodTopology([H.PageKey(2, [-1, -1, -1]), H.PageKey(0, [0, -4, -4])])
    assert info.value.kind == "UnbalancedFace"
    with pytest.raises(H.TerrainLodTopologyError) as info:
        H.TerrainLodTopology([])
    assert info.value.kind == "Empty"
    with pytest.raises(H.TerrainLodTopologyError) as info:
        H.TerrainLodTopology([H.PageKey(0, [1, 1, 1]), H.PageKey(0, [1, 1, 1])])
    assert info.value.kind == "DuplicatePage"
    with pytest.raises(H.AddressError):
        H.TerrainLodTopology([H.PageKey(58, [0, 0, 0])])
    with pytest.raises(H.TerrainLodTopologyError) as info:
        H.HorizonLodFixturePlan.build([63_710_000, -1, 0], 10, 32)
    assert info.value.kind == "PageBudget"
    with pytest.raises(H.TerrainLodTopologyError) as info:
        H.HorizonLodFixturePlan.build([0, 0, 0], 0, 32)
    assert info.value.kind == "UnsupportedRootLod"
    with pytest.raises(H.TerrainLodTopologyError) as info:
        H.HorizonLodFixturePlan.build_with_minimum_lod([0, 0, 0], 4, 4, 32)
    assert info.value.kind == "UnsupportedMinimumLod"


def test_horizon_plan_is_deterministic_bounded_balanced_and_exact():
    focus = [63_710_000, -1, -17]
    first, second = H.HorizonLodFixturePlan.build(focus, 11, 96), H.HorizonLodFixturePlan.build(focus, 11, 96)
    assert first == second
    stats = first.topology().stats()
    assert stats.minimum_lod == 0 and stats.maximum_lod >= 7 and stats.pages <= 96 and stats.transition_faces > 0
    root, masks, ostats = LT.horizon_plan(focus, 11, 0, 96)
    assert (first.root().lod, first.root().page_xyz) == root
    assert {(k.lod, k.page_xyz): v for k, v in first.topology().transition_masks().items()} == masks
    assert (stats.pages, stats.minimum_lod, stats.maximum_lod, stats.transition_faces) == tuple(ostats.values())


def test_horizon_plan_crosses_signed_boundaries_and_coarsens_with_altitude():
    plans = [H.HorizonLodFixturePlan.build(f, 11, 96) for f in
             ([63_710_000, -1, -1], [63_710_032, -1, 0], [-63_710_001, -1, -33], [-63_710_033, -1, 32])]
    assert all(p.topology().stats().pages <= 96 for p in plans)
    assert plans[0].topology().pages() != plans[2].topology().pages()
    ground = H.HorizonLodFixturePlan.build_with_minimum_lod([63_710_000, -1, 17], 11, 0, 96)
    orbit = H.HorizonLodFixturePlan.build_with_minimum_lod([63_710_000, -1, 17], 11, 5, 96)
    assert ground.topology().stats().minimum_lod == 0 and orbit.topology().stats().minimum_lod == 5
    assert orbit.topology().stats().pages < ground.topology().stats().pages


def test_randomized_signed_horizon_neighbourhoods_match_the_oracle():
    """PV/src/lod_topology.rs:508-533 with its xorshift64 seed; 48 of the 512 cases, each checked
    page-for-page and mask-for-mask against the Python restatement, at both page edges."""
    state = 0x4D595DF4D0F33173

    def nxt():
        nonlocal state
        state ^= (state << 13) & 0xFFFFFFFFFFFFFFFF
        state ^= state >> 7
        state ^= (state << 17) & 0xFFFFFFFFFFFFFFFF
        return state

    def coordinate():
        magnitude = nxt() % 127_420_000
        return magnitude if nxt() & 1 == 0 else -magnitude
    for case in range(48):
        focus = [coordinate(), -1, coordinate()]
        minimum = nxt() % 6
        edge = 32 if case % 3 else 64
        plan = H.HorizonLodFixturePlan.build_with_minimum_lod(focus, 11, minimum, 192, edge=edge)
        root, masks, stats = LT.horizon_plan(focus, 11, minimum, 192, edge=edge)
        assert (plan.root().lod, plan.root().page_xyz) == root, case
        assert {(k.lod, k.page_xyz): v for k, v in plan.topology().transition_masks().items()} == masks, case
        assert plan.topology().stats().minimum_lod == minimum and plan.topology().stats().pages <= 192
        # re-deriving the masks from the bare page list gives the same answer (single ownership)
        again = H.TerrainLodTopology(plan.topology().pages(), edge=edge)
        assert again == plan.topology()


def test_partition_is_deterministic_balanced_and_lod_aware():
    costs = np.array([H.chunk_cost(64, m) for m in [0, 0x3F, 0, 1, 0, 0x15, 0, 0, 3, 0, 0, 0x3F]], dtype=np.uint64)
    owner = H.partition_chunks(costs, 4)
    assert np.array_equal(owner, H.partition_chunks(costs, 4)) and set(owner) == {0, 1, 2, 3}
    loads = [costs[owner == r].sum() for r in range(4)]
    assert max(loads) - min(loads) <= costs.max()
    assert H.chunk_cost(64, 0) == 4 * 66 ** 3 and H.chunk_cost(32, 0x3F) == 4 * 34 ** 3 + 6 * 12 * 67 * 67
    uniform = H.partition_chunks(np.full(4096, 7, dtype=np.uint64), 8)
    assert np.array_equal(np.bincount(uniform), np.full(8, 512)) and np.array_equal(uniform[:16], np.arange(16) % 8)


def test_packed_vertex_table_is_the_row_table_without_its_unused_tails():
    """The regular kernel reads edge codes from a packed copy (shared memory is scarce): same data."""
    import re
    from pathlib import Path
    text = (Path(__file__).resolve().parent.parent / "helio_b200" / "csrc" / "transvoxel_tables.inc").read_text()

    def table(name):
        body = re.search(name + r"(?:\[\d+\])+\s*=\s*\{(.*?)\};", text, re.S).group(1)
        return [int(x, 0) for x in re.findall(r"0x[0-9a-fA-F]+|\d+", body)]

    info, rows = table("HVX_REGULAR_CASE_INFO"), table("HVX_REGULAR_VERTEX_EDGE")
    base, packed = table("HVX_REGULAR_VERTEX_BASE"), table("HVX_REGULAR_VERTEX_PACKED")
    assert len(info) == 256 and len(rows) == 256 * 12 and len(base) == 256 and len(packed) == 1536
    at = 0
    for case in range(256):
        nv = info[case] & 15
        assert base[case] == at
        assert packed[at:at + nv] == rows[case * 12:case * 12 + nv]
        for code in packed[at:at + nv]:       # an edge joins corner c0 to c0 + one bit (what the vertex code relies on)
            c0, c1 = code >> 4, code & 15
            assert c1 > c0 and (c0 ^ c1) in (1, 2, 4) and c1 < 8
        at += nv
    assert at == 1536


def test_planet_scale_page_set_is_reproducible_and_partitions_evenly():
    """BASELINE config 4 (tools/bench_planet.py): the horizon plans drawn like the reference's randomized test
    (PV/src/lod_topology.rs:507-520) give the same page list every time, only coarse pages own transition faces,
    and the LPT partition hands every rank the same bytes to within a page."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
    from bench_planet import next_random, planet_page_set

    x = 0x4D595DF4D0F33173                       # one xorshift64 step (13, 7, 17), spelled out
    x ^= (x << 13) & (2**64 - 1)
    x ^= x >> 7
    x ^= (x << 17) & (2**64 - 1)
    assert next_random(0x4D595DF4D0F33173) == x == 0x76e0cf04b412ebd1
    pages, lods, masks = planet_page_set(12)
    again = planet_page_set(12)
    assert np.array_equal(pages, again[0]) and np.array_equal(lods, again[1]) and np.array_equal(masks, again[2])
    assert 12 * 20 < len(pages) <= 12 * 192
    assert np.all(masks[lods == 0] == 0) and np.any(masks != 0) and np.all(masks < 64)
    costs = np.array([H.chunk_cost(32, int(m)) for m in masks], dtype=np.uint64)
    for ranks in (2, 4, 8):
        owner = H.partition_chunks(costs, ranks)
        load = np.array([costs[owner == r].sum() for r in range(ranks)], dtype=np.float64)
        assert set(owner) == set(range(ranks))
        assert load.max() - load.min() <= float(costs.max()), (ranks, load)


def test_start_order_of_a_hinted_batch():
    """hvx_start_order (the schedule hvx_extract_regular derives from hvx_chunk_desc.cost_hint): a permutation; descending
    hints when the batch is uniform or the spread is off; a few heavy chunks among many light ones are spread evenly
    over the first spread_pct per cent, heaviest first, and the order ends on light chunks."""
    rng = np.random.default_rng(3)
    # the headline's shape: 298 heavy of 4096
    hints = np.zeros(4096, dtype=np.uint32)
    heavy_ids = rng.choice(4096, 298, replace=False)
    hints[heavy_ids] = rng.integers(15_000, 30_000, 298)
    plain = H.start_order(hints, 0)
    assert sorted(plain.tolist()) == list(range(4096))
    assert np.all(np.diff(hints[plain].astype(np.int64)) <= 0), "descending"
    assert np.array_equal(plain[:298], sorted(heavy_ids, key=lambda i: (-int(hints[i]), i)))   # ties in index order
    spread = H.start_order(hints, 75)
    assert sorted(spread.tolist()) == list(range(4096))
    pos = np.flatnonzero(hints[spread] > 0)
    assert len(pos) == 298 and pos[0] == 0 and pos[-1] < 0.75 * 4096
    assert np.all(np.diff(hints[spread][pos].astype(np.int64)) <= 0), "heavy chunks still heaviest first"
    gaps = np.diff(pos)
    assert gaps.min() >= 10 and gaps.max() <= 11, "evenly spread (3072 / 298 = 10.3)"
    assert np.all(hints[spread][int(0.75 * 4096):] == 0), "the order ends on light chunks"
    # an all-surface batch has no heavy chunk (nothing above four times the median): plain descending order
    uniform = rng.integers(20_000, 26_000, 500).astype(np.uint32)
    assert np.array_equal(H.start_order(uniform, 75), H.start_order(uniform, 0))
    # more than a quarter heavy: not "a few"
    many = np.where(np.arange(100) % 3 == 0, 5000, 1).astype(np.uint32)
    assert np.array_equal(H.start_order(many, 75), H.start_order(many, 0))
    assert len(H.start_order(np.zeros(0, dtype=np.uint32))) == 0
    with pytest.raises(H.HvxError):
        H.start_order(hints, 101)
