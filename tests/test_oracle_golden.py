"""CPU: the oracle (oracle/hvx_oracle.c) against every known answer the reference pins for this path
(tests/golden/known_answers.json) and against the reference's own CPU unit tests, restated.

No GPU, no reference code: the reference is Rust + wgpu and cannot run in this image.
"""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle as O
from hvx_testutil import ALL, FIXTURE_PAGES

GOLDEN = json.loads((Path(__file__).parent / "golden" / "known_answers.json").read_text())
WEIGHTS = GOLDEN["transition_case_weights"]["value"]


def test_table_audit_fingerprint_and_totals():
    """PV/src/transvoxel.rs:420-435 every_official_case_is_index_safe_and_deterministic."""
    audit, want = O.validate_tables(), GOLDEN["table_audit"]
    assert audit.fingerprint == int(want["fingerprint"], 16)
    for key in ("regular_cases", "transition_cases", "regular_vertices", "regular_triangles", "transition_vertices",
                "transition_triangles", "max_regular_vertices", "max_regular_triangles", "max_transition_vertices",
                "max_transition_triangles"):
        assert getattr(audit, key) == want[key], key
    again = O.validate_tables()
    assert again.fingerprint == audit.fingerprint
    assert O.lib().hvxo_table_revision().decode() == GOLDEN["table_revision"]["value"]


def test_cellword_layout():
    """helio-planet-voxel-core/src/types.rs:395-404."""
    g = GOLDEN["cellword"]
    assert O.cellword(*g["args"]) == int(g["value"], 16)
    assert O.cellword(32767, 0, 0) == int(g["air"], 16)
    word = O.cellword(-123, 17, 9)
    assert np.array([word & 0xFFFF], dtype=np.uint16).view(np.int16)[0] == -123 and (word >> 16) & 0xFF == 17 and word >> 24 == 9


def test_transition_inverse_flag_reverses_winding_and_case_weights():
    """PV/src/transvoxel.rs:437-484."""
    flipped = next(c for c in range(512) if O.case_topology(1, c)["reverse"] and O.case_topology(1, c)["triangle_count"])
    topo = O.case_topology(1, flipped)
    # the oracle returns triangles with the inverse flip applied; undo it through a non-inverse twin
    twin = next(c for c in range(512) if not O.case_topology(1, c)["reverse"]
                and O.case_topology(1, c)["class_index"] == topo["class_index"])
    a, b = topo["triangles"][:3], O.case_topology(1, twin)["triangles"][:3]
    assert a == [b[0], b[2], b[1]]
    for sample, weight in enumerate(WEIGHTS):
        slab = np.full(6 * 3 * 67 * 67, O.cellword(1, 0, 0), dtype=np.uint32)
        su, sv = sample % 3 + 1, sample // 3 + 1
        slab[su + sv * 67 + 67 * 67] = O.cellword(-1, 1, 0)
        mesh = O.extract_transition(slab, 1)
        assert mesh.cell_words[0, 0] & 0x1FF == weight


def test_regular_winding_matches_negative_solid_density_gradient():
    """PV/src/transvoxel.rs:552-573: case 0x55 (x=0 solid, x=1 air) must face +X."""
    topo = O.case_topology(0, 0x55)
    corners = [[i & 1, (i >> 1) & 1, (i >> 2) & 1] for i in range(8)]
    pos = [(np.array(corners[(c & 0xFF) >> 4], float) + np.array(corners[c & 0x0F], float)) * 0.5 for c in topo["codes"]]
    for t in range(topo["triangle_count"]):
        a, b, c = (pos[i] for i in topo["triangles"][3 * t:3 * t + 3])
        assert np.cross(b - a, c - a)[0] > 0


def test_transition_boundary_vertices_match_both_lod_face_contours():
    """PV/src/transvoxel.rs:486-542."""
    full_edges = [(0, 1), (1, 2), (3, 4), (4, 5), (6, 7), (7, 8), (0, 3), (3, 6), (1, 4), (4, 7), (2, 5), (5, 8)]
    half_edges = [(9, 10), (10, 12), (11, 12), (9, 11)]
    dup = [0, 2, 6, 8]

    def solid(case, corner):
        return bool(case & WEIGHTS[corner if corner < 9 else dup[corner - 9]])
    for case in range(512):
        topo = O.case_topology(1, case)
        actual = {tuple(sorted(((c & 0xFF) >> 4, c & 0x0F))) for c in topo["codes"]}
        assert actual & set(full_edges) == {e for e in full_edges if solid(case, e[0]) != solid(case, e[1])}, case
        assert actual & set(half_edges) == {e for e in half_edges if solid(case, e[0]) != solid(case, e[1])}, case


@pytest.mark.parametrize("name", list(FIXTURE_PAGES))
def test_published_fixture_counts(name):
    """docs/planetary_voxel_extraction_benchmark.md:59-70 + PV/src/transvoxel.rs:587-615."""
    want = GOLDEN["fixtures"][name]
    samples = O.fixture_fill(want["kind"], want["page"])
    mesh = O.extract_regular(samples)
    assert len(mesh.vertices) == want["vertices"] and len(mesh.indices) == 3 * want["triangles"]
    assert 32 * len(mesh.vertices) + 4 * len(mesh.indices) == want["indexed_bytes"]
    metrics = O.fixture_metrics(samples)
    assert metrics.active_cells == mesh.classify[1] > 0 and metrics.active_microbrick_mask != 0
    assert metrics.solid_samples + metrics.air_samples == 34 ** 3
    assert mesh.indices.max() < len(mesh.vertices)
    again = O.extract_regular(O.fixture_fill(want["kind"], want["page"]))
    assert again.vertices.tobytes() == mesh.vertices.tobytes() and np.array_equal(again.indices, mesh.indices)


def test_first_sphere_vertex():
    mesh = O.extract_regular(O.fixture_fill(O.FIELD_SPHERE, [0, 0, 0]))
    assert tuple(mesh.vertices["position"][0]) == (12.0, 0.0, 0.0) and tuple(mesh.vertices["normal"][0]) == (1.0, 0.0, 0.0)
    assert mesh.vertices["material"][0] == 1


def test_dirty_secondary_and_overflow_contracts():
    samples = O.fixture_fill(O.FIELD_PLANE, [0, -1, 0])
    dirty = O.extract_regular(samples, dirty_microbricks=GOLDEN["dirty_microbrick"]["dirty"])
    assert dirty.classify[0] == GOLDEN["dirty_microbrick"]["visited_cells"] and dirty.classify[1] > 0
    visited = dirty.cell_ranges[:, 0] != 0xFFFFFFFF
    lin = np.flatnonzero(visited)
    assert np.all(((lin % 32) // 8) + ((lin // 32) % 32 // 8) * 4 + (lin // 1024 // 8) * 16 == 12)
    g = GOLDEN["secondary_position"]
    sec = O.extract_regular(samples, transition_mask=g["mask"])
    x, z = sec.vertices["position"][:, 0], sec.vertices["position"][:, 2]
    inside = (z >= g["z_range"][0]) & (z < g["z_range"][1])
    assert np.any((np.abs(x - g["x_present"]) <= 1e-5) & inside) and not np.any((np.abs(x - g["x_absent"]) <= 1e-5) & inside)
    tiny = O.extract_regular(samples, max_vertices=1, max_indices=1)
    assert list(tiny.counters) == [4096, 6144, 0, 0, 1, 1, 1, 0]


def test_adjacent_page_halos_and_material_seam():
    """PV/src/fixture.rs:255-310."""
    for kind in range(6):
        left = O.fixture_fill(kind, [-1, -1, -1]).reshape(34, 34, 34)
        right = O.fixture_fill(kind, [0, -1, -1]).reshape(34, 34, 34)
        assert np.array_equal(left[:, :, 33], right[:, :, 1]) and np.array_equal(left[:, :, 32], right[:, :, 0])
    seam = O.fixture_fill(O.FIELD_MATERIAL_SEAM, [-1, -1, -1]).reshape(34, 34, 34)
    plane = O.fixture_fill(O.FIELD_PLANE, [-1, -1, -1]).reshape(34, 34, 34)
    assert np.array_equal(seam & 0xFFFF, plane & 0xFFFF)
    assert (seam[32, 32, 32] >> 16) & 0xFF == 1 and (seam[32, 32, 33] >> 16) & 0xFF == 2
    plane0 = O.extract_regular(O.fixture_fill(O.FIELD_PLANE, [0, -1, 0]))
    assert plane0.cell_words[31 + 31 * 32 + 31 * 1024, 0] & 0xFF not in (0, 255)  # +face cell uses the halo


@pytest.mark.parametrize("case", range(3))
def test_transition_slab_route_equals_the_reference_analytic_route(case):
    """The GPU path reads gradients from the slab halo; the reference CPU code re-samples the
    analytic field (PV/src/transvoxel_transition.rs:412-424).  Both must agree bit for bit."""
    g = GOLDEN["transition_cases"]["cases"][case]
    slabs = O.slab_fill(g["kind"], g["page"], g["lod"])
    mesh = O.extract_transition(slabs, g["mask"])
    parts_v, parts_i, base = [], [], 0
    for face in range(6):
        if (g["mask"] >> face) & 1:
            v, i = O.extract_transition_face_analytic(g["kind"], g["page"], g["lod"], face)
            parts_v.append(v)
            parts_i.append(i + base)
            base += len(v)
    assert np.concatenate(parts_v).tobytes() == mesh.vertices.tobytes()
    assert np.array_equal(np.concatenate(parts_i), mesh.indices)
    assert mesh.counters[1] == bin(g["mask"]).count("1") and mesh.counters[8] == 1
    assert np.all(np.isin(mesh.vertices["flags"], [1 << f for f in range(6) if (g["mask"] >> f) & 1]))
    assert len(mesh.vertices) <= 6 * 32 * 32 * 12 and mesh.indices.max() < len(mesh.vertices)
    # slab layout: logical face grid at 1..=65 on layer 1 (PV/src/transvoxel_transition.rs:637-649)
    assert slabs.size == 6 * 3 * 67 * 67 == GOLDEN["buffer_budgets"]["slab_bytes"] // 4


def test_every_transition_case_is_finite_index_safe_and_winds_with_the_gradient():
    """PV/src/transvoxel_transition.rs:599-623, 810-832."""
    air, solid = O.cellword(1, 0, 0), O.cellword(-1, 7, 0)
    for case in range(512):
        slab = np.full(6 * 3 * 67 * 67, air, dtype=np.uint32)
        for s in range(9):
            if case & WEIGHTS[s]:
                slab[5 * 3 * 67 * 67 + (s % 3 + 1) + (s // 3 + 1) * 67 + 67 * 67] = solid
        mesh = O.extract_transition(slab, 1 << 5)
        topo = O.case_topology(1, case)
        first = mesh.cell_ranges[5 * 1024 + 1, 0] if case else 0
        assert mesh.cell_words[5 * 1024, 0] & 0x1FF == case
        assert first == topo["vertex_count"] or case == 0
        assert np.isfinite(mesh.vertices["position"]).all() and np.isfinite(mesh.vertices["normal"]).all()
    for face in range(6):  # planar field: solid for u == 0, air beyond -> triangles face +u
        basis_u = [(0, 1, 0), (0, 1, 0), (0, 0, 1), (0, 0, 1), (1, 0, 0), (1, 0, 0)][face]
        slab = np.full(6 * 3 * 67 * 67, air, dtype=np.uint32)
        base = face * 3 * 67 * 67
        cu, cv = 7, 11
        for layer in range(3):
            for dv in range(-1, 4):
                for du in range(-1, 4):
                    u = 2 * cu + du
                    slab[base + (u + 1) + (2 * cv + dv + 1) * 67 + layer * 67 * 67] = solid if u <= 2 * cu else air
        mesh = O.extract_transition(slab, 1 << face)
        lo, hi = mesh.cell_ranges[face * 1024 + cu + cv * 32], mesh.cell_ranges[face * 1024 + cu + cv * 32 + 1]
        tris = mesh.indices[lo[1]:hi[1]].reshape(-1, 3)
        assert len(tris) > 0
        for a, b, c in mesh.vertices["position"][tris]:
            assert np.dot(np.cross(b - a, c - a), basis_u) > 0, face


def test_neighbouring_transition_cells_share_identical_seam_vertices():
    """PV/src/transvoxel_transition.rs:671-698 (quantised 1e-5 like the reference)."""
    mesh = O.extract_transition(O.slab_fill(O.FIELD_PLANE, [0, -1, 0], 1), 1 << 5)
    ranges = mesh.cell_ranges[5 * 1024:6 * 1024]
    nonempty = 0
    for u in range(31):
        def keys(cell, boundary):
            lo = ranges[cell, 0]
            hi = ranges[cell + 1, 0] if cell + 1 < 1024 else len(mesh.vertices)
            v = mesh.vertices[lo:hi]
            on = np.abs(v["position"][:, 0] - boundary) <= 1e-6
            return {tuple(np.round(p * 1e5).astype(int)) + tuple(np.round(n * 1e5).astype(int)) for p, n in
                    zip(v["position"][on], v["normal"][on])}
        left, right = keys(u + 31 * 32, u + 1.0), keys(u + 1 + 31 * 32, u + 1.0)
        assert left == right
        nonempty += bool(left)
    assert nonempty == 31
    # half-resolution side sits a quarter cell inside the page (PV/src/transvoxel_transition.rs:566-582)
    z = mesh.vertices["position"][:, 2]
    assert z.max() == 32.0 and z.min() >= GOLDEN["face_half_width"]["positive"] - 1e-6


def test_edge64_generalisation_is_consistent_with_edge32():
    """A 64^3 chunk is exactly eight 32^3 pages: the union of the octants' cell classifications must
    equal the big chunk's (this is the only pin the edge-64 oracle has besides its code path)."""
    for kind, page in [(O.FIELD_SPHERE, [0, 0, 0]), (O.FIELD_TERRAIN_FBM, [0, -1, 0]), (O.FIELD_CAVE, [-1, -1, -1])]:
        big = O.extract_regular(O.fixture_fill(kind, page, edge=64), edge=64)
        cases64 = (big.cell_words[:, 0] & 0xFF).reshape(64, 64, 64)
        total_v = 0
        for oz in range(2):
            for oy in range(2):
                for ox in range(2):
                    sub = [2 * page[0] + ox, 2 * page[1] + oy, 2 * page[2] + oz]
                    small = O.extract_regular(O.fixture_fill(kind, sub, edge=32), edge=32)
                    total_v += len(small.vertices)
                    cases32 = (small.cell_words[:, 0] & 0xFF).reshape(32, 32, 32)
                    assert np.array_equal(cases64[32 * oz:32 * oz + 32, 32 * oy:32 * oy + 32, 32 * ox:32 * ox + 32], cases32)
        assert total_v == len(big.vertices)


def test_buffer_budgets_match_the_reference():
    b = GOLDEN["buffer_budgets"]
    assert 34 ** 3 * 4 == b["sample_bytes"] and 32 ** 3 * 16 == b["cell_bytes"] == b["offset_bytes"]
    assert 128 * 16 == b["block_bytes"] and 393_216 * 32 == b["vertex_bytes"] and 491_520 * 4 == b["index_bytes"]
    assert 6144 * 16 == b["transition_cell_bytes"] and 73_728 * 32 == b["transition_vertex_bytes"]
    assert 221_184 * 4 == b["transition_index_bytes"]


def test_terrain_fbm_field_basics():
    """noise.rs terrain_sdf: a pure function of position; heightfield => sdf(x, y+d, z) - sdf(x, y, z) == d."""
    a = O.terrain_sdf(1.5, 0.25, -3.0)
    assert a == O.terrain_sdf(1.5, 0.25, -3.0)
    assert np.isclose(O.terrain_sdf(1.5, 1.25, -3.0) - a, 1.0, atol=1e-6)
    heights = [0.25 - O.terrain_sdf(x * 0.7, 0.25, x * -0.3) for x in range(200)]
    assert -6.0001 <= min(heights) and max(heights) <= 2.0001 and max(heights) - min(heights) > 1.0
