"""Shared helpers for the parity tests."""
import numpy as np

ALL = (1 << 64) - 1

# fixture -> (ExtractionFixtureKind, page) used by every reference test
# (PV/tests/gpu_transvoxel_emission.rs:490-500)
FIXTURE_PAGES = {
    "plane": (0, [0, -1, 0]), "sphere": (1, [0, 0, 0]), "cave": (2, [0, 0, 0]), "sharp_corner": (3, [0, 0, 0]),
    "thin_slab": (4, [0, -1, 0]), "material_seam": (5, [0, -1, 0]),
}


def ulp_distance(a, b):
    """Element-wise distance in units in the last place between two float32 arrays (+0 == -0)."""
    ai = np.ascontiguousarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    bi = np.ascontiguousarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, np.int64(-(1 << 31)) - ai, ai)
    bi = np.where(bi < 0, np.int64(-(1 << 31)) - bi, bi)
    return np.abs(ai - bi)


def assert_vertices_equal(got, want, label=""):
    """Bit-exact vertex parity (the bar: <= 1 ULP, tolerance 1e-6 relative; we require 0 ULP and
    report the worst ULP distance if that ever fails)."""
    assert got.shape == want.shape, f"{label}: vertex count {got.shape} != {want.shape}"
    assert np.array_equal(got["material"], want["material"]), f"{label}: material"
    assert np.array_equal(got["flags"], want["flags"]), f"{label}: flags"
    if got.tobytes() == want.tobytes():
        return
    dp = ulp_distance(got["position"], want["position"])
    dn = ulp_distance(got["normal"], want["normal"])
    bad = int(np.argmax(dp.max(axis=1) + dn.max(axis=1)))
    raise AssertionError(
        f"{label}: vertices differ: max ULP position {dp.max()} normal {dn.max()}; first worst #{bad}: "
        f"got {got[bad]} want {want[bad]}")


def check_offsets(offsets, blocks, want_ranges, generation, label=""):
    """PV/tests/gpu_transvoxel_emission.rs:103-123: block.first + offset.first == expected range."""
    visited = want_ranges[:, 0] != 0xFFFFFFFF
    gen = offsets["generation_low"].astype(np.uint64) | (offsets["generation_high"].astype(np.uint64) << np.uint64(32))
    assert np.all(gen[visited] == generation), f"{label}: offset generation"
    block_of = np.arange(len(offsets)) // 256
    fv = blocks["first_vertex"][block_of] + offsets["first_vertex"]
    fi = blocks["first_index"][block_of] + offsets["first_index"]
    assert np.array_equal(fv[visited], want_ranges[visited, 0]), f"{label}: vertex ranges"
    assert np.array_equal(fi[visited], want_ranges[visited, 1]), f"{label}: index ranges"
    return gen
