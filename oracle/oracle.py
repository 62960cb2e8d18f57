"""ctypes binding of the CPU oracle (oracle/hvx_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, bench.py's cpu_baseline /
``--impl reference`` legs and ``__graft_entry__.smoke()``.  The product package
``helio_b200`` never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None

FIELD_PLANE, FIELD_SPHERE, FIELD_CAVE, FIELD_SHARP_CORNER, FIELD_THIN_SLAB, FIELD_MATERIAL_SEAM = range(6)
FIELD_TERRAIN_FBM = 16
FIELD_DENSE_RANDOM = 17
FIELD_NAMES = {
    "plane": FIELD_PLANE, "sphere": FIELD_SPHERE, "cave": FIELD_CAVE, "sharp_corner": FIELD_SHARP_CORNER,
    "thin_slab": FIELD_THIN_SLAB, "material_seam": FIELD_MATERIAL_SEAM,
    "terrain_fbm": FIELD_TERRAIN_FBM, "dense_random": FIELD_DENSE_RANDOM,
}

MESHLET_DTYPE = np.dtype([(n, "<u4") for n in ("first_index", "index_count", "first_vertex", "vertex_count",
                                                 "bounds_offset", "generation_low", "generation_high", "_pad")])
MESHLET_BOUNDS_DTYPE = np.dtype([("center", "<f4", 3), ("radius", "<f4"), ("cone_apex", "<f4", 3), ("cone_cutoff", "<f4"),
                                 ("cone_axis", "<f4", 3), ("_pad", "<f4")])
VERTEX_DTYPE = np.dtype([("position", "<f4", 3), ("material", "<u4"), ("normal", "<f4", 3), ("flags", "<u4")])
assert VERTEX_DTYPE.itemsize == 32


class TableAudit(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "regular_cases", "transition_cases", "regular_vertices", "regular_triangles",
        "transition_vertices", "transition_triangles", "max_regular_vertices", "max_regular_triangles",
        "max_transition_vertices", "max_transition_triangles")] + [("fingerprint", C.c_uint64)]


class FixtureMetrics(C.Structure):
    _fields_ = [("solid_samples", C.c_uint32), ("air_samples", C.c_uint32), ("active_cells", C.c_uint32),
                ("_pad", C.c_uint32), ("active_microbrick_mask", C.c_uint64), ("fingerprint", C.c_uint64)]


def build(force: bool = False) -> Path:
    """Compile oracle/libhvx_oracle.so with the committed Makefile (gcc, -ffp-contract=off)."""
    so = _HERE / "libhvx_oracle.so"
    src_mtime = max((_HERE / f).stat().st_mtime for f in ("hvx_oracle.c", "brick_oracle.c", "hvx_oracle.h", "tables.inc", "mc_tables.inc"))
    if force or not so.exists() or so.stat().st_mtime < src_mtime:
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-C", str(_HERE), "CC=gcc"], check=True, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(str(build()))
        L = _LIB
        u32p, i64p = C.POINTER(C.c_uint32), C.POINTER(C.c_int64)
        L.hvxo_cellword.restype = C.c_uint32
        L.hvxo_cellword.argtypes = [C.c_int16, C.c_uint8, C.c_uint8]
        L.hvxo_validate_tables.argtypes = [C.POINTER(TableAudit)]
        L.hvxo_table_revision.restype = C.c_char_p
        L.hvxo_sample_canonical.restype = C.c_uint32
        L.hvxo_sample_canonical.argtypes = [C.c_int, i64p, C.c_uint32]
        L.hvxo_fixture_fill.argtypes = [C.c_int, C.c_int, C.c_uint32, i64p, u32p]
        L.hvxo_fixture_metrics_of.argtypes = [C.c_int, u32p, C.POINTER(FixtureMetrics)]
        L.hvxo_slab_fill.argtypes = [C.c_int, C.c_int, C.c_uint32, i64p, u32p]
        L.hvxo_terrain_sdf_rolling.restype = C.c_float
        L.hvxo_terrain_sdf_rolling.argtypes = [C.c_float] * 3
        L.hvxo_extract_regular.argtypes = [C.c_int, u32p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32,
                                           C.c_void_p, C.c_uint32, u32p, C.c_uint32, u32p, u32p, u32p, u32p]
        L.hvxo_extract_transition.argtypes = [C.c_int, u32p, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32,
                                              C.c_void_p, C.c_uint32, u32p, C.c_uint32, u32p, u32p, u32p]
        L.hvxo_extract_transition_face_analytic.argtypes = [C.c_int, C.c_int, C.c_uint32, i64p, C.c_int, C.c_void_p,
                                                            C.c_uint32, u32p, C.c_uint32, u32p, u32p]
        L.hvxo_batch_regular.restype = C.c_int64
        L.hvxo_batch_regular.argtypes = [C.c_int, C.c_int, C.c_uint32, i64p, C.c_uint32, C.c_int, u32p, C.c_int,
                                         C.POINTER(C.c_uint64)]
        L.hvxo_build_meshlets.argtypes = [C.c_void_p, C.c_uint32, u32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                          C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32]
        L.hvxo_batch_fill.argtypes = [C.c_int, C.c_int, C.c_uint32, i64p, C.c_uint32, C.c_int, u32p]
        L.hvxo_page_hash.restype = C.c_uint32
        L.hvxo_page_hash.argtypes = [u32p, C.POINTER(C.c_int32), C.c_uint32]
        L.hvxo_gather_surface.restype = None
        L.hvxo_gather_surface.argtypes = [C.c_void_p, C.c_void_p, u32p, C.c_void_p, u32p, u32p, C.c_void_p, u32p]
        L.hvxo_case_topology.argtypes = [C.c_int, C.c_uint32, u32p, u32p, u32p, u32p, C.POINTER(C.c_uint16),
                                         C.POINTER(C.c_uint8)]
    return _LIB


def _u32p(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_uint32))


def _xyz(page_xyz):
    return (C.c_int64 * 3)(*[int(v) for v in page_xyz])


def cellword(density: int, material: int, flags: int) -> int:
    return int(lib().hvxo_cellword(density, material, flags))


def validate_tables() -> TableAudit:
    audit = TableAudit()
    rc = lib().hvxo_validate_tables(C.byref(audit))
    if rc:
        raise ValueError(f"table audit failed: {rc}")
    return audit


def case_topology(kind: int, case_index: int):
    cls, rev, nv, nt = (C.c_uint32() for _ in range(4))
    codes = (C.c_uint16 * 12)()
    tris = (C.c_uint8 * 36)()
    rc = lib().hvxo_case_topology(kind, case_index, C.byref(cls), C.byref(rev), C.byref(nv), C.byref(nt), codes, tris)
    if rc:
        raise ValueError("bad case")
    return dict(class_index=cls.value, reverse=bool(rev.value), vertex_count=nv.value, triangle_count=nt.value,
                codes=list(codes)[:nv.value], triangles=list(tris)[:3 * nt.value])


def sample_canonical(kind: int, position, lod: int = 0) -> int:
    return int(lib().hvxo_sample_canonical(kind, _xyz(position), lod))


def fixture_fill(kind: int, page_xyz, lod: int = 0, edge: int = 32) -> np.ndarray:
    out = np.empty((edge + 2) ** 3, dtype=np.uint32)
    rc = lib().hvxo_fixture_fill(kind, edge, lod, _xyz(page_xyz), _u32p(out))
    if rc:
        raise ValueError(f"fixture_fill failed: {rc}")
    return out


def fixture_metrics(samples: np.ndarray, edge: int = 32) -> FixtureMetrics:
    m = FixtureMetrics()
    rc = lib().hvxo_fixture_metrics_of(edge, _u32p(np.ascontiguousarray(samples, dtype=np.uint32)), C.byref(m))
    if rc:
        raise ValueError("metrics failed")
    return m


def slab_fill(kind: int, page_xyz, lod: int, edge: int = 32) -> np.ndarray:
    out = np.empty(6 * 3 * (2 * edge + 3) ** 2, dtype=np.uint32)
    rc = lib().hvxo_slab_fill(kind, edge, lod, _xyz(page_xyz), _u32p(out))
    if rc:
        raise ValueError(f"slab_fill failed: {rc}")
    return out


def terrain_sdf(x: float, y: float, z: float) -> float:
    return float(lib().hvxo_terrain_sdf_rolling(x, y, z))


class Mesh:
    """Result of one oracle extraction."""

    def __init__(self, vertices, indices, cell_words, cell_ranges, classify, counters):
        self.vertices = vertices
        self.indices = indices
        self.cell_words = cell_words
        self.cell_ranges = cell_ranges
        self.classify = classify
        self.counters = counters


def extract_regular(samples: np.ndarray, edge: int = 32, generation: int = 1, dirty_microbricks: int = (1 << 64) - 1,
                    transition_mask: int = 0, max_vertices: int | None = None, max_indices: int | None = None,
                    debug: bool = True) -> Mesh:
    samples = np.ascontiguousarray(samples, dtype=np.uint32)
    assert samples.size == (edge + 2) ** 3
    cells = edge ** 3
    maxv = 0xFFFFFFFF if max_vertices is None else max_vertices
    maxi = 0xFFFFFFFF if max_indices is None else max_indices
    cls = np.zeros(4, dtype=np.uint32)
    em = np.zeros(8, dtype=np.uint32)
    # pass 1: counts only
    lib().hvxo_extract_regular(edge, _u32p(samples), generation, dirty_microbricks, transition_mask, maxv, maxi,
                               None, 0, None, 0, None, None, _u32p(cls), _u32p(em))
    nv, ni = int(em[0]), int(em[1])
    verts = np.zeros(nv, dtype=VERTEX_DTYPE)
    idx = np.zeros(ni, dtype=np.uint32)
    words = np.zeros((cells, 4), dtype=np.uint32) if debug else None
    ranges = np.full((cells, 2), 0xFFFFFFFF, dtype=np.uint32) if debug else None
    lib().hvxo_extract_regular(edge, _u32p(samples), generation, dirty_microbricks, transition_mask, maxv, maxi,
                               verts.ctypes.data_as(C.c_void_p), nv, _u32p(idx), ni, _u32p(words), _u32p(ranges),
                               _u32p(cls), _u32p(em))
    return Mesh(verts, idx, words, ranges, cls, em)


def extract_transition(slabs: np.ndarray, transition_mask: int, edge: int = 32, generation: int = 1,
                       max_vertices: int | None = None, max_indices: int | None = None, debug: bool = True) -> Mesh:
    slabs = np.ascontiguousarray(slabs, dtype=np.uint32)
    assert slabs.size == 6 * 3 * (2 * edge + 3) ** 2
    cells = 6 * edge * edge
    maxv = 0xFFFFFFFF if max_vertices is None else max_vertices
    maxi = 0xFFFFFFFF if max_indices is None else max_indices
    ctr = np.zeros(12, dtype=np.uint32)
    rc = lib().hvxo_extract_transition(edge, _u32p(slabs), transition_mask, generation, maxv, maxi, None, 0, None, 0,
                                       None, None, _u32p(ctr))
    if rc:
        raise ValueError(f"extract_transition failed: {rc}")
    nv, ni = int(ctr[2]), int(ctr[3])
    verts = np.zeros(nv, dtype=VERTEX_DTYPE)
    idx = np.zeros(ni, dtype=np.uint32)
    words = np.zeros((cells, 4), dtype=np.uint32) if debug else None
    ranges = np.full((cells, 2), 0xFFFFFFFF, dtype=np.uint32) if debug else None
    lib().hvxo_extract_transition(edge, _u32p(slabs), transition_mask, generation, maxv, maxi,
                                  verts.ctypes.data_as(C.c_void_p), nv, _u32p(idx), ni, _u32p(words), _u32p(ranges),
                                  _u32p(ctr))
    return Mesh(verts, idx, words, ranges, None, ctr)


def extract_transition_face_analytic(kind: int, page_xyz, lod: int, face: int, edge: int = 32):
    vcap, icap = edge * edge * 12, edge * edge * 36
    verts = np.zeros(vcap, dtype=VERTEX_DTYPE)
    idx = np.zeros(icap, dtype=np.uint32)
    nv, ni = C.c_uint32(), C.c_uint32()
    rc = lib().hvxo_extract_transition_face_analytic(kind, edge, lod, _xyz(page_xyz), face,
                                                     verts.ctypes.data_as(C.c_void_p), vcap, _u32p(idx), icap,
                                                     C.byref(nv), C.byref(ni))
    if rc:
        raise ValueError(f"analytic face failed: {rc}")
    return verts[:nv.value].copy(), idx[:ni.value].copy()


def build_meshlets(vertices: np.ndarray, indices: np.ndarray, first_index=0, first_vertex=0, first_bounds=0,
                   generation=1, flags=0):
    """build_terrain_meshlets: (descriptors, bounds); raises ValueError with the reference's error kind."""
    vertices = np.ascontiguousarray(vertices, dtype=VERTEX_DTYPE)
    indices = np.ascontiguousarray(indices, dtype=np.uint32)
    cap = (indices.size + 62) // 63 + 1
    meshlets = np.zeros(cap, dtype=MESHLET_DTYPE)
    bounds = np.zeros(cap, dtype=MESHLET_BOUNDS_DTYPE)
    rc = lib().hvxo_build_meshlets(vertices.ctypes.data_as(C.c_void_p), vertices.size, _u32p(indices), indices.size,
                                   first_index, first_vertex, first_bounds, generation, flags,
                                   meshlets.ctypes.data_as(C.c_void_p), bounds.ctypes.data_as(C.c_void_p), cap)
    if rc < 0:
        raise ValueError({-1: "IncompleteTriangle", -2: "IndexOutOfBounds", -3: "NonFinitePosition", -4: "capacity"}[rc])
    return meshlets[:rc], bounds[:rc]


def max_threads() -> int:
    return int(lib().hvxo_max_threads())


def batch_fill(kind: int, pages: np.ndarray, edge: int, lod: int = 0, threads: int = 1) -> np.ndarray:
    pages = np.ascontiguousarray(pages, dtype=np.int64).reshape(-1, 3)
    out = np.empty(pages.shape[0] * (edge + 2) ** 3, dtype=np.uint32)
    rc = lib().hvxo_batch_fill(kind, edge, lod, pages.ctypes.data_as(C.POINTER(C.c_int64)), pages.shape[0], threads,
                               _u32p(out))
    if rc:
        raise ValueError(f"batch_fill failed: {rc}")
    return out


def batch_regular(kind: int, pages: np.ndarray, edge: int, lod: int = 0, threads: int = 1, do_fill: bool = True,
                  samples: np.ndarray | None = None):
    """CPU baseline: fill (optional) + regular extraction of every chunk, OpenMP over chunks."""
    pages = np.ascontiguousarray(pages, dtype=np.int64).reshape(-1, 3)
    totals = (C.c_uint64 * 4)()
    smp = None if samples is None else _u32p(np.ascontiguousarray(samples, dtype=np.uint32))
    cells = lib().hvxo_batch_regular(kind, edge, lod, pages.ctypes.data_as(C.POINTER(C.c_int64)), pages.shape[0],
                                     1 if do_fill else 0, smp, threads, totals)
    return int(cells), [int(t) for t in totals]


def gather_surface(residency: np.ndarray, table: np.ndarray, atlas: np.ndarray, job: np.ndarray):
    """One job of the surface gather (hvxo_gather_surface): (regular[39304], transition[80802], counters, indirect[24]).

    residency / table / job are the structured arrays of oracle/residency.py; slab words of faces outside
    the job's mask keep the sentinel 0xFFFFFFFF."""
    from . import residency as R
    residency = np.ascontiguousarray(residency, dtype=R.RESIDENCY_DTYPE)
    table = np.ascontiguousarray(table, dtype=R.ENTRY_DTYPE)
    job = np.ascontiguousarray(job, dtype=R.JOB_DTYPE)
    atlas = np.ascontiguousarray(atlas, dtype=np.uint32).reshape(-1)
    regular = np.full(34 ** 3, 0xFFFFFFFF, dtype=np.uint32)
    transition = np.full(6 * 3 * 67 * 67, 0xFFFFFFFF, dtype=np.uint32)
    counters = np.zeros(1, dtype=R.COUNTERS_DTYPE)
    indirect = np.zeros(24, dtype=np.uint32)
    lib().hvxo_gather_surface(residency.ctypes.data_as(C.c_void_p), table.ctypes.data_as(C.c_void_p), _u32p(atlas),
                              job.ctypes.data_as(C.c_void_p), _u32p(regular), _u32p(transition),
                              counters.ctypes.data_as(C.c_void_p), _u32p(indirect))
    return regular, transition, counters[0], indirect


def page_hash(planet_id, relative_min, lod: int) -> int:
    pid = np.ascontiguousarray(planet_id, dtype=np.uint32)
    rel = np.ascontiguousarray(relative_min, dtype=np.int32)
    return int(lib().hvxo_page_hash(_u32p(pid), rel.ctypes.data_as(C.POINTER(C.c_int32)), lod))


def brick_extract(voxel_words: np.ndarray, data_offset: int, origin, voxel_size: float):
    """Legacy 8^3-brick marching cubes (oracle/brick_oracle.c) for one brick, cells in linear order.
    Returns (vertices [n,4], normals [n,4], indices [n], raw_count); n = min(raw_count, 2048) is what the shader
    publishes, entries of cells dropped by the overflow rule are left zero."""
    L = lib()
    L.hvxo_brick_extract.argtypes = [C.POINTER(C.c_uint32), C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                     C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    words = np.ascontiguousarray(voxel_words, dtype=np.uint32)
    os_ = (C.c_float * 4)(*[float(v) for v in origin], float(voxel_size))
    v = np.zeros((2048, 4), dtype=np.float32)
    n = np.zeros((2048, 4), dtype=np.float32)
    i = np.zeros(2048, dtype=np.uint32)
    raw = C.c_uint32()
    rc = L.hvxo_brick_extract(_u32p(words), data_offset, os_, v.ctypes.data_as(C.POINTER(C.c_float)),
                              n.ctypes.data_as(C.POINTER(C.c_float)), _u32p(i), C.byref(raw))
    assert rc == 0
    count = min(raw.value, 2048)
    return v[:count], n[:count], i[:count], raw.value
