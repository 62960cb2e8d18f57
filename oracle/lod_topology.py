"""CPU ORACLE (test infrastructure) for the mixed-LOD topology: a plain-Python restatement of
PV/src/lod_topology.rs, used to check helio_b200/csrc/lod_topology.cpp.  Python ints are unbounded,
so the reference's i128/u128 arithmetic needs no special care.

Not imported by the product package.
"""
from __future__ import annotations

MAX_ADDRESSABLE_LOD = 57
TANGENT_AXES = (0, 2)


class TopologyError(Exception):
    def __init__(self, kind, detail=None):
        super().__init__(f"{kind}: {detail}")
        self.kind = kind


def _bounds(page, edge):
    """PageBounds::new (PV/src/lod_topology.rs:238-246): page = (lod, (x, y, z))."""
    lod, xyz = page
    span = edge << lod
    mn = tuple(v * span for v in xyz)
    return mn, tuple(v + span for v in mn)


def _overlaps(a, b):
    return all(a[0][i] < b[1][i] and b[0][i] < a[1][i] for i in range(3))


def _shared_face(a, b):
    """PV/src/lod_topology.rs:252-272: (axis, a_is_positive) or None."""
    for axis in range(3):
        pos, neg = a[1][axis] == b[0][axis], b[1][axis] == a[0][axis]
        if not pos and not neg:
            continue
        if all(a[0][o] < b[1][o] and b[0][o] < a[1][o] for o in range(3) if o != axis):
            return axis, pos
    return None


def topology(pages, edge=32):
    """TerrainLodTopology::new (PV/src/lod_topology.rs:27-100) -> ({page: mask}, stats dict)."""
    unique = set()
    for page in pages:
        page = (int(page[0]), tuple(int(v) for v in page[1]))
        if page[0] > MAX_ADDRESSABLE_LOD:
            raise TopologyError("Address")
        if any(abs(v * (edge << page[0])) >= 2 ** 63 for v in page[1]):
            raise TopologyError("Address")
        if page in unique:
            raise TopologyError("DuplicatePage", page)
        unique.add(page)
    if not unique:
        raise TopologyError("Empty")
    ordered = sorted(unique)
    bounds = [_bounds(p, edge) for p in ordered]
    masks = {p: 0 for p in ordered}
    for l in range(len(ordered)):
        for r in range(l + 1, len(ordered)):
            if _overlaps(bounds[l], bounds[r]):
                raise TopologyError("OverlappingPages", (ordered[l], ordered[r]))
            face = _shared_face(bounds[l], bounds[r])
            if face is None:
                continue
            axis, l_positive = face
            diff = abs(ordered[l][0] - ordered[r][0])
            if diff > 1:
                raise TopologyError("UnbalancedFace", (ordered[l], ordered[r]))
            if diff == 1:
                if ordered[l][0] > ordered[r][0]:
                    coarse, positive = ordered[l], l_positive
                else:
                    coarse, positive = ordered[r], not l_positive
                masks[coarse] |= 1 << (2 * axis + (1 if positive else 0))
    stats = dict(pages=len(ordered), minimum_lod=ordered[0][0], maximum_lod=ordered[-1][0],
                 transition_faces=sum(bin(m).count("1") for m in masks.values()))
    return masks, stats


def _address(lod, cell, edge):
    span = edge << lod
    return lod, tuple(c // span for c in cell)  # Python // is Euclidean for a positive divisor


def _tangent_children(parent):
    lod, (x, _, z) = parent
    if lod == 0:
        raise TopologyError("CannotRefineLod0")
    return [(lod - 1, (2 * x, -1, 2 * z)), (lod - 1, (2 * x + 1, -1, 2 * z)),
            (lod - 1, (2 * x, -1, 2 * z + 1)), (lod - 1, (2 * x + 1, -1, 2 * z + 1))]


def _tangent_shared_edge(a, b):
    for axis, other in ((0, 2), (2, 0)):
        if (a[1][axis] == b[0][axis] or b[1][axis] == a[0][axis]) and a[0][other] < b[1][other] and b[0][other] < a[1][other]:
            return True
    return False


def _balance(leaves, edge, max_pages):
    """balance_tangent_leaves (PV/src/lod_topology.rs:309-346)."""
    while True:
        ordered = sorted(leaves)
        bounds = [_bounds(p, edge) for p in ordered]
        coarse = None
        for l in range(len(ordered)):
            for r in range(l + 1, len(ordered)):
                if _tangent_shared_edge(bounds[l], bounds[r]) and abs(ordered[l][0] - ordered[r][0]) > 1:
                    coarse = ordered[l] if ordered[l][0] > ordered[r][0] else ordered[r]
                    break
            if coarse is not None:
                break
        if coarse is None:
            return
        if len(leaves) + 3 > max_pages:
            raise TopologyError("PageBudget", len(leaves) + 3)
        leaves.remove(coarse)
        leaves.update(_tangent_children(coarse))


def horizon_plan(focus, root_lod, minimum_lod, max_pages, edge=32):
    """HorizonLodFixturePlan::build_with_minimum_lod (PV/src/lod_topology.rs:169-217)."""
    if root_lod == 0 or root_lod > MAX_ADDRESSABLE_LOD:
        raise TopologyError("UnsupportedRootLod", root_lod)
    if minimum_lod >= root_lod:
        raise TopologyError("UnsupportedMinimumLod")
    if max_pages < 4:
        raise TopologyError("PageBudget", 4)
    focus = tuple(int(v) for v in focus)
    lod, xyz = _address(root_lod, focus, edge)
    root = (lod, (xyz[0], -1, xyz[2]))
    leaves = {root}
    for target_lod in range(root_lod - 1, minimum_lod - 1, -1):
        lod, xyz = _address(target_lod + 1, focus, edge)
        target = (lod, (xyz[0], -1, xyz[2]))
        if target not in leaves:
            raise TopologyError("MissingRefinementParent", target)
        leaves.remove(target)
        leaves.update(_tangent_children(target))
        _balance(leaves, edge, max_pages)
        if len(leaves) > max_pages:
            raise TopologyError("PageBudget", len(leaves))
    masks, stats = topology(leaves, edge)
    # validate_tangent_root (PV/src/lod_topology.rs:123-146)
    rb = _bounds(root, edge)
    covered = 0
    for page in leaves:
        b = _bounds(page, edge)
        if not all(rb[0][a] <= b[0][a] and b[1][a] <= rb[1][a] for a in TANGENT_AXES):
            raise TopologyError("OutsideTangentRoot", page)
        covered += (b[1][0] - b[0][0]) * (b[1][2] - b[0][2])
    if covered != (rb[1][0] - rb[0][0]) * (rb[1][2] - rb[0][2]):
        raise TopologyError("TangentCoverage")
    return root, masks, stats
