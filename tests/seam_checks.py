"""Seam checks shared by the CPU (oracle) and GPU versions of tests/test_lod_seams.py."""
import numpy as np

import helio_b200 as H
from oracle import oracle as O

EDGE = 32
# transition_face_basis (PV/src/transvoxel_transition.rs:399-410): origin corner, u axis, v axis, outward normal
BASIS = [
    ((0, 0, 1), (0, 1, 0), (0, 0, -1), (-1, 0, 0)), ((1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 0, 0)),
    ((1, 0, 0), (0, 0, 1), (-1, 0, 0), (0, -1, 0)), ((0, 1, 0), (0, 0, 1), (1, 0, 0), (0, 1, 0)),
    ((0, 1, 0), (1, 0, 0), (0, -1, 0), (0, 0, -1)), ((0, 0, 1), (1, 0, 0), (0, 1, 0), (0, 0, 1)),
]


def quantize(a, scale=100_000.0):
    return np.rint(np.asarray(a, dtype=np.float64) * scale).astype(np.int64)


def face_uv(positions, face):
    origin, u_axis, v_axis, _ = (np.array(b, dtype=np.float64) for b in BASIS[face])
    rel = positions.astype(np.float64) - origin * EDGE
    return rel @ u_axis, rel @ v_axis



def plan_pages(kind):
    focus = [4000, -20, -17] if kind == O.FIELD_PLANE else [40, -20, -17]
    plan = H.HorizonLodFixturePlan.build_with_minimum_lod(focus, 3, 0, 192, EDGE)
    keys = list(plan.topology().transition_masks().items())
    assert sorted({k.lod for k, _ in keys}) == [0, 1, 2]
    pages = np.array([k.page_xyz for k, _ in keys], dtype=np.int64)
    lods = np.array([k.lod for k, _ in keys], dtype=np.uint8)
    masks = [m for _, m in keys]
    return keys, pages, lods, masks


def check_plan_seams(keys, pages, lods, masks, regular, transition):
    """regular[g]: vertex array of page g; transition[g]: (vertices, indices) of coarse page g (mask != 0)."""
    n = len(keys)
    checked_faces = checked_vertices = 0
    for g in range(n):
        if not masks[g]:
            continue
        tv, ti = transition[g]
        assert len(ti) % 3 == 0 and (len(tv) == 0 or int(ti.max()) < len(tv))
        lod = int(lods[g])
        span = EDGE << lod                                   # the coarse page in LOD0 cells
        origin = pages[g] * span
        for face in range(6):
            if not (masks[g] >> face) & 1:
                continue
            axis, positive = face // 2, face & 1
            on_face = tv[tv["flags"] == (1 << face)]
            plane = float(EDGE if positive else 0)
            full_res = on_face[np.abs(on_face["position"][:, axis] - plane) <= 1e-6]     # depth 0: the fine side
            theirs = set()
            neighbours = 0
            covered = np.zeros(len(full_res), dtype=bool)    # the plan is a tangent sheet: not every face is fully lined with finer pages
            for h in range(n):
                if int(lods[h]) != lod - 1:
                    continue
                fspan = EDGE << (lod - 1)
                forigin = pages[h] * fspan
                lo, hi = forigin, forigin + fspan
                touching = (lo[axis] == origin[axis] + span) if positive else (hi[axis] == origin[axis])
                inside = all(lo[a] >= origin[a] and hi[a] <= origin[a] + span for a in range(3) if a != axis)
                if not (touching and inside):
                    continue
                neighbours += 1
                assert not (masks[h] >> (face ^ 1)) & 1, "seams are coarse-owned: the fine side keeps primary positions"
                flo, fhi = (lo - origin) / (1 << lod), (hi - origin) / (1 << lod)     # the fine page in coarse local cells
                covered |= np.all([(full_res["position"][:, a] >= flo[a] - 1e-6) & (full_res["position"][:, a] <= fhi[a] + 1e-6)
                                   for a in range(3) if a != axis], axis=0)
                fv = regular[h]
                fplane = 0.0 if positive else float(EDGE)
                boundary = fv[np.abs(fv["position"][:, axis] - fplane) <= 1e-6]
                # fine local cells -> coarse local cells
                mapped = (forigin - origin)[None, :].astype(np.float64) / (1 << lod) + boundary["position"].astype(np.float64) * 0.5
                theirs |= {tuple(q) for q in quantize(mapped, 10_000.0)}
            assert neighbours >= 1, f"page {keys[g][0]} face {face}: no finer neighbour"
            ours = {tuple(q) for q in quantize(full_res["position"][covered], 10_000.0)}
            assert ours == theirs, (f"page {keys[g][0]} face {face}: {len(ours - theirs)} transition vertices without a fine twin, "
                                    f"{len(theirs - ours)} fine boundary vertices without a transition twin")
            checked_faces += 1
            checked_vertices += len(ours)
    assert checked_faces == sum(bin(m).count("1") for m in masks) and checked_vertices > 100
    return checked_faces, checked_vertices


def random_slabs(seed):
    """Hashed densities on every fine sample of every face plane (the reference's splitmix-style hash of the fine
    sample coordinates, PV/src/transvoxel_transition.rs:863-886): |density| in [2048, 32767], never zero, so no vertex
    sits on a corner.  The three layers of a slab hold the same plane."""
    w = 2 * EDGE + 3
    v, u = np.meshgrid(np.arange(w, dtype=np.uint64), np.arange(w, dtype=np.uint64), indexing="ij")
    slabs = np.zeros((6, 3, w, w), dtype=np.uint32)
    for face in range(6):
        h = np.uint64(seed * 6 + face) ^ (u * np.uint64(0x9E3779B97F4A7C15)) ^ (v * np.uint64(0xD1B54A32D192ED03))
        h ^= h >> np.uint64(30)
        h *= np.uint64(0xBF58476D1CE4E5B9)
        h ^= h >> np.uint64(27)
        h *= np.uint64(0x94D049BB133111EB)
        h ^= h >> np.uint64(31)
        # the reference draws |density| from [1, 32767] on an 8 x 8 patch; on the whole 32 x 32 face that produces a few
        # slivers thinner than its own 1e-12 area bound, so the floor is raised (every edge parameter in [1/17, 16/17])
        magnitude = ((h >> np.uint64(17)) % np.uint64(30720) + np.uint64(2048)).astype(np.int64)
        density = np.where(h & np.uint64(1) == 0, magnitude, -magnitude)
        words = (density.astype(np.int16).view(np.uint16).astype(np.uint32)) | ((density <= 0).astype(np.uint32) << np.uint32(16))
        slabs[face, :, :, :] = words[None, :, :]
    return slabs.reshape(-1)



def check_random_faces(verts, idx, words, first_v, first_i, seed):
    """words: packed transition cell records [6 * E^2]; first_v / first_i: chunk-local first vertex / index per cell."""
    per_face = EDGE * EDGE
    nv = ((words >> 17) & 0xF).astype(np.int64)
    nt = ((words >> 21) & 0xF).astype(np.int64)
    first_v, first_i = first_v.astype(np.int64), first_i.astype(np.int64)
    assert int(nv.sum()) == len(verts) and int(3 * nt.sum()) == len(idx)
    qpos, qnrm = quantize(verts["position"]), quantize(verts["normal"])
    for face in range(6):
        cell0 = face * per_face
        sl = slice(int(first_v[cell0]), int(first_v[cell0 + per_face - 1] + nv[cell0 + per_face - 1]))
        assert np.all(verts["flags"][sl] == (1 << face)), f"face {face}: flags"
        fu, fv = face_uv(verts["position"], face)

        def keys(cell, axis_values, boundary):
            a, b = int(first_v[cell]), int(first_v[cell] + nv[cell])
            on = np.nonzero(np.abs(axis_values[a:b] - boundary) <= 1e-6)[0] + a
            return {(tuple(qpos[k]), tuple(qnrm[k]), int(verts["material"][k])) for k in on}

        nonempty = 0
        for v in range(EDGE):
            for u in range(EDGE - 1):
                left, right = keys(cell0 + v * EDGE + u, fu, u + 1.0), keys(cell0 + v * EDGE + u + 1, fu, u + 1.0)
                assert left == right, f"face {face} seed {seed} u seam {u + 1} row {v}"
                nonempty += bool(left)
        for v in range(EDGE - 1):
            for u in range(EDGE):
                bottom, top = keys(cell0 + v * EDGE + u, fv, v + 1.0), keys(cell0 + (v + 1) * EDGE + u, fv, v + 1.0)
                assert bottom == top, f"face {face} seed {seed} v seam {v + 1} column {u}"
                nonempty += bool(bottom)
        assert nonempty > 500, f"face {face}: a random field crosses most seams"
        a, b = int(first_i[cell0]), int(first_i[cell0 + per_face - 1] + 3 * nt[cell0 + per_face - 1])
        tri = idx[a:b].reshape(-1, 3).astype(np.int64)
        p = verts["position"][tri].astype(np.float64)                     # [t, 3, 3]
        area = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
        assert np.all((area * area).sum(axis=1) > 1.0e-12), f"face {face} seed {seed}: degenerate triangle"
        corners = [tuple(sorted(tuple(qpos[k]) for k in t)) for t in tri]
        assert all(t[0] != t[1] and t[1] != t[2] for t in corners), f"face {face} seed {seed}: collapsed triangle"
        assert len(set(corners)) == len(corners), f"face {face} seed {seed}: duplicated triangle"
