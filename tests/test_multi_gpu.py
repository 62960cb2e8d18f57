"""Multi-process sharding: static LPT partition + optional final gather reproduce the single-process
result byte for byte.  CPU: world_size 2 over gloo (host logic only, the oracle stands
in for the kernels).  GPU: world_size 2 over NCCL with the CUDA extractor, when two GPUs exist."""
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import helio_b200 as H
from oracle import oracle as O

HERE = Path(__file__).resolve().parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_world(world, backend, tmp_path):
    port = _free_port()
    procs = [subprocess.Popen([sys.executable, str(HERE / "_dist_worker.py"), str(r), str(world), str(port), backend,
                               str(tmp_path)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(world)]
    for r, proc in enumerate(procs):
        out, _ = proc.communicate(timeout=300)
        assert proc.returncode == 0, f"rank {r} failed:\n{out[-3000:]}"
    return np.load(tmp_path / "gathered.npz")


def _expected():
    sys.path.insert(0, str(HERE))
    from _dist_worker import chunk_list
    chunks = chunk_list()
    meshes = [O.extract_regular(O.fixture_fill(O.FIELD_PLANE, page, lod=lod), transition_mask=mask, debug=False)
              for lod, page, mask in chunks]
    return chunks, meshes


def _check(got, chunks, meshes, world):
    assert got["owner"].size == len(chunks) and set(got["owner"]) == set(range(world))
    verts = got["vertices"].view(np.uint8).view(H.VERTEX_DTYPE)
    tv = ti = 0
    for g, mesh in enumerate(meshes):
        fv, nv, fi, ni = (int(x) for x in got["ranges"][g])
        assert (fv, nv, fi, ni) == (tv, len(mesh.vertices), ti, len(mesh.indices)), g
        assert verts[fv:fv + nv].tobytes() == mesh.vertices.tobytes(), g
        assert np.array_equal(got["indices"][fi:fi + ni].view(np.uint32), mesh.indices), g
        tv += nv
        ti += ni
    assert tv == len(verts) and ti == got["indices"].size
    assert any(len(m.vertices) for m in meshes) and any(mask for _, _, mask in chunks)


def test_two_ranks_over_gloo_reproduce_the_single_process_result(tmp_path):
    chunks, meshes = _expected()
    _check(_run_world(2, "gloo", tmp_path), chunks, meshes, 2)


@pytest.mark.gpu
def test_two_gpus_over_nccl_reproduce_the_single_gpu_result(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    chunks, meshes = _expected()
    _check(_run_world(2, "nccl", tmp_path), chunks, meshes, 2)


@pytest.mark.gpu
def test_single_gpu_batch_equals_the_oracle_for_a_mixed_lod_plan():
    """The same planet patch on one GPU: regular + transition extraction of every page of the plan."""
    chunks, meshes = _expected()
    n = len(chunks)
    batch = H.ChunkBatchExtractor(0, edge=32, max_chunks=n, max_vertices=16384, max_indices=24576,
                                  max_transition_vertices=8192, max_transition_indices=24576)
    pages, lods, masks = [p for _, p, _ in chunks], [l for l, _, _ in chunks], [m for _, _, m in chunks]
    batch.fill_density(int(H.ExtractionFixtureKind.Plane), pages, lods)
    batch.extract_regular(None, n, transition_mask=masks)
    verts, idx, ranges = batch.ctx.read_meshes(0, 0, n)
    for g, mesh in enumerate(meshes):
        r = ranges[g]
        assert verts[r["first_vertex"]:r["first_vertex"] + r["vertex_count"]].tobytes() == mesh.vertices.tobytes(), g
        assert np.array_equal(idx[r["first_index"]:r["first_index"] + r["index_count"]], mesh.indices), g
    coarse = [g for g in range(n) if lods[g] >= 1]
    batch.fill_slabs(int(H.ExtractionFixtureKind.Plane), [pages[g] for g in coarse], [lods[g] for g in coarse])
    batch.extract_transition(None, len(coarse), [masks[g] for g in coarse])
    counters = batch.transition_counters(len(coarse))
    for j, g in enumerate(coarse):
        want = O.extract_transition(O.slab_fill(O.FIELD_PLANE, pages[g], lods[g]), masks[g], debug=False)
        assert counters["required_vertices"][j] == len(want.vertices) and counters["active_faces"][j] == bin(masks[g]).count("1")
        v, i = batch.chunk_mesh(j, kind=1)
        assert v.tobytes() == want.vertices.tobytes() and np.array_equal(i, want.indices), g
    batch.close()
