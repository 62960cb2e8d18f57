"""CPU: the publication oracle against the reference's own sequence of known answers.

PV/tests/gpu_surface_publication.rs:182-520 (publication_is_atomic_generation_safe_and_visibility_gated):
stale metadata -> overflow -> incomplete -> success (bank flip, counts, meshlet counts, draw arguments,
feedback) -> visibility on / off."""
import numpy as np

from oracle import publication as P


def reference_fixture():
    pub = P.Publisher(1, 32, 64, 16, 48)
    pub.states[0] = (41, 0, 0, 1, 7, 11, 5, 7, 0, 0, (0, 0))
    pub.regular_draws[0] = (11, 1, 7, 3, 0)
    pub.transition_draws[0] = (7, 1, 5, 2, 0)
    job = pub.job(0, 43)
    meta = np.zeros(1, dtype=P.PAGE_META_DTYPE)
    meta[0] = ((0, 0, 0), 0, 0, 42, 0, 0)
    regular = np.array([0, 0, 19, 27, 0, 0, 1, 0], dtype=np.uint32)
    transition = np.array([0, 0, 0, 0, 13, 21, 0, 0, 1, 0, 0, 0], dtype=np.uint32)
    return pub, job, meta, regular, transition


def run_reference_sequence(publish, refresh, read):
    """Drive `publish(meta, regular, transition)` / `refresh(visible)` through the reference test's steps."""
    pub0, job, meta, regular, transition = reference_fixture()
    old_state, old_r, old_t = pub0.states[0].copy(), pub0.regular_draws[0].copy(), pub0.transition_draws[0].copy()
    publish(meta, regular, transition)                                           # metadata generation 42 != job 43
    s, r, t, f = read()
    assert s == old_state and r == old_r and t == old_t
    assert tuple(f)[:5] == (1, 0, 1, 0, 0)
    meta["generation_low"] = 43
    overflow = regular.copy()
    overflow[4] = 1
    publish(meta, overflow, transition)
    s, r, t, f = read()
    assert s == old_state and r == old_r and tuple(f)[:5] == (2, 0, 1, 1, 0)
    incomplete = regular.copy()
    incomplete[6] = 0
    publish(meta, incomplete, transition)
    s, r, t, f = read()
    assert s == old_state and r == old_r and tuple(f)[:5] == (3, 0, 1, 1, 1)
    publish(meta, regular, transition)
    s, r, t, f = read()
    assert tuple(s)[:10] == (43, 0, 1, 1, 19, 27, 13, 21, 1, 1)
    assert tuple(r) == (27, 0, 64, 32, 0) and tuple(t) == (21, 0, 48, 16, 0)
    assert tuple(f)[:5] == (4, 1, 1, 1, 1)
    refresh([1])
    s, r, t, f = read()
    assert r["instance_count"] == 1 and t["instance_count"] == 1
    refresh([0])
    s, r, t, f = read()
    assert r["instance_count"] == 0 and t["instance_count"] == 0


def test_publication_is_atomic_generation_safe_and_visibility_gated():
    pub, job, _, _, _ = reference_fixture()
    run_reference_sequence(lambda m, rc, tc: pub.publish(job, m, rc, tc), pub.refresh_visibility,
                           lambda: (pub.states[0].copy(), pub.regular_draws[0].copy(), pub.transition_draws[0].copy(), pub.feedback[0].copy()))


def test_meshes_land_in_the_inactive_bank_and_only_emitted_elements_move():
    pub = P.Publisher(3, 8, 12, 4, 6)
    job = pub.job(2, 5)
    meta = np.zeros(3, dtype=P.PAGE_META_DTYPE)
    meta[2] = ((0, 0, 0), 0, 2, 5, 0, 0)
    sv = np.zeros(8, dtype=pub.vertices.dtype)
    sv["material"] = np.arange(8) + 100
    si = np.arange(12, dtype=np.uint32) + 7
    rc = np.array([5, 9, 5, 9, 0, 0, 1, 0], dtype=np.uint32)
    tc = np.array([0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0], dtype=np.uint32)
    pub.vertices["material"] = 0xAAAA
    pub.indices[:] = 0xBBBB
    pub.publish(job, meta, rc, tc, sv, si, np.zeros(4, dtype=sv.dtype), np.zeros(6, dtype=np.uint32))
    bank = 2 * 2 + 1                                         # active bank 0 -> the mesh goes to bank 1 of slot 2
    assert list(pub.vertices["material"][bank * 8:bank * 8 + 8]) == [100, 101, 102, 103, 104, 0xAAAA, 0xAAAA, 0xAAAA]
    assert list(pub.indices[bank * 12:bank * 12 + 12]) == list(range(7, 16)) + [0xBBBB] * 3
    assert (pub.vertices["material"][:bank * 8] == 0xAAAA).all() and pub.states[2]["active_bank"] == 1
    assert tuple(pub.regular_draws[2]) == (9, 0, bank * 12, bank * 8, 2)
    pub.publish(pub.job(2, 6), meta, rc, tc, sv, si, None, None)   # stale: metadata still says generation 5
    assert pub.states[2]["generation_low"] == 5 and pub.feedback[0]["stale_rejections"] == 1
