"""TEST INFRASTRUCTURE -- CPU restatement of the reference's bounded extraction publisher.

Follows PV/src/extraction.rs: SurfaceCounts::validate :230-243, BoundedExtractionPublisher
:342-603, RangeAllocator :605-664.  Pinned by replaying the reference's own unit tests
(PV/src/extraction.rs:731-868) in tests/test_extraction_publisher.py; the product's C++
(helio_b200/csrc/extraction_publisher.cpp) is then checked against this model on random
operation sequences.  Only tests/ may import this module.
"""
from __future__ import annotations


class Error(Exception):
    def __init__(self, kind, **fields):
        super().__init__(kind)
        self.kind, self.fields = kind, fields


class RangeAllocator:
    def __init__(self, capacity):
        self.capacity = capacity
        self.free = [[0, capacity]]          # [first, count], sorted by first, coalesced

    def reserve(self, count):
        if count == 0:
            return (0, 0)
        for i, (first, have) in enumerate(self.free):
            if have >= count:
                self.free[i] = [first + count, have - count]
                if have == count:
                    del self.free[i]
                return (first, count)
        return None

    def release(self, released):
        first, count = released
        if count == 0:
            return
        at = 0
        while at < len(self.free) and self.free[at][0] < first:
            at += 1
        self.free.insert(at, [first, count])
        i = max(at - 1, 0)
        while i + 1 < len(self.free):
            end = self.free[i][0] + self.free[i][1]
            if end < self.free[i + 1][0]:
                i += 1
                continue
            assert end == self.free[i + 1][0]
            self.free[i][1] += self.free[i + 1][1]
            del self.free[i + 1]

    def used(self):
        return self.capacity - sum(c for _, c in self.free)


def validate_counts(vertices, indices, meshlets):
    if indices % 3:
        raise Error("NonTriangleIndexCount", indices=indices)
    if (vertices == 0 or indices == 0) and (vertices, indices, meshlets) != (0, 0, 0):
        raise Error("IncompleteSurfaceCounts")
    if meshlets == 0 and indices != 0:
        raise Error("IncompleteSurfaceCounts")


class Publisher:
    """Reservations / surfaces are tuples: reservation = (key, generation, allocation), surface =
    (generation, allocation), allocation = ((vf, vc), (if, ic), (mf, mc))."""

    def __init__(self, max_page_slots, max_pending_pages, max_vertices, max_indices, max_meshlets):
        self.max_pending = max_pending_pages
        self.v, self.i, self.m = RangeAllocator(max_vertices), RangeAllocator(max_indices), RangeAllocator(max_meshlets)
        self.pages = {}                      # key -> [current | None, pending | None]
        self.c = dict(reservations=0, publications=0, replacements=0, cancellations=0, evictions=0, stale_rejected=0,
                      backpressured=0, pending_high_water=0, vertex_high_water=0, index_high_water=0, meshlet_high_water=0)

    def current(self, key):
        return self.pages.get(key, [None, None])[0]

    def pending(self, key):
        return self.pages.get(key, [None, None])[1]

    def counters(self):
        out = dict(self.c)
        out["current_pages"] = sum(1 for s in self.pages.values() if s[0] is not None)
        out["pending_pages"] = sum(1 for s in self.pages.values() if s[1] is not None)
        out["used_vertices"], out["used_indices"], out["used_meshlets"] = self.v.used(), self.i.used(), self.m.used()
        return out

    def _release(self, allocation):
        self.v.release(allocation[0])
        self.i.release(allocation[1])
        self.m.release(allocation[2])

    def reserve(self, key, generation, counts):
        validate_counts(*counts)
        current = self.current(key)
        if current is not None:
            if generation < current[0]:
                self.c["stale_rejected"] += 1
                return ("Stale", current[0])
            if generation == current[0]:
                return ("Current", current)
        pending = self.pending(key)
        if pending is not None:
            if generation < pending[1]:
                self.c["stale_rejected"] += 1
                return ("Stale", pending[1])
            if generation == pending[1]:
                if tuple(a[1] for a in pending[2]) != tuple(counts):
                    raise Error("GenerationConflict", key=key, generation=generation)
                return ("DuplicatePending", pending)
            self.cancel_pending(key, pending[1])
        if sum(1 for s in self.pages.values() if s[1] is not None) >= self.max_pending:
            self.c["backpressured"] += 1
            raise Error("PendingCapacity", maximum=self.max_pending)
        v = self.v.reserve(counts[0])
        if v is None:
            self.c["backpressured"] += 1
            raise Error("ArenaCapacity", capacity="Vertices")
        i = self.i.reserve(counts[1])
        if i is None:
            self.v.release(v)
            self.c["backpressured"] += 1
            raise Error("ArenaCapacity", capacity="Indices")
        m = self.m.reserve(counts[2])
        if m is None:
            self.i.release(i)
            self.v.release(v)
            self.c["backpressured"] += 1
            raise Error("ArenaCapacity", capacity="Meshlets")
        reservation = (key, generation, (v, i, m))
        self.pages.setdefault(key, [None, None])[1] = reservation
        self.c["reservations"] += 1
        now = self.counters()
        self.c["pending_high_water"] = max(self.c["pending_high_water"], now["pending_pages"])
        self.c["vertex_high_water"] = max(self.c["vertex_high_water"], now["used_vertices"])
        self.c["index_high_water"] = max(self.c["index_high_water"], now["used_indices"])
        self.c["meshlet_high_water"] = max(self.c["meshlet_high_water"], now["used_meshlets"])
        return ("Reserved", reservation)

    def publish(self, reservation):
        key = reservation[0]
        if key not in self.pages:
            raise Error("ReservationMissing", key=key)
        state = self.pages[key]
        generations = [g for g in (state[0][0] if state[0] else None, state[1][1] if state[1] else None) if g is not None]
        if generations and reservation[1] < max(generations):
            self.c["stale_rejected"] += 1
            return ("Stale", max(generations))
        if state[1] != reservation:
            raise Error("ReservationMismatch", key=key, generation=reservation[1])
        state[1] = None
        current = (reservation[1], reservation[2])
        replaced, state[0] = state[0], current
        if replaced is not None:
            self._release(replaced[1])
            self.c["replacements"] += 1
        self.c["publications"] += 1
        return ("Published", current, replaced)

    def cancel_pending(self, key, generation):
        state = self.pages.get(key)
        if state is None or state[1] is None:
            return False
        if state[1][1] != generation:
            raise Error("ReservationMismatch", key=key, generation=generation)
        self._release(state[1][2])
        state[1] = None
        self.c["cancellations"] += 1
        if state[0] is None:
            del self.pages[key]
        return True

    def evict(self, key, generation):
        state = self.pages.get(key)
        if state is None:
            return ("Missing", None)
        newest = max([state[0][0] if state[0] else 0, state[1][1] if state[1] else 0])
        if generation < newest:
            self.c["stale_rejected"] += 1
            return ("Stale", newest)
        del self.pages[key]
        if state[0]:
            self._release(state[0][1])
        if state[1]:
            self._release(state[1][2])
        self.c["evictions"] += 1
        return ("Evicted", None)
