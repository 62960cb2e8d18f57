"""Small workload for compute-sanitizer (memcheck / racecheck): every kernel, both edges, a few chunks."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import helio_b200 as H  # noqa: E402

for edge in (32, 64):
    pages = [[0, -1, 0], [0, 0, 0], [1, -1, 1], [-1, -1, 0], [0, 1, 0], [2, -1, -2]]
    n = len(pages)
    b = H.ChunkBatchExtractor(0, edge=edge, max_chunks=n, max_vertices=40000, max_indices=60000,
                              max_transition_vertices=8192, max_transition_indices=24576, debug_records=(edge == 32))
    for kind in (16, 1, 17):
        b.fill_density(kind, pages)
        b.extract_regular(None, n, transition_mask=[0, 0x3F, 1, 2, 0, 0x15], dirty_microbricks=[(1 << 64) - 1] * 5 + [0xFF00FF])
        b.ctx.build_meshlets(n, 0)
    b.fill_slabs(16, pages, [1] * n)
    b.extract_transition(None, n, [0x3F, 0x15, 0, 1, 0x2A, 0x3F])
    b.ctx.build_meshlets(n, 1)
    c = b.counters(n)
    t = b.transition_counters(n)
    print(edge, int(c["emitted_vertices"].sum()), int(t["emitted_vertices"].sum()))
    b.close()
print("sanitize workload done")
