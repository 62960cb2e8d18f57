#!/usr/bin/env python3
"""Golden per-chunk fingerprints of the headline workload (4096 fBm chunks of 64^3, BASELINE configs[1]).

Runs the CPU oracle (oracle/, no GPU, no reference tree) over the bench's 16x16x16 chunk grid and
stores, for every chunk that has a mesh: vertex count, index count, CRC-32 of the vertex bytes and
of the index bytes, plus the batch totals.  tests/test_gpu_fullsize.py replays the same grid on the
GPU and compares every entry -- bit-exact parity at the full benchmark size through hashes.
"""
import json
import sys
import zlib
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import oracle as O  # noqa: E402

EDGE, GRID = 64, 16


def pages():
    xs = np.arange(-GRID // 2, GRID // 2, dtype=np.int64)
    z, y, x = np.meshgrid(xs, xs, xs, indexing="ij")
    return np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)


def one(page):
    s = O.fixture_fill(O.FIELD_TERRAIN_FBM, [int(v) for v in page], edge=EDGE)
    m = O.extract_regular(s, edge=EDGE, debug=False)
    if len(m.vertices) == 0:
        return None
    return [len(m.vertices), len(m.indices), zlib.crc32(m.vertices.tobytes()), zlib.crc32(m.indices.tobytes())]


def main():
    pg = pages()
    with ThreadPoolExecutor(8) as ex:
        res = list(ex.map(one, pg))
    chunks = {str(i): r for i, r in enumerate(res) if r is not None}
    out = {"edge": EDGE, "grid": GRID, "field": "terrain_fbm", "chunks_total": len(pg), "chunks_with_mesh": len(chunks),
           "vertices_total": sum(r[0] for r in chunks.values()), "indices_total": sum(r[1] for r in chunks.values()),
           "chunks": chunks}
    path = ROOT / "tests" / "golden" / "terrain_4096x64.json"
    path.write_text(json.dumps(out, separators=(",", ":")) + "\n")
    print(path, out["chunks_with_mesh"], out["vertices_total"], out["indices_total"])


if __name__ == "__main__":
    main()
