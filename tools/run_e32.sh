for v in 0; do timeout 120 python - <<'PY'
import os, sys, numpy as np, torch
sys.path.insert(0, '.')
import helio_b200 as H
xs = np.arange(-8, 8); z, y, x = np.meshgrid(xs, xs, xs, indexing='ij')
pages = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1).astype(np.int64)
b = H.ChunkBatchExtractor(0, edge=32, max_chunks=len(pages), max_vertices=12288, max_indices=18432)
s = torch.cuda.Stream(); b.ctx.set_stream(s.cuda_stream)
b.fill_density(16, pages); d = H.make_descs(len(pages))
for _ in range(3): b.ctx.extract_regular(None, d, len(pages))
s.synchronize(); ts = []
for _ in range(10):
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s); b.ctx.extract_regular(None, d, len(pages)); e.record(s); e.synchronize(); ts.append(a.elapsed_time(e))
c = b.counters(len(pages)); v = int(c['emitted_vertices'].astype(np.int64).sum()); i = int(c['emitted_indices'].astype(np.int64).sum())
t = float(np.median(ts)); nb = len(pages) * 34**3 * 4 + 32 * v + 4 * i
print('e32 variant', os.environ.get('HVX_REGULAR_VARIANT'), round(t, 4), 'ms', round(nb / t / 1e6), 'GB/s', v)
PY
done
