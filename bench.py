#!/usr/bin/env python3
"""bench.py -- Transvoxel chunk-extraction throughput on B200 (BASELINE.json metric).

Workload (BASELINE configs[1]): a 16x16x16 grid of 64^3-cell chunks of procedural
fBm terrain (helio-pass-sdf `terrain_sdf`, TerrainConfig::rolling(), 0.1 m voxels), 4096 chunks,
4.71 GB of CellWord samples.  One "step" = one regular-cell extraction pass over the whole batch.

  value        cells/s with the samples already resident in HBM (CUDA events on the launch stream)
  e2e          the same metric through ONE C-ABI call with HOST buffers (hvx_extract_regular_to_host): pinned
               host samples -> H2D in sub-batches beside the kernels -> packed mesh + counters -> D2H behind
               them.  h2d_gbs_peak is the raw pinned upload bandwidth of this box measured in the same run
               (all ranks at once), so h2d_frac says how much of the link the call uses.
  e2e_variants fill_extract_readback (procedural density on the device, no sample upload), sparse_upload
               (chunks the producer knows to be empty are flagged HVX_CHUNK_UNIFORM and not uploaded),
               unpipelined (round 1's path: monolithic copy, kernel, read-back, counters)
  configs      BASELINE configs[2..4] as sub-records: the planet-scale page set (STRONG scaling at the launched
               N), the three-level LOD seam path, the incremental-edit latency path, the whole per-page pass
  roofline     algorithmic bytes (4*(E+2)^3 + 32*V + 4*I per chunk) / kernel time vs measured HBM peak
  cpu_baseline the CPU oracle (restating the reference's Rust CPU extractor) on this box's cores,
               on a bounded sample of the same chunks

`--impl reference` times that CPU implementation alone (the reference itself is Rust + wgpu and
cannot be built in this image).  Multi-GPU (`torchrun ... --gpus N`): the chunk
list grows with N (weak scaling), is partitioned by the LPT scheduler, no data-path collective.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

EDGE = 64
GRID = 16                      # chunks per axis at N=1
FIELD_TERRAIN_FBM = 16
MAX_VERTICES, MAX_INDICES = 49_152, 73_728   # per-chunk slot capacity (surface chunks need ~25k / ~37k)


WORKLOAD = "terrain"   # --workload: "terrain" (headline), "surface" (every chunk crosses the surface), "empty"


def chunk_grid(n_gpus: int) -> np.ndarray:
    """Global chunk list: page_xyz in [-8, 8)^3 per GPU, extended along +x for N > 1 (x fastest)."""
    xs = np.arange(-GRID // 2, -GRID // 2 + GRID * n_gpus, dtype=np.int64)
    ys = np.arange(-GRID // 2, GRID // 2, dtype=np.int64)
    zs = np.arange(-GRID // 2, GRID // 2, dtype=np.int64)
    if WORKLOAD == "surface":      # diagnostics: 64 x 1 x 64 chunks, all on the surface layer y = -1
        xs = np.arange(-32, -32 + 64 * n_gpus, dtype=np.int64)
        ys = np.array([-1], dtype=np.int64)
        zs = np.arange(-32, 32, dtype=np.int64)
    elif WORKLOAD == "empty":      # diagnostics: nothing but air
        ys = ys + 16
    z, y, x = np.meshgrid(zs, ys, xs, indexing="ij")
    return np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)


def host_threads() -> int:
    """Every host thread this process may use (torchrun exports OMP_NUM_THREADS=1, so ask the OS)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def bind_to_gpu_numa_node(local_rank):
    """Run this rank on the CPUs of the NUMA node its GPU hangs off, so the pinned host buffers of the e2e leg
    are first-touched in memory local to that GPU's PCIe root (matters when 8 ranks upload at once)."""
    try:
        import torch
        props = torch.cuda.get_device_properties(local_rank)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        node = int(Path(f"/sys/bus/pci/devices/{bdf}/numa_node").read_text())
        if node < 0:
            return None
        cpus = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def profiled_traffic(alg_bytes):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch.  ncu cannot run inside the timed process, so this
    is an OFFLINE number: the committed ncu --set full capture of this same command (profiles/r02_traffic.json, or
    round 1's), reported only when it was taken on exactly this workload (same algorithmic bytes).  -> (bytes, source)"""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        path = ROOT / "profiles" / name
        if path.exists():
            rec = json.loads(path.read_text())
            if rec.get("algorithmic_bytes_per_launch") == alg_bytes:
                return rec["dram_bytes_per_launch"], f"offline ncu --set full capture, profiles/{name}"
    return None, None


def measured_peak_gbs():
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        return float(json.loads(path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.device_index = device_index
        self.lines = []
        self.proc = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.device_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            self.proc.wait()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, value in zip(names, parts[4:8]):
                if value.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def run_reference(args):
    """CPU arm: the oracle's restatement of the reference CPU extractor, all host threads."""
    from oracle import oracle as O
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    pages = chunk_grid(1)
    sample = pages[args.cpu_offset::args.cpu_stride]          # bounded, y-uniform sample of the workload
    words = (EDGE + 2) ** 3
    samples = O.batch_fill(FIELD_TERRAIN_FBM, sample, EDGE, threads=threads)   # setup (untimed)
    for _ in range(args.warmup):
        O.batch_regular(FIELD_TERRAIN_FBM, sample[:threads], EDGE, threads=threads, do_fill=False,
                        samples=samples[:threads * words])
    t0 = time.perf_counter()
    cells = 0
    for _ in range(args.steps):
        c, totals = O.batch_regular(FIELD_TERRAIN_FBM, sample, EDGE, threads=threads, do_fill=False, samples=samples)
        cells += c
    dt = time.perf_counter() - t0
    value = cells / dt
    desc = f"{len(sample)} of 4096 chunks (every {args.cpu_stride}th, all y layers), samples resident in host RAM"
    print(json.dumps({
        "impl": "reference", "metric": "voxel_cells_per_sec", "value": value, "unit": "cells/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i16/f32", "data": "synthetic",
        "chunks_per_s": value / EDGE ** 3,
        "config": {"workload": "fbm_terrain_4096x64^3", "edge": EDGE, "chunks_per_step": len(sample),
                   "implementation": "C restatement of the reference's Rust CPU extractor (oracle/), OpenMP over chunks"},
        "cpu_baseline": {"value": value, "unit": "cells/s", "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_ours(args):
    import torch
    import torch.distributed as dist

    import helio_b200 as H
    from helio_b200 import _ffi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        with quiet_stdout():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            torch.cuda.set_device(local_rank)
            dist.barrier()          # creates the communicator now, with stdout parked
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 and not args.no_numa else None
    print(f"[bench] rank {rank}: GPU {local_rank} numa node {numa_node}, {len(os.sched_getaffinity(0))} cpus", file=sys.stderr)

    # ---- scheduler: global chunk list -> this rank's shard (no data-path collective) ----------
    pages_all = chunk_grid(world)
    if args.chunks:
        pages_all = pages_all[:args.chunks * world]
    costs = np.full(len(pages_all), H.chunk_cost(EDGE, 0), dtype=np.uint64)
    owner = H.partition_chunks(costs, world)
    pages = np.ascontiguousarray(pages_all[owner == rank])
    n = len(pages)
    words = (EDGE + 2) ** 3

    batch = H.ChunkBatchExtractor(local_rank, edge=EDGE, max_chunks=n, max_vertices=MAX_VERTICES,
                                  max_indices=MAX_INDICES)
    ctx = batch.ctx
    if args.spread:
        ctx.debug_set_mode(args.spread << 12)
    stream = torch.cuda.Stream(device=device)
    ctx.set_stream(stream.cuda_stream)
    descs = H.make_descs(n)

    # ---- setup (untimed): procedural density on the device, K1 --------------------------------
    t_fill0, t_fill1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ctx.fill_density(FIELD_TERRAIN_FBM, pages)          # warm
        t_fill0.record(stream)
        ctx.fill_density(FIELD_TERRAIN_FBM, pages)
        t_fill1.record(stream)
    stream.synchronize()
    fill_ms = t_fill0.elapsed_time(t_fill1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    # ---- device-resident extraction: warm-up, then K timed steps -------------------------------
    for _ in range(args.warmup):
        ctx.extract_regular(None, descs, n)
    if args.warmup and not args.no_hints:
        # steady state of a renderer that re-extracts resident chunks: every chunk's vertex count of the previous
        # pass goes into its descriptor as the scheduler's cost hint, so the batch starts its heaviest chunks first
        prev = batch.counters(n)["required_vertices"]
        descs = H.make_descs(n, cost_hint=[int(v) for v in prev])
        ctx.extract_regular(None, descs, n)
    barrier()
    launches0 = ctx.launch_count
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    clocks = ClockSampler(local_rank)       # sampled from here to the end of the e2e legs (stopped before the CPU baseline)
    clocks.__enter__()
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        starts[k].record(stream)
        ctx.extract_regular(None, descs, n)
        stops[k].record(stream)
    barrier()
    wall = time.perf_counter() - t0
    launches = ctx.launch_count - launches0
    kernel_ms = [s.elapsed_time(e) for s, e in zip(starts, stops)]
    total_ms = starts[0].elapsed_time(stops[-1])
    if world > 1:
        t = torch.tensor([total_ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    counters = batch.counters(n)
    total_v = int(counters["emitted_vertices"].astype(np.int64).sum())
    total_i = int(counters["emitted_indices"].astype(np.int64).sum())
    overflow = int((counters["vertex_overflow"] | counters["index_overflow"]).sum())
    surface_chunks = int((counters["emitted_vertices"] > 0).sum())
    cells_per_step = n * EDGE ** 3
    value = world * cells_per_step * args.steps / (total_ms * 1e-3)

    alg_bytes = n * words * 4 + 32 * total_v + 4 * total_i
    avg_kernel_s = float(np.mean(kernel_ms)) * 1e-3
    peak, peak_src = measured_peak_gbs()
    achieved = alg_bytes / avg_kernel_s / 1e9

    # ---- e2e through the C ABI with HOST buffers ------------------------------------------------
    e2e, e2e_variants = None, {}
    if not args.no_e2e:
        import ctypes as C
        host_samples = torch.empty(n * words, dtype=torch.int32, pin_memory=True)
        ctx.synchronize()
        # one-time (untimed) copy of the generated samples to the pinned host buffer
        rc = _ffi.load().hvx_read(ctx._handle, _ffi.BUF_SAMPLES, 0, n * words * 4, C.c_void_p(host_samples.data_ptr()))
        assert rc == 0
        host_v = torch.empty((total_v + 1024) * 8, dtype=torch.int32, pin_memory=True)
        host_i = torch.empty(total_i + 1024, dtype=torch.int32, pin_memory=True)
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        h2d_bytes = n * words * 4 + n * 32
        d2h_bytes = 32 * total_v + 4 * total_i + n * 16 + n * 32

        def timed_e2e(step, warm=2):
            """wall clock over e2e_steps calls, barrier + device sync on both sides, max over ranks"""
            for _ in range(warm):
                step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                out = step()
            barrier()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], device=device, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            return dt / e2e_steps, out

        def check(out):
            tv, ti, _ranges, c = out
            assert tv == total_v and ti == total_i and int(c["emitted_vertices"].astype(np.int64).sum()) == total_v

        # the raw link: the same 4.71 GB pinned buffer into the sample arena with plain async copies, all ranks at once

        def raw_upload():
            rc = _ffi.load().hvx_write(ctx._handle, _ffi.BUF_SAMPLES, 0, n * words * 4, C.c_void_p(host_samples.data_ptr()))
            assert rc == 0
        raw_s, _ = timed_e2e(raw_upload, warm=1)
        h2d_peak = n * words * 4 / raw_s / 1e9

        # headline: ONE call, pipelined upload -> extraction -> packed read-back
        step_s, out = timed_e2e(lambda: ctx.extract_regular_to_host(host_samples, descs, n, host_v, host_i))
        check(out)
        e2e = {"value": world * cells_per_step / step_s, "unit": "cells/s", "h2d_bytes_per_step": h2d_bytes,
               "d2h_bytes_per_step": d2h_bytes, "ms_per_step": step_s * 1e3, "steps": e2e_steps,
               "path": "hvx_extract_regular_to_host(pinned host samples): sub-batched H2D beside the kernels, packed meshes + counters D2H behind them",
               "h2d_gbs_achieved": h2d_bytes / step_s / 1e9, "h2d_gbs_peak": h2d_peak,
               "h2d_frac": (h2d_bytes / step_s / 1e9) / h2d_peak,
               "h2d_peak_how": f"hvx_write of the same {n * words * 4 / 1e9:.2f} GB pinned buffer, {world} rank(s) at once, same run",
               "numa_node": numa_node}

        # round 1's path for comparison: monolithic copy, kernel, read-back, counters
        def unpipelined():
            raw_upload()
            ctx.extract_regular(None, descs, n)
            tv, ti, r = ctx.read_meshes(0, 0, n, vertices_out=host_v, indices_out=host_i)
            return tv, ti, r, batch.counters(n)
        step_s, out = timed_e2e(unpipelined, warm=1)
        check(out)
        e2e_variants["unpipelined"] = {"value": world * cells_per_step / step_s, "unit": "cells/s", "ms_per_step": step_s * 1e3,
                                       "path": "hvx_write(samples) + hvx_extract_regular + hvx_read_meshes + counters (round 1's sequence)"}

        # production flow 1: procedural density on the device -> extraction -> packed read-back (no sample upload)
        def fill_extract_readback():
            ctx.fill_density(FIELD_TERRAIN_FBM, pages)
            return ctx.extract_regular_to_host(None, descs, n, host_v, host_i)
        step_s, out = timed_e2e(fill_extract_readback, warm=1)
        check(out)
        e2e_variants["fill_extract_readback"] = {
            "value": world * cells_per_step / step_s, "unit": "cells/s", "ms_per_step": step_s * 1e3,
            "h2d_bytes_per_step": n * 24 + n * 32, "d2h_bytes_per_step": d2h_bytes,
            "path": "hvx_fill_density(page list) + hvx_extract_regular_to_host(arena): host sends page coordinates, gets packed meshes"}

        # sparse upload: the producer (an octree that keeps min / max density per node) knows which chunks hold no
        # surface; they are flagged HVX_CHUNK_UNIFORM and neither uploaded nor walked
        empty = counters["required_vertices"] == 0
        sparse = H.make_descs(n, cost_hint=[int(v) for v in counters["required_vertices"]],
                              flags=[_ffi.HVX_CHUNK_UNIFORM if e else 0 for e in empty])
        step_s, out = timed_e2e(lambda: ctx.extract_regular_to_host(host_samples, sparse, n, host_v, host_i), warm=1)
        check(out)
        e2e_variants["sparse_upload"] = {
            "value": world * cells_per_step / step_s, "unit": "cells/s", "ms_per_step": step_s * 1e3,
            "h2d_bytes_per_step": int((~empty).sum()) * words * 4 + n * 32, "d2h_bytes_per_step": d2h_bytes,
            "uploaded_chunks": int((~empty).sum()),
            "path": "hvx_extract_regular_to_host with HVX_CHUNK_UNIFORM on the chunks that produced no vertex on the previous pass"}
        ctx.extract_regular(host_samples, descs, n)      # leave the arena whole again
        ctx.synchronize()
        del host_samples, host_v, host_i

    clocks.__exit__(None, None, None)

    # ---- CPU baseline on this box's cores (rank 0, N=1 only) ------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        from oracle import oracle as O
        threads = host_threads()
        idx = np.arange(args.cpu_offset, n, args.cpu_stride)
        sample_pages = pages[idx]
        host = np.empty(len(idx) * words, dtype=np.uint32)
        for j, i in enumerate(idx):
            host[j * words:(j + 1) * words] = ctx.read(_ffi.BUF_SAMPLES, int(i) * words, words)
        t0 = time.perf_counter()
        cells, totals = O.batch_regular(FIELD_TERRAIN_FBM, sample_pages, EDGE, threads=threads, do_fill=False, samples=host)
        dt = time.perf_counter() - t0
        # the sample's GPU result must equal the CPU result (counts; full parity lives in tests/)
        assert totals[0] == int(counters["required_vertices"][idx].astype(np.int64).sum()), "CPU/GPU vertex totals differ"
        cpu = {"value": cells / dt, "unit": "cells/s", "cores": threads, "kind": "port",
               "sample": f"{len(idx)} of {n} chunks (every {args.cpu_stride}th), {dt:.2f} s wall, samples resident in host RAM"}
        del host

    kernel_name = ctx.regular_kernel_name()
    batch.close()

    # ---- the other BASELINE configurations, as sub-records --------------------------------------
    configs = {}
    if not args.no_configs and WORKLOAD == "terrain":
        sys.path.insert(0, str(ROOT / "tools"))
        import bench_cases as BC
        configs["planet"] = BC.planet_case(H, torch, dist, device, local_rank, rank, world, steps=args.steps, warmup=args.warmup)
        if world == 1:
            configs["lod_seam"] = BC.lod_seam_case(H, torch, device, local_rank)
            configs["edit_latency"] = BC.edit_latency_case(H, torch, device, local_rank)
            configs["page_pass"] = BC.page_pass_case(H, torch, device, local_rank)

    if rank == 0:
        clk = clocks.summary()
        traffic, traffic_src = profiled_traffic(alg_bytes) if WORKLOAD == "terrain" and world == 1 else (None, None)
        print(json.dumps({
            "metric": "voxel_cells_per_sec", "value": value, "unit": "cells/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i16/f32",
            "data": "synthetic", "chunks_per_s": value / EDGE ** 3,
            "config": {"workload": "fbm_terrain_4096x64^3" if WORKLOAD == "terrain" else f"diagnostic_{WORKLOAD}_4096x64^3", "edge": EDGE, "chunks_per_gpu": n,
                       "grid": f"{GRID * world}x{GRID}x{GRID}", "surface_chunks_per_gpu": surface_chunks,
                       "vertices_per_step_per_gpu": total_v, "indices_per_step_per_gpu": total_i,
                       "l2_policy": "inputs larger than L2 (4.71 GB samples per GPU vs 126 MB L2)",
                       "slot_capacity": [MAX_VERTICES, MAX_INDICES], "overflowed_chunks": overflow,
                       "partition": "LPT static, no data-path collective",
                       "chunk_order": "identity" if args.no_hints or not args.warmup else "by the previous pass's vertex counts (hvx_chunk_desc.cost_hint): heavy chunks heaviest first, spread over the first 75 % of the start order" + (f" (--spread {args.spread})" if args.spread else "")},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "kernel": kernel_name, "kernel_ms": float(np.mean(kernel_ms)),
                         "frac_of_8TBps_nominal": achieved / 8000.0},
            "fill_kernel": {"ms": fill_ms, "GB/s": n * words * 4 / (fill_ms * 1e-3) / 1e9},
            "cpu_baseline": cpu, "e2e": e2e, "e2e_variants": e2e_variants, "configs": configs,
            "gpu_launches": launches, "clocks": clk,
            "wall_s_timed_region": wall,
        }))
    if world > 1:
        dist.destroy_process_group()


class quiet_stdout:
    """Send file descriptor 1 to stderr for a while: torch prints "NCCL version ..." on stdout when the
    communicator is created, and stdout is reserved for the one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self._saved, 1)
        os.close(self._saved)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--chunks", type=int, default=0, help="chunks per GPU (default: the full 4096)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the sub-records of the other BASELINE configurations")
    ap.add_argument("--no-hints", action="store_true", help="do not feed the previous pass's vertex counts back as cost hints")
    ap.add_argument("--spread", type=int, default=0, help="diagnostic: share (per cent) of the start order the heavy chunks are spread over; 255 = plain descending order; 0 = library default")
    ap.add_argument("--no-numa", action="store_true", help="do not bind ranks to their GPU's NUMA node (N > 1)")
    ap.add_argument("--cpu-stride", type=int, default=1, help="CPU baseline runs every k-th chunk of the workload (default: all of them)")
    ap.add_argument("--cpu-offset", type=int, default=0)
    ap.add_argument("--workload", choices=["terrain", "surface", "empty"], default="terrain")
    args = ap.parse_args()
    global WORKLOAD
    WORKLOAD = args.workload
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
