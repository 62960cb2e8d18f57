"""Legacy 8^3-brick marching-cubes extraction: the compute half of ``VoxelMeshPass``.

Mirrors crates/passes/3d/helio-pass-voxel-mesh/src/lib.rs (``VoxelMeshPass::{new, mark_dirty,
clear_brick_slot, execute}``, ``ActiveBrickRange``) over ``hvx_brick_*`` (helio_b200/csrc/brick_extract.cu).
The render half (multi_draw_indexed_indirect over the arenas) stays with wgpu.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi
from .context import Context

VOXEL_MESH_MAX_BRICKS = 1024          # lib.rs:27
VOXEL_MESH_MAX_DIRTY = 4096           # lib.rs:28
VOXEL_MESH_BRICK_VOXEL_WORDS = 183    # lib.rs:34, ceil(9*9*9 / 4)
MAX_SURFACE_VERTS_PER_BRICK = 2048    # helio-voxel-core/src/constants.rs:15
MAX_SURFACE_INDICES_PER_BRICK = 2048
PADDED_BRICK_SIZE = 9

BRICK_META_DTYPE = np.dtype([("data_offset", "<u4"), ("occupancy", "<u4")])
DIRTY_BRICK_DTYPE = np.dtype([("brick_slot", "<u4"), ("volume_id", "<u4"), ("_pad", "<u4", (2,)), ("origin_size", "<f4", (4,))])
BRICK_MESHLET_DTYPE = np.dtype([(n, "<u4") for n in ("vertex_offset", "index_offset", "vertex_count", "index_count",
                                                      "brick_index", "volume_id", "_pad0", "_pad1")])
BRICK_DRAW_DTYPE = np.dtype([("index_count", "<u4"), ("instance_count", "<u4"), ("first_index", "<u4"),
                             ("base_vertex", "<i4"), ("first_instance", "<u4")])


def pack_brick(voxels_9x9x9) -> np.ndarray:
    """A padded brick (material bytes indexed [z][y][x], 9 per axis) -> its 183 little-endian words."""
    flat = np.zeros(VOXEL_MESH_BRICK_VOXEL_WORDS * 4, dtype=np.uint8)
    flat[:729] = np.asarray(voxels_9x9x9, dtype=np.uint8).reshape(729)
    return flat.view("<u4").copy()


class ActiveBrickRange:
    """lib.rs:80-115: the highest occupied slot bounds the indirect multi-draw."""

    def __init__(self, capacity):
        self.slots = [False] * capacity
        self._draw_count = 0

    def set(self, brick_slot, active):
        if not 0 <= brick_slot < len(self.slots):
            return False
        self.slots[brick_slot] = bool(active)
        if active:
            self._draw_count = max(self._draw_count, brick_slot + 1)
        elif brick_slot + 1 == self._draw_count:
            occupied = [i for i, s in enumerate(self.slots) if s]
            self._draw_count = occupied[-1] + 1 if occupied else 0
        return True

    def draw_count(self):
        return self._draw_count


class VoxelMeshExtractor:
    """The extraction state of ``VoxelMeshPass``: brick metadata + voxel data in, per-slot vertex / normal /
    index ranges, ``GpuBrickMeshlet`` descriptors and ``DrawIndexedIndirect`` records out."""

    def __init__(self, device=0, max_bricks=VOXEL_MESH_MAX_BRICKS, max_dirty=VOXEL_MESH_MAX_DIRTY, context=None):
        self._own_ctx = context is None
        self.ctx = context or Context(device, edge=32, max_chunks=1, max_vertices=1, max_indices=1)
        self._lib = _ffi.load()
        handle = C.c_void_p()
        self.ctx._check(self._lib.hvx_brick_mesher_create(self.ctx._handle, max_bricks, C.byref(handle)))
        self._handle, self.max_bricks, self.max_dirty = handle, max_bricks, max_dirty
        self.dirty_bricks = []
        self.active_bricks = ActiveBrickRange(max_bricks)
        self._meta = np.zeros(max_bricks, dtype=BRICK_META_DTYPE)
        self._voxels = None

    def close(self):
        if self._handle:
            self._lib.hvx_brick_mesher_destroy(self._handle)
            self._handle = None
            if self._own_ctx:
                self.ctx.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # the reference writes these two buffers with queue.write_buffer (brick_meta_buf / voxel_data_buf accessors)
    def write_brick_meta(self, meta):
        meta = np.ascontiguousarray(meta, dtype=BRICK_META_DTYPE)
        self._meta[:len(meta)] = meta

    def write_voxel_data(self, words):
        """``words``: numpy uint32 array (host) or a CUDA tensor of int32 / uint32 words (used in place)."""
        self._voxels = words if hasattr(words, "data_ptr") else np.ascontiguousarray(words, dtype=np.uint32)

    def mark_dirty(self, brick_slot, volume_id, origin, voxel_size, occupied):
        """lib.rs:559-588: out-of-range slots and list overflow are dropped (the reference logs a warning)."""
        if brick_slot >= self.max_bricks or len(self.dirty_bricks) >= self.max_dirty:
            return False
        self.dirty_bricks.append((brick_slot, volume_id, tuple(origin), float(voxel_size)))
        self.active_bricks.set(brick_slot, occupied)
        return True

    def clear_brick_slot(self, brick_slot):
        """lib.rs:590-615."""
        if not self.active_bricks.set(brick_slot, False):
            return False
        self.dirty_bricks = [d for d in self.dirty_bricks if d[0] != brick_slot]
        self.ctx._check(self._lib.hvx_brick_clear_slot(self._handle, brick_slot))
        return True

    def dirty_array(self):
        arr = np.zeros(len(self.dirty_bricks), dtype=DIRTY_BRICK_DTYPE)
        for i, (slot, volume, origin, size) in enumerate(self.dirty_bricks):
            arr[i]["brick_slot"], arr[i]["volume_id"] = slot, volume
            arr[i]["origin_size"] = (*origin, size)
        return arr

    def extract(self, dirty, meta=None):
        """One launch over an explicit DirtyBrick array (DIRTY_BRICK_DTYPE, host) -- or, for the device-resident
        steady state, CUDA tensors holding the same bytes for both ``dirty`` and ``meta`` (n_dirty = bytes / 32)."""
        if self._voxels is None:
            raise ValueError("write_voxel_data first")
        if hasattr(self._voxels, "data_ptr"):
            ptr, n_words = C.c_void_p(self._voxels.data_ptr()), self._voxels.numel()
        else:
            ptr, n_words = C.c_void_p(self._voxels.ctypes.data), self._voxels.size
        if hasattr(dirty, "data_ptr"):
            d_ptr, n_dirty = C.c_void_p(dirty.data_ptr()), dirty.numel() * dirty.element_size() // 32
            m_ptr, n_meta = C.c_void_p(meta.data_ptr()), meta.numel() * meta.element_size() // 8
        else:
            dirty = np.ascontiguousarray(dirty, dtype=DIRTY_BRICK_DTYPE)
            d_ptr, n_dirty = C.c_void_p(dirty.ctypes.data), len(dirty)
            m_ptr, n_meta = C.c_void_p(self._meta.ctypes.data), self.max_bricks
        self.ctx._check(self._lib.hvx_brick_extract(self._handle, m_ptr, n_meta, ptr, n_words, d_ptr, n_dirty))

    def rejected(self):
        """Dirty entries of device-resident lists the kernel's bounds check skipped (host lists raise instead)."""
        return int(self._read(_ffi.BRICK_REJECTED, np.uint32, 0, 1)[0])

    def execute(self):
        """The compute step of ``VoxelMeshPass::execute`` (lib.rs:653-667); returns ``draw_count``."""
        if self.dirty_bricks:
            self.extract(self.dirty_array())
        self.dirty_bricks.clear()
        return self.active_bricks.draw_count()

    def _read(self, buffer_id, dtype, first, count):
        out = np.empty(count, dtype=dtype)
        if count:
            size = np.dtype(dtype).itemsize
            self.ctx._check(self._lib.hvx_brick_read(self._handle, buffer_id, first * size, count * size,
                                                     out.ctypes.data_as(C.c_void_p)))
        return out

    def descriptors(self, first=0, count=None):
        return self._read(_ffi.BRICK_DESCRIPTORS, BRICK_MESHLET_DTYPE, first, self.max_bricks - first if count is None else count)

    def indirect_draws(self, first=0, count=None):
        return self._read(_ffi.BRICK_DRAWS, BRICK_DRAW_DTYPE, first, self.max_bricks - first if count is None else count)

    def brick_mesh(self, brick_slot, count=None):
        """(vertices [n,4], normals [n,4], indices [n]) of one slot; n = the published vertex count."""
        if count is None:
            count = int(self.descriptors(brick_slot, 1)[0]["vertex_count"])
        base = brick_slot * MAX_SURFACE_VERTS_PER_BRICK
        v = self._read(_ffi.BRICK_VERTICES, np.float32, base * 4, count * 4).reshape(-1, 4)
        n = self._read(_ffi.BRICK_NORMALS, np.float32, base * 4, count * 4).reshape(-1, 4)
        i = self._read(_ffi.BRICK_INDICES, np.uint32, base, count)
        return v, n, i
