"""Error types mirroring the reference's thiserror enums.

PV/src/transvoxel_gpu.rs:445-459            TransvoxelGpuError
PV/src/transvoxel_transition_gpu.rs:712-732 TransvoxelTransitionGpuError
PV/src/transvoxel_transition.rs:443-453     TransvoxelTransitionError
helio-planet-voxel-core/src/types.rs:356-364 AddressError
PV/src/lod_topology.rs:369-401              TerrainLodTopologyError
"""
from __future__ import annotations

import re

from . import _ffi


class HvxError(RuntimeError):
    """Base class; ``status`` is the negative ``hvx_status`` the C ABI returned."""

    def __init__(self, message, status=None):
        super().__init__(message)
        self.status = status


class TransvoxelGpuError(HvxError):
    pass


class TransvoxelTransitionGpuError(HvxError):
    pass


def _parse_counts(message):
    numbers = [int(v) for v in re.findall(r"\d+", message)]
    return (numbers + [None, None])[:2]


class SampleCount(TransvoxelGpuError):
    def __init__(self, message, status=None):
        super().__init__(message, status)
        self.actual, self.expected = _parse_counts(message)


class InvalidExtractionCapacity(TransvoxelGpuError):
    def __init__(self, message, status=None, max_vertices=None, max_indices=None):
        super().__init__(message, status)
        self.max_vertices, self.max_indices = max_vertices, max_indices


class DeviceLimit(TransvoxelGpuError):
    pass


class TransitionSampleCount(TransvoxelTransitionGpuError):
    def __init__(self, message, status=None):
        super().__init__(message, status)
        self.actual, self.expected = _parse_counts(message)


class TransitionMask(TransvoxelTransitionGpuError):
    def __init__(self, message, status=None, mask=None):
        super().__init__(message, status)
        self.mask = mask


class TransitionInvalidExtractionCapacity(TransvoxelTransitionGpuError):
    def __init__(self, message, status=None, max_vertices=None, max_indices=None):
        super().__init__(message, status)
        self.max_vertices, self.max_indices = max_vertices, max_indices


class TransitionDeviceLimit(TransvoxelTransitionGpuError):
    pass


class FinestLodHasNoFinerNeighbor(HvxError):
    pass


class AddressError(HvxError):
    pass


class BatchCapacity(HvxError):
    pass


class CudaError(HvxError):
    pass


class TerrainLodTopologyError(HvxError):
    """``kind`` is the reference variant name (Empty, DuplicatePage, OverlappingPages, ...)."""

    def __init__(self, message, status=None, kind=None):
        super().__init__(message, status)
        self.kind = kind


_TOPOLOGY_KINDS = {
    _ffi.HVX_E_TOPOLOGY_EMPTY: "Empty", _ffi.HVX_E_TOPOLOGY_DUPLICATE: "DuplicatePage",
    _ffi.HVX_E_TOPOLOGY_OVERLAP: "OverlappingPages", _ffi.HVX_E_TOPOLOGY_UNBALANCED: "UnbalancedFace",
    _ffi.HVX_E_TOPOLOGY_ROOT_LOD: "UnsupportedRootLod", _ffi.HVX_E_TOPOLOGY_MINIMUM_LOD: "UnsupportedMinimumLod",
    _ffi.HVX_E_TOPOLOGY_PAGE_BUDGET: "PageBudget", _ffi.HVX_E_TOPOLOGY_MISSING_PARENT: "MissingRefinementParent",
    _ffi.HVX_E_TOPOLOGY_COVERAGE: "TangentCoverage",
}


def raise_for_status(status, message, kind="regular", **extra):
    transition = kind == "transition"
    if status == _ffi.HVX_E_SAMPLE_COUNT:
        raise (TransitionSampleCount if transition else SampleCount)(message, status)
    if status == _ffi.HVX_E_INVALID_CAPACITY:
        cls = TransitionInvalidExtractionCapacity if transition else InvalidExtractionCapacity
        raise cls(message, status, extra.get("max_vertices"), extra.get("max_indices"))
    if status == _ffi.HVX_E_DEVICE_LIMIT:
        raise (TransitionDeviceLimit if transition else DeviceLimit)(message, status)
    if status == _ffi.HVX_E_TRANSITION_MASK:
        found = re.search(r"0x[0-9a-fA-F]+|\d+", message)
        raise TransitionMask(message, status, int(found.group(0), 0) if found else None)
    if status == _ffi.HVX_E_FINEST_LOD:
        raise FinestLodHasNoFinerNeighbor(message, status)
    if status == _ffi.HVX_E_ADDRESS:
        raise AddressError(message or "planetary page coordinate arithmetic overflowed", status)
    if status == _ffi.HVX_E_BATCH_CAPACITY:
        raise BatchCapacity(message, status)
    if status == _ffi.HVX_E_CUDA:
        raise CudaError(message, status)
    if status in _TOPOLOGY_KINDS:
        raise TerrainLodTopologyError(message or _TOPOLOGY_KINDS[status], status, _TOPOLOGY_KINDS[status])
    raise HvxError(message or f"hvx status {status}", status)
