"""One process per GPU: the path's single exchange step.

Each rank probes its own device (independent devices, no data-path
collective); the only exchange is an all-gather of the fixed 512-byte
``cro_probe_result`` structs, over NCCL/NVLink on a GPU box and
over gloo in the CPU tests.  The single-process form of the same step is
``cro_probe_all`` in the C library (what a Go operator calls); this module is
the torchrun form bench.py uses.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional

from . import ProbeResult


class DevBuf:
    """Zero-copy torch view of a device pointer owned by libcroprobe."""

    def __init__(self, ptr: int, nbytes: int) -> None:
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


def result_to_bytes(r: ProbeResult) -> bytes:
    return ctypes.string_at(ctypes.addressof(r), ctypes.sizeof(ProbeResult))


def results_from_bytes(raw: bytes) -> List[ProbeResult]:
    n = len(raw) // ctypes.sizeof(ProbeResult)
    return [ProbeResult.from_buffer_copy(raw, i * ctypes.sizeof(ProbeResult)) for i in range(n)]


def all_gather_results(dist, local: ProbeResult, send=None, recv=None) -> List[ProbeResult]:
    """All-gathers the per-rank result structs; every rank returns the same list, ordered by rank.

    ``send`` / ``recv`` may be preallocated device tensors (``send`` viewing the
    library's own result buffer: the kernel-side struct is the send buffer);
    otherwise CPU tensors are built from ``local`` (gloo)."""
    import torch
    world = dist.get_world_size()
    size = ctypes.sizeof(ProbeResult)
    if send is None:
        send = torch.frombuffer(bytearray(result_to_bytes(local)), dtype=torch.uint8)
    if recv is None:
        recv = torch.empty(world * size, dtype=torch.uint8, device=send.device)
    dist.all_gather_into_tensor(recv, send)
    raw = bytes(recv.cpu().numpy().tobytes())
    return results_from_bytes(raw)


def check_gathered(results: List[ProbeResult], world: int) -> Optional[str]:
    """Sanity of a gathered array: one entry per rank, distinct devices, all probes ok."""
    if len(results) != world:
        return "expected %d results, got %d" % (world, len(results))
    uuids = [r.gpu_uuid for r in results]
    if len(set(uuids)) != len(uuids):
        return "two ranks probed the same device: %r" % uuids
    for i, r in enumerate(results):
        if r.status != 0:
            return "rank %d probe status %d" % (i, r.status)
    return None
