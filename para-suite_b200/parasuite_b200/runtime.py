"""Context + batch-level calls over the C ABI (host side of the drop-in boundary).

`Context` owns one GPU (one process per GPU under torch.distributed).  All compute goes through
libparasuite_b200.so; there is no CPU path here.  torch is used only as plumbing: device memory for
resident batches, streams, and the NCCL all-reduce of the count vector.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import abi
from .batch import PackedReference, ReadBatch


def _check(lib, ctx, st, fault=None):
    if st == abi.PS_OK:
        return
    msg = lib.ps_last_error(ctx).decode() if ctx else lib.ps_strerror(st).decode()
    if st == abi.PS_ERR_REFERENCE_WOULD_THROW:
        raise abi.ReferenceWouldThrow(st, msg, fault)
    raise abi.PsError(st, msg or lib.ps_strerror(st).decode())


class DeviceBatch:
    """A ReadBatch resident in HBM (torch tensors hold the memory; the kernels see raw pointers)."""

    def __init__(self, host: ReadBatch, device):
        import torch
        self.n_reads = host.n_reads
        self.host = host
        self._t = {}
        for f in ReadBatch.FIELDS:
            a = getattr(host, f)
            if a is None:
                self._t[f] = None
                continue
            t = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8))
            self._t[f] = t.to(device, non_blocking=False)
        s = host.as_struct()
        for f in ReadBatch.FIELDS:
            setattr(s, f, None if self._t[f] is None else self._t[f].data_ptr())
        self.struct = s
        self.nbytes = sum(t.numel() for t in self._t.values() if t is not None)


class PinnedBatch:
    """A ReadBatch copied into page-locked host memory (what the native batcher fills)."""

    def __init__(self, host: ReadBatch):
        import torch
        self.n_reads = host.n_reads
        self._t = {}
        s = host.as_struct()
        for f in ReadBatch.FIELDS:
            a = getattr(host, f)
            if a is None:
                continue
            t = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8)).clone().pin_memory()
            self._t[f] = t
            setattr(s, f, t.data_ptr())
        self.struct = s
        self.h2d_bytes = (host.n_reads * 8 + host.bases_bytes + host.qual_bytes + host.cigar_count * 4 +
                          host.exc_count * 4 + (host.n_tiles + 1) * 28)


class Context:
    def __init__(self, device: int = 0):
        self.lib = abi.load_library()
        h = C.c_void_p()
        st = self.lib.ps_create(C.byref(h), device)
        if st != abi.PS_OK:
            raise abi.PsError(st, self.lib.ps_strerror(st).decode())
        self.h = h
        self.device = device
        self.ref: Optional[PackedReference] = None
        self._max_len = 0
        self._infer_q = False
        self._keep = []

    def close(self):
        if self.h:
            self.lib.ps_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- reference ------------------------------------------------------------------------------
    def upload_reference(self, ref: PackedReference):
        s = ref.as_struct()
        _check(self.lib, self.h, self.lib.ps_reference_upload(self.h, C.byref(s)))
        self.ref = ref

    # ---- error profile (ErrorProfiling.java:146-409) ---------------------------------------------
    def profile_begin(self, max_read_length: int, infer_qualities: bool = False):
        o = abi.ps_profile_opts(max_read_length, int(infer_qualities))
        _check(self.lib, self.h, self.lib.ps_profile_begin(self.h, C.byref(o)))
        self._max_len = max_read_length
        self._infer_q = infer_qualities

    def profile_batch(self, batch):
        """Host-resident batch (ReadBatch or PinnedBatch): H2D inside the call."""
        s = batch.struct if hasattr(batch, "struct") else batch.as_struct()
        self._keep.append(batch)
        _check(self.lib, self.h, self.lib.ps_profile_batch(self.h, C.byref(s)))

    def profile_batch_device(self, dbatch: DeviceBatch, stream: int = 0):
        _check(self.lib, self.h, self.lib.ps_profile_batch_device(self.h, C.byref(dbatch.struct), stream or None))

    def profile_acc_tensor(self):
        """The int64 accumulator vector as a torch tensor aliasing library memory (for dist.all_reduce)."""
        import torch
        p = C.c_void_p()
        n = C.c_size_t()
        _check(self.lib, self.h, self.lib.ps_profile_acc_device(self.h, C.byref(p), C.byref(n)))

        class _Alias:
            __cuda_array_interface__ = {"shape": (n.value,), "typestr": "<i8", "data": (p.value, False), "version": 3}

        return torch.as_tensor(_Alias(), device=f"cuda:{self.device}")

    def profile_end(self) -> dict:
        m = self._max_len
        out = {
            "position_conversions": np.zeros((m, 4, 4), dtype=np.int32),
            "quality_per_mismatch": np.zeros((4, 4), dtype=np.int32),
            "quality_per_mismatch_counts": np.zeros((4, 4), dtype=np.int32),
            "insertions_per_pos": np.zeros(m, dtype=np.float64),
            "deletions_per_pos": np.zeros(m, dtype=np.float64),
            "counters": np.zeros(abi.PS_PC_COUNT, dtype=np.int32),
            "wide": np.zeros(self.lib.ps_profile_acc_len(m, int(self._infer_q)), dtype=np.int64),
        }
        if self._infer_q:
            out["quality_hist"] = np.zeros((m, 256), dtype=np.int64)
        r = abi.ps_profile_result()
        r.position_conversions = out["position_conversions"].ctypes.data
        r.quality_per_mismatch = out["quality_per_mismatch"].ctypes.data
        r.quality_per_mismatch_counts = out["quality_per_mismatch_counts"].ctypes.data
        r.insertions_per_pos = out["insertions_per_pos"].ctypes.data
        r.deletions_per_pos = out["deletions_per_pos"].ctypes.data
        r.counters = out["counters"].ctypes.data
        r.quality_hist = out["quality_hist"].ctypes.data if self._infer_q else None
        r.wide = out["wide"].ctypes.data
        st = self.lib.ps_profile_end(self.h, C.byref(r))
        self._keep.clear()
        _check(self.lib, self.h, st, fault=(r.fault.code, r.fault.read_ordinal))
        return out

    def profile(self, batch: ReadBatch, max_read_length: int, infer_qualities: bool = False) -> dict:
        self.profile_begin(max_read_length, infer_qualities)
        self.profile_batch(batch)
        return self.profile_end()

    # ---- instrumentation -------------------------------------------------------------------------
    def kernel_launches(self) -> int:
        return int(self.lib.ps_kernel_launches(self.h))

    def kernel_times_reset(self, enabled: bool = True):
        self.lib.ps_kernel_times_reset(self.h, int(enabled))

    def kernel_times_ms(self) -> np.ndarray:
        buf = np.zeros(512, dtype=np.float32)
        n = self.lib.ps_kernel_times(self.h, buf.ctypes.data, buf.size)
        return buf[:max(n, 0)].copy()
