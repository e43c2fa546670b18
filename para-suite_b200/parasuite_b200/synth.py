"""Seeded synthetic workloads -- binding of lib/libps_synth.so (csrc/synth.cpp)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import abi
from .batch import PackedReference, ReadBatch

_lib = None


class ps_synth_params(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("n_reads", C.c_uint64), ("read_len", C.c_uint32), ("mode", C.c_uint32),
                ("threads", C.c_uint32), ("special_ppm", C.c_uint32), ("region_lo", C.c_uint64),
                ("region_hi", C.c_uint64), ("n_ppm", C.c_uint32), ("reserved", C.c_uint32)]


def _load():
    global _lib
    if _lib is None:
        p = os.path.join(abi.LIB_DIR, "libps_synth.so")
        if not os.path.exists(p):
            raise abi.NativeLibraryMissing(f"{p} not built: run `python __graft_entry__.py build`")
        lib = C.CDLL(p)
        lib.ps_synth_reference.restype = C.c_int
        lib.ps_synth_reference.argtypes = [C.c_uint64, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                           C.c_int]
        lib.ps_synth_reads.restype = C.c_void_p
        lib.ps_synth_reads.argtypes = [C.POINTER(ps_synth_params), C.POINTER(abi.ps_reference)]
        lib.ps_synth_batch.restype = C.POINTER(abi.ps_read_batch)
        lib.ps_synth_batch.argtypes = [C.c_void_p]
        lib.ps_synth_n_clusters.restype = C.c_uint64
        lib.ps_synth_n_clusters.argtypes = [C.c_void_p]
        lib.ps_synth_free.argtypes = [C.c_void_p]
        _lib = lib
    return _lib


def default_threads() -> int:
    return max(1, min(32, os.cpu_count() or 1))


# GRCh38 primary assembly lengths (chr1..22, X, Y, M): 3.1 Gb, config 3/4
GRCH38_LENGTHS = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717,
                  133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285,
                  58617616, 64444167, 46709983, 50818468, 156040895, 57227415, 16569]
GRCH38_NAMES = [f"chr{i}" for i in range(1, 23)] + ["chrX", "chrY", "chrM"]


def synth_reference(seed: int, contig_lengths, names=None, n_run: int = 10000, threads: int = 0) -> PackedReference:
    lib = _load()
    lens = np.asarray(contig_lengths, dtype=np.uint64)
    n = int(lens.sum())
    w2, w1 = PackedReference.words_for(n)
    seq2 = np.zeros(w2, dtype=np.uint32)
    inv = np.zeros(w1, dtype=np.uint32)
    lib.ps_synth_reference(seed, len(lens), lens.ctypes.data, n_run, seq2.ctypes.data, inv.ctypes.data,
                           threads or default_threads())
    if names is None:
        names = [f"chr{i + 1}" for i in range(len(lens))]
    return PackedReference(list(names), [int(x) for x in lens], seq2, inv)


class _SynthHandle:
    def __init__(self, h):
        self.h = h

    def __del__(self):
        if self.h and _lib is not None:
            _lib.ps_synth_free(self.h)
            self.h = None


def _view(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    ct = {np.uint8: C.c_uint8, np.uint32: C.c_uint32, np.uint64: C.c_uint64}[dtype]
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,))


def synth_reads(ref: PackedReference, n_reads: int, read_len: int, seed: int = 0x5EED0002, mode: int = 0,
                special_ppm: int = 0, n_ppm: int = 1000, region=None, threads: int = 0) -> ReadBatch:
    """Generate a coordinate-sorted SoA batch.  Arrays are views into generator-owned memory (kept alive)."""
    lib = _load()
    P = ps_synth_params(seed, n_reads, read_len, mode, threads or default_threads(), special_ppm,
                        region[0] if region else 0, region[1] if region else 0, n_ppm, 0)
    rs = ref.as_struct()
    h = lib.ps_synth_reads(C.byref(P), C.byref(rs))
    if not h:
        raise ValueError("synthetic workload parameters out of range (region too small for the read count?)")
    b = lib.ps_synth_batch(h).contents
    n = int(b.n_reads)
    nt = (n + abi.PS_TILE_READS - 1) // abi.PS_TILE_READS
    batch = ReadBatch(
        n, _view(b.meta, n, np.uint32), _view(b.ref_start, n, np.uint32),
        _view(b.bases2, int(b.bases_bytes) + 64, np.uint8), _view(b.qual, int(b.qual_bytes) + 64, np.uint8),
        _view(b.cigar, int(b.cigar_count) + 16, np.uint32), _view(b.tile_base_off, nt + 1, np.uint64),
        _view(b.tile_qual_off, nt + 1, np.uint64), _view(b.tile_cigar_off, nt + 1, np.uint64),
        _view(b.tile_exc_off, nt + 1, np.uint32), _view(b.exc, int(b.exc_count) + 16, np.uint32),
        uniform_len=int(b.uniform_len), uniform_ncigar=int(b.uniform_ncigar), bases_bytes=int(b.bases_bytes),
        qual_bytes=int(b.qual_bytes), cigar_count=int(b.cigar_count), exc_count=int(b.exc_count))
    batch._owner = _SynthHandle(h)
    batch.n_clusters_generated = int(lib.ps_synth_n_clusters(h))
    return batch


def bridge_cluster(ref: PackedReference, start_global: int, count: int = 256, read_len: int = 36) -> ReadBatch:
    """`count` identical plus-strand reads `read_len`M at global offset `start_global`: the reference window with a C at
    the first T.  bench.py puts such a cluster on either side of a mid-contig shard cut, so that one cluster spans the
    cut (open cluster of shard s-1 + head partial of shard s) and the halo merge has real work.  No N calls."""
    L, T = read_len, abi.PS_TILE_READS
    pos = np.arange(start_global, start_global + L, dtype=np.int64)
    codes = ((ref.seq2[pos >> 4] >> ((pos & 15) * 2).astype(np.uint32)) & 3).astype(np.uint8)
    invalid = ((ref.inv[pos >> 5] >> (pos & 31).astype(np.uint32)) & 1).astype(bool)
    t = np.nonzero((codes == 3) & ~invalid)[0]
    if len(t):
        codes[t[0]] = 1
    bb = (L + 3) // 4
    padded = np.zeros(bb * 4, dtype=np.uint8)
    padded[:L] = codes
    row = (padded[0::4] | (padded[1::4] << 2) | (padded[2::4] << 4) | (padded[3::4] << 6)).astype(np.uint8)
    nt = (count + T - 1) // T
    edges = np.minimum(np.arange(nt + 1, dtype=np.uint64) * np.uint64(T), np.uint64(count))
    return ReadBatch(
        count, np.full(count, L | (1 << 16), dtype=np.uint32), np.full(count, start_global, dtype=np.uint32),
        np.concatenate((np.tile(row, count), np.zeros(64, np.uint8))),
        np.concatenate((np.full(count * L, 30, dtype=np.uint8), np.zeros(64, np.uint8))),
        np.concatenate((np.full(count, L << 4, dtype=np.uint32), np.zeros(16, np.uint32))), edges * np.uint64(bb),
        edges * np.uint64(L), edges.copy(), np.zeros(nt + 1, dtype=np.uint32), np.zeros(16, dtype=np.uint32),
        uniform_len=L, uniform_ncigar=1, bases_bytes=count * bb, qual_bytes=count * L, cigar_count=count, exc_count=0)


def concat_uniform(head, body: ReadBatch, tail=None) -> ReadBatch:
    """head ++ body ++ tail for uniform batches (one length, one cigar op per read).  head and tail hold no N calls and
    head is a whole number of tiles, so the body's exception list keeps its (read in tile, position) keys."""
    T = abi.PS_TILE_READS
    L, bb = body.uniform_len, (body.uniform_len + 3) // 4
    parts = [p for p in (head, body, tail) if p is not None]
    for p in parts:
        if p.uniform_len != L or p.uniform_ncigar != 1 or (p is not body and p.exc_count):
            raise ValueError("concat_uniform: parts must be uniform, head / tail without N calls")
    if head is not None and head.n_reads % T:
        raise ValueError("concat_uniform: head must be a whole number of tiles")
    n = sum(p.n_reads for p in parts)
    nt = (n + T - 1) // T
    edges = np.minimum(np.arange(nt + 1, dtype=np.uint64) * np.uint64(T), np.uint64(n))
    k_head = head.n_reads // T if head is not None else 0
    nt_body = (body.n_reads + T - 1) // T
    teo = np.zeros(nt + 1, dtype=np.uint32)
    teo[k_head:k_head + nt_body + 1] = body.tile_exc_off[:nt_body + 1]
    teo[k_head + nt_body + 1:] = body.tile_exc_off[nt_body]
    pad8, pad32 = np.zeros(64, np.uint8), np.zeros(16, np.uint32)
    return ReadBatch(
        n, np.concatenate([p.meta[:p.n_reads] for p in parts]), np.concatenate([p.ref_start[:p.n_reads] for p in parts]),
        np.concatenate([p.bases2[:p.n_reads * bb] for p in parts] + [pad8]),
        np.concatenate([p.qual[:p.n_reads * L] for p in parts] + [pad8]),
        np.concatenate([p.cigar[:p.n_reads] for p in parts] + [pad32]), edges * np.uint64(bb), edges * np.uint64(L),
        edges.copy(), teo, np.concatenate((body.exc[:body.exc_count], pad32)), uniform_len=L, uniform_ncigar=1,
        bases_bytes=n * bb, qual_bytes=n * L, cigar_count=n, exc_count=body.exc_count)


def trim_uniform(batch: ReadBatch, min_len: int, seed: int = 1) -> ReadBatch:
    """A uniform single-M batch with every read cut to its own length in [min_len, L], 3' end of the stored sequence
    removed (what adapter trimming leaves): a ragged batch, still one `xM` op per read, same starts.  Special reads
    (unmapped, POS 0, duplicates) keep their flags.  Vectorised, for workloads of millions of reads."""
    if not (batch.uniform_len and batch.uniform_ncigar == 1):
        raise ValueError("trim_uniform: uniform single-op batch expected")
    n, L, T = batch.n_reads, batch.uniform_len, abi.PS_TILE_READS
    bb = (L + 3) // 4
    rng = np.random.default_rng(seed)
    Lr = rng.integers(min_len, L + 1, size=n, dtype=np.int64)
    meta = (batch.meta[:n] & np.uint32(0xFFFF0000)) | Lr.astype(np.uint32)
    cigar = (Lr.astype(np.uint32) << np.uint32(4))
    q2 = batch.qual[:n * L].reshape(n, L)
    qual = q2[np.arange(L)[None, :] < Lr[:, None]]
    nb = (Lr + 3) // 4
    b2 = batch.bases2[:n * bb].reshape(n, bb).copy()
    last = nb - 1                                       # clear the bits behind the last kept base
    keep_bits = ((Lr - 1) % 4 + 1) * 2
    rows = np.arange(n)
    b2[rows, last] &= ((1 << keep_bits) - 1).astype(np.uint8)
    bases = b2[np.arange(bb)[None, :] < nb[:, None]]
    nt = (n + T - 1) // T
    edges = np.minimum(np.arange(nt + 1, dtype=np.int64) * T, n)
    cq = np.concatenate(([0], np.cumsum(Lr)))
    cb = np.concatenate(([0], np.cumsum(nb)))
    # N calls: keep those inside the kept part; the list stays sorted by (read in tile, position) within each tile
    exc = batch.exc[:batch.exc_count]
    teo = batch.tile_exc_off[:nt + 1].astype(np.int64)
    tile_of = np.repeat(np.arange(nt), np.diff(teo))
    read_of = tile_of * T + (exc >> np.uint32(16)).astype(np.int64)
    keep = (exc & np.uint32(0xFFFF)).astype(np.int64) < Lr[read_of] if len(exc) else np.zeros(0, dtype=bool)
    new_teo = np.concatenate(([0], np.cumsum(np.bincount(tile_of[keep], minlength=nt)))).astype(np.uint32)
    has = np.zeros(n, dtype=bool)
    has[read_of[keep]] = True
    inv_bit = np.uint32(abi.PS_RF_HAS_INVALID << 24)
    meta = np.where(has, meta | inv_bit, meta & ~inv_bit).astype(np.uint32)
    pad8, pad32 = np.zeros(64, np.uint8), np.zeros(16, np.uint32)
    return ReadBatch(
        n, meta, batch.ref_start[:n].copy(), np.concatenate((bases, pad8)), np.concatenate((qual, pad8)),
        np.concatenate((cigar, pad32)), cb[edges].astype(np.uint64), cq[edges].astype(np.uint64), edges.astype(np.uint64),
        new_teo, np.concatenate((exc[keep], pad32)), uniform_len=0, uniform_ncigar=1, bases_bytes=int(cb[-1]),
        qual_bytes=int(cq[-1]), cigar_count=n, exc_count=int(keep.sum()))
