"""`comb` tool (CombineGenomeTranscript.java): transcript hits lifted to genomic coordinates and merged with the genomic
hits -- native, host only (csrc/liftover.cpp).  Thin ctypes wrappers."""
from __future__ import annotations

import ctypes as C
from typing import Tuple

from . import abi


def liftover_hit(transcript_name: str, aln_start: int, aln_end: int, read_len: int, cigar: str) -> Tuple[int, str, int]:
    """One hit -> (new start or -1, new cigar, missedTranscriptAlignments increment)."""
    lib = abi.load_library()
    start, missed = C.c_int32(-1), C.c_uint32(0)
    buf = C.create_string_buffer(1 << 16)
    st = lib.ps_liftover_hit(transcript_name.encode(), aln_start, aln_end, read_len, cigar.encode(), C.byref(start), buf,
                             len(buf), C.byref(missed))
    if st == abi.PS_ERR_REFERENCE_WOULD_THROW:
        raise abi.ReferenceWouldThrow(st, buf.value.decode(), (0, 0))
    if st != abi.PS_OK:
        raise abi.PsError(st, lib.ps_strerror(st).decode())
    return start.value, buf.value.decode(), missed.value


def comb_bam(genomic_bam: str, transcript_bam: str, out_bam: str) -> dict:
    lib = abi.load_library()
    stats = abi.ps_comb_stats()
    err = C.create_string_buffer(1024)
    st = lib.ps_comb_bam(genomic_bam.encode(), transcript_bam.encode(), out_bam.encode(), C.byref(stats), err, len(err))
    if st == abi.PS_ERR_REFERENCE_WOULD_THROW:
        raise abi.ReferenceWouldThrow(st, err.value.decode(), (0, 0))
    if st != abi.PS_OK:
        raise abi.PsError(st, err.value.decode() or lib.ps_strerror(st).decode())
    return {k: int(getattr(stats, k)) for k, _ in abi.ps_comb_stats._fields_}
