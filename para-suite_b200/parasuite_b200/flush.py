"""Host side of the `clust` tool after the record loop: the per-cluster flush (PileupClusters.java:178-344) and the
end-of-run files (:502-545), on the cluster / site records of the pileup kernels.

The arithmetic (SNP filter, HashMap-order anchor tie-break, sorted T>C fractions, allele statistics) is native
(csrc/flush.cpp, ps_flush_*); this module wraps it and formats the numeric output files the way Java prints doubles.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import abi


def java_double(x: float) -> str:
    """Double.toString (JDK 19+ shortest-repr rules; older JDKs may print a longer digit string in rare cases)."""
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Infinity" if x > 0 else "-Infinity"
    if x == 0:
        return "-0.0" if str(x).startswith("-") else "0.0"
    sign = "-" if x < 0 else ""
    a = abs(x)
    digits, exp = f"{a:.17e}".split("e")          # placeholder, replaced by the shortest repr below
    r = repr(a)
    if "e" in r or "E" in r:
        mant, e = r.lower().split("e")
        exp10 = int(e)
    else:
        mant, exp10 = r, 0
    if "." in mant:
        ip, fp = mant.split(".")
    else:
        ip, fp = mant, ""
    ds = (ip + fp).lstrip("0")
    # decimal exponent of the first significant digit
    point = len(ip) + exp10 - (len(ip + fp) - len((ip + fp).lstrip("0")))
    ds = ds.rstrip("0") or "0"
    if 1e-3 <= a < 1e7:
        if point <= 0:
            s = "0." + "0" * (-point) + ds
        elif point >= len(ds):
            s = ds + "0" * (point - len(ds)) + ".0"
        else:
            s = ds[:point] + "." + ds[point:]
        return sign + s
    return sign + ds[0] + "." + (ds[1:] or "0") + "E" + str(point - 1)


class Flush:
    def __init__(self, contig_names: Sequence[str], min_read_coverage: int = 1,
                 snps: Iterable[Tuple[str, int, str, str]] = (), vcf: Optional[str] = None):
        self.lib = abi.load_library()
        names = (C.c_char_p * len(contig_names))(*[n.encode() for n in contig_names])
        h = C.c_void_p()
        st = self.lib.ps_flush_create(C.byref(h), min_read_coverage, len(contig_names), names)
        if st != abi.PS_OK:
            raise abi.PsError(st, self.lib.ps_strerror(st).decode())
        self.h = h
        for chrom, pos, ref, alt in snps:
            self.lib.ps_flush_add_snp(h, chrom.encode(), pos, ref.encode(), alt.encode())
        if vcf:
            st = self.lib.ps_flush_load_vcf(h, vcf.encode())
            if st != abi.PS_OK:
                raise abi.PsError(st, self.lib.ps_flush_error(h).decode())

    def clusters(self, clusters: np.ndarray, sites: np.ndarray) -> np.ndarray:
        """Flush closed clusters in order (may be called chunk by chunk); returns one FLUSH_ROW per cluster."""
        clusters = np.ascontiguousarray(clusters)
        sites = np.ascontiguousarray(sites)
        rows = np.zeros(len(clusters), dtype=abi.FLUSH_ROW_DTYPE)
        st = self.lib.ps_flush_clusters(self.h, clusters.ctypes.data, len(clusters), sites.ctypes.data if len(sites) else None,
                                        rows.ctypes.data)
        if st != abi.PS_OK:
            raise abi.PsError(st, self.lib.ps_flush_error(self.h).decode() or self.lib.ps_strerror(st).decode())
        return rows

    def totals(self) -> dict:
        t = abi.ps_flush_totals()
        self.lib.ps_flush_totals_get(self.h, C.byref(t), None, 0)
        afi = np.zeros(int(t.n_allele_frequency), dtype=np.float64)
        self.lib.ps_flush_totals_get(self.h, C.byref(t), afi.ctypes.data if len(afi) else None, len(afi))
        return {"snp_hit": int(t.snp_hit), "high_frequent_error": int(t.high_frequent_error),
                "num_crosslinked_clusters": int(t.num_crosslinked_clusters),
                "num_allele_positions": int(t.num_allele_positions),
                "allele_positions": [int(x) for x in t.allele_positions], "allele_frequency_information": afi}

    # ---- end-of-run files (PileupClusters.java:529-545) ---------------------------------------------------
    def sitefrequency_lines(self) -> List[str]:
        t = self.totals()
        n = t["num_crosslinked_clusters"]
        with np.errstate(divide="ignore", invalid="ignore"):
            return [java_double(float(np.float64(v) / np.float64(n))) for v in t["allele_frequency_information"]]

    def sitepositions_lines(self) -> List[str]:
        t = self.totals()
        n = t["num_allele_positions"]
        with np.errstate(divide="ignore", invalid="ignore"):
            return [java_double(float(np.float64(v) / np.float64(n))) for v in t["allele_positions"]]

    def close(self):
        if getattr(self, "h", None):
            self.lib.ps_flush_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ClustWriter:
    """The output files of the `clust` tool (PileupClusters.java:262-343, :367-487, :502-545) written natively
    (csrc/clust_writer.cpp): <out>, <out>.ccr.fasta, <out>.ccr.tsv, <out>.report, <bam>.sitefrequency.tsv,
    <bam>.sitepositions.tsv.  Feed it, in file order, the host batches the kernels saw and the clusters that closed
    with each of them; it runs them through `flush` itself."""

    def __init__(self, flush: Flush, fasta_path: str, out_path: str, bam_path: str):
        self.lib = flush.lib
        self.flush = flush
        h = C.c_void_p()
        st = self.lib.ps_clust_writer_open(C.byref(h), flush.h, fasta_path.encode(), out_path.encode(), bam_path.encode())
        self.h = h
        if st != abi.PS_OK:
            msg = self.lib.ps_clust_writer_error(h).decode() if h else ""
            self.close()
            raise abi.PsError(st, msg or self.lib.ps_strerror(st).decode())

    def feed(self, batch, first_ordinal: int, clusters: np.ndarray, sites: np.ndarray, open_first_read: Optional[int] = None):
        s = batch.struct if hasattr(batch, "struct") else batch.as_struct()
        clusters = np.ascontiguousarray(clusters)
        sites = np.ascontiguousarray(sites)
        st = self.lib.ps_clust_writer_feed(self.h, C.byref(s), int(first_ordinal), clusters.ctypes.data if len(clusters) else None,
                                           len(clusters), sites.ctypes.data if len(sites) else None,
                                           0 if open_first_read is None else 1, int(open_first_read or 0))
        if st == abi.PS_ERR_REFERENCE_WOULD_THROW:
            f = abi.ps_fault()
            self.lib.ps_clust_writer_fault(self.h, C.byref(f))
            raise abi.ReferenceWouldThrow(st, self.lib.ps_clust_writer_error(self.h).decode(), (f.code, f.read_ordinal))
        if st != abi.PS_OK:
            raise abi.PsError(st, self.lib.ps_clust_writer_error(self.h).decode() or self.lib.ps_strerror(st).decode())

    def finish(self, counters: dict) -> dict:
        ctr = abi.ps_pileup_counters()
        ctr.skipped_due_indel = int(counters["skipped_due_indel"])
        ctr.double_stranded = int(counters["double_stranded"])
        ctr.num_reads_processed = int(counters.get("num_reads_processed", 0))
        st = self.lib.ps_clust_writer_finish(self.h, C.byref(ctr))
        if st != abi.PS_OK:
            raise abi.PsError(st, self.lib.ps_clust_writer_error(self.h).decode())
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.lib.ps_clust_writer_stats(self.h, C.byref(a), C.byref(b), C.byref(c))
        return {"rows": a.value, "ccr_rows": b.value, "ccr_start_before_contig": c.value}

    def close(self):
        if getattr(self, "h", None):
            self.lib.ps_clust_writer_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
