"""
Hand-derived known-answer vectors for the hot path.  The reference ships no tests
or golden outputs, so these were derived by hand from the Java source
(ErrorProfiling.java:146-409, PileupClusters.java:137-500,585-673) and are checked against BOTH
oracle restatements (oracle/py_oracle.py and oracle/parasuite_oracle.cpp) and the CUDA path.

Expected values are written as  "pos:ref>read"  increments of positionConversions.
"""

KAT_REF = "ACGTTGCAAGGCTTACGATCGGATCCTTAGNNacgtTTGACC"      # chr1, 1-based
KAT_MAXLEN = 12
BASES = "ACGT"


def q(lo, n):
    return bytes(range(lo, lo + n))


# id -> (records, expected)
# record: (flag, pos, cigar, seq, qual)
PROFILE_KATS = {
    "A": dict(
        reads=[(0, 3, "8M", "GTCGCAAG", q(30, 8))],
        conv="0:G>G 1:T>T 2:T>C 3:G>G 4:C>C 5:A>A 6:A>A 7:G>G",
        qsum={"A>A": 71, "C>C": 34, "G>G": 100, "T>C": 32, "T>T": 31},
        counters=dict(num_reads_processed=1, total_bases_checked=8)),
    "B": dict(
        reads=[(16, 3, "8M", "GTCGCAAG", q(30, 8))],
        conv="0:C>C 1:T>T 2:T>T 3:G>G 4:C>C 5:A>G 6:A>A 7:C>C",
        qsum={"A>G": 35, "A>A": 36, "C>C": 101, "G>G": 33, "T>T": 63},
        counters=dict(num_reads_processed=1, total_bases_checked=8)),
    "C": dict(
        reads=[(0, 5, "2S6M", "GGTGCAAG", bytes([20] * 8))],
        conv="0:T>G 1:G>G 2:C>T 3:A>G 4:A>C 5:G>A",
        qsum={"T>G": 20, "G>G": 20, "C>T": 20, "A>G": 20, "A>C": 20, "G>A": 20},
        counters=dict(indel_read=1, total_bases_checked=6)),
    "D": dict(
        reads=[(0, 5, "6M2S", "TGCAAGTT", bytes([20] * 8))],
        conv="0:T>T 1:G>G 2:C>C 3:A>A 4:A>A 5:G>G",
        counters=dict(indel_read=1, total_bases_checked=6)),
    "E": dict(
        reads=[(16, 5, "6M2S", "TGCAAGTT", bytes([20] * 8))],
        conv="2:C>C 3:T>T 4:T>T 5:G>G 6:C>C 7:A>A",
        counters=dict(indel_read=1, total_bases_checked=6)),
    "F": dict(
        reads=[(0, 3, "4M1D4M", "GTTGAAGG", bytes([20] * 8))],
        conv="0:G>G 1:T>T 2:T>T 3:G>G 5:A>A 6:A>A 7:G>G 8:G>G",
        qsum={}, dels={6: 1}, counters=dict(indel_read=1, total_bases_checked=8)),
    "G": dict(
        reads=[(0, 3, "4M1I4M", "GTTGACAAG", bytes([20] * 9))],
        conv="0:G>G 1:T>T 2:T>T 3:G>G 5:C>C 6:A>A 7:A>A 8:G>G",
        qsum={}, ins={6: 1}, counters=dict(indel_read=1, total_bases_checked=8)),
    "H": dict(
        reads=[(0, 3, "3M1I2M1D3M", "GTTAGCAGG", bytes([20] * 9))],
        conv="0:G>G 1:T>T 2:T>T 3:G>A 4:C>G 5:A>C 6:A>A 7:G>G 8:G>G",
        qsum={}, counters=dict(indel_read=0, total_bases_checked=9)),
    "I": dict(
        reads=[(0, 3, "3M2I2M1D3M", "GTTAAGCAGG", bytes([20] * 10))],
        conv="", ins={6: 1, 7: 1}, dels={9: 1},
        counters=dict(longer_indels=1, indel_read=1, skipped_reads=1, total_bases_checked=0)),
    "J": dict(
        reads=[(0, 3, "4M3N4M", "GTTGAGGC", bytes([20] * 8))],
        conv="", counters=dict(indel_read=1, skipped_reads=1, total_bases_checked=0)),
    "K": dict(
        reads=[(0, 29, "8M", "AGNNACGT", bytes([20] * 8))],
        conv="0:A>A 1:G>G 4:A>A 5:C>C 6:G>G 7:T>T",
        counters=dict(total_bases_checked=6)),
    "L": dict(
        reads=[(0x400, 3, "8M", "GTCGCAAG", q(30, 8)),
               (0x4, 3, "8M", "GTCGCAAG", q(30, 8)),
               (0, 0, "8M", "GTCGCAAG", q(30, 8))],
        conv="", counters=dict(duplicates=1, unmapped=1, start_zero=1, num_reads_processed=0,
                               total_bases_checked=0)),
}

# ---- pileup ------------------------------------------------------------------------------
PILEUP_REF = KAT_REF + "ATTTGCATGCATTTACG"
PILEUP_READS = [
    (0, 3, "8M", "GTCGCAAG"),
    (0, 4, "8M", "TCGCAAGG"),
    (16, 6, "8M", "GCAGGGCT"),
    (0, 9, "4M1D4M", "AGGCTACG"),
    (0, 20, "8M", "CGGATCCT"),
    (16, 22, "2S6M", "AAGATCCT"),
    (0, 40, "8M", "ACCATTTG"),
]
# closed clusters in order (the open one starting at 40 is never emitted by the reference)
PILEUP_EXPECT = [
    dict(cluster_id="cl_2_chr1", start=3, end=13, first_reverse=False, num_reads=3, num_t2c=3,
         sites={5: (2, 2), 9: (1, 3)}, combined="+/-", mask=[1, 2, 4]),
    dict(cluster_id="cl_3_chr1", start=9, end=17, first_reverse=False, num_reads=1, num_t2c=0,
         sites={}, combined="+", mask=[]),
    dict(cluster_id="cl_4_chr1", start=20, end=27, first_reverse=False, num_reads=2,
         combined="+/-"),
]
PILEUP_DOUBLE_STRANDED = 2


def parse_conv(s):
    """'2:T>C 3:G>G' -> {(2, 3, 1): 1, ...} keyed (pos, ref_idx, read_idx)."""
    out = {}
    for tok in s.split():
        p, rr = tok.split(":")
        a, b = rr.split(">")
        k = (int(p), BASES.index(a), BASES.index(b))
        out[k] = out.get(k, 0) + 1
    return out
