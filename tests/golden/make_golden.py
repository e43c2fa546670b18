#!/usr/bin/env python
"""Generates tests/golden/parasuite_golden_v1.json with the literal Python restatement of the Java loops
(oracle/py_oracle.py) -- NOT with the C++ oracle and not with the CUDA path, so the fixture pins all three.

  python tests/golden/make_golden.py            (from the repo root; deterministic: seeded)

The reference (Java) cannot run in this image and ships no expected outputs, so these are golden vectors of the
restatement ("parity unpinned"): they freeze its behaviour on a record set that walks every CIGAR
class, both strands, N / IUPAC / soft-masked reference bases, N base calls, special flags, SNP filtering and the
HashMap-order anchor tie-break."""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(REPO, "para-suite_b200"), os.path.join(REPO, "oracle"), os.path.join(REPO, "tests")):
    sys.path.insert(0, p)

import py_oracle as po  # noqa: E402
from helpers import kat_records, random_genome, random_records, to_py  # noqa: E402
from kat_vectors import PILEUP_READS  # noqa: E402

MAX_LEN = 64
MIN_COV = 2


def build():
    rng = random.Random(20161007)
    contigs = random_genome(rng, n_contigs=2, length=12000, n_frac=0.01, lower_frac=0.1)
    genome = po.Genome(dict(contigs))
    recs = random_records(rng, contigs, 700, kinds=("M", "M", "M", "clip", "indel", "splice", "wild"), Lrange=(12, 40),
                          flags_special=0.04)
    keep = []
    for r in recs:          # records the JVM survives in both tools
        try:
            po.profile(to_py([r]), genome, MAX_LEN)
            if r.pos > 0:
                po.pileup(to_py([r]), genome, po.SnpDb([]), 1)
            keep.append(r)
        except po.ReferenceWouldThrow:
            pass
    prof = po.profile(to_py(keep), genome, MAX_LEN).wrapped()
    pile_recs = [r for r in keep if r.pos > 0]
    # SNPs: every 4th T>C site of an SNP-free run, named as SNPCalling.querySNP asks for them ("chr" stripped)
    dry = po.pileup(to_py(pile_recs), genome, po.SnpDb([]), MIN_COV)
    snps = []
    for k, c in enumerate(dry.clusters):
        if k % 4 == 0 and c.sites:
            snps.append([c.chrom[3:], c.sites[0][0], "T", "C"])
    st = po.pileup(to_py(pile_recs), genome, po.SnpDb([tuple(s) for s in snps]), MIN_COV)
    clusters = []
    for c in st.clusters:
        clusters.append({
            "running_id": c.running_id, "chrom": c.chrom, "start": c.start, "end": c.end, "first_reverse": bool(c.first_reverse),
            "num_reads": c.num_reads, "num_t2c": c.num_t2c, "combined_strand": c.combined_strand,
            "mask51": [j for j, b in enumerate(c.mask51) if b], "sites_in_insertion_order": [list(s) for s in c.sites],
            "emitted": c.emitted, "num_t2c_sites": c.num_t2c_sites, "fraction": c.fraction, "best_pos": c.best_pos,
            "best_value": c.best_value, "best_count": c.best_count})
    return {
        "version": 1, "generator": "tests/golden/make_golden.py (oracle/py_oracle.py)", "max_read_length": MAX_LEN,
        "min_read_coverage": MIN_COV,
        "contigs": [[n, s.decode()] for n, s in contigs],
        "records": [[r.flag, r.rname, r.pos, r.cigar, r.seq.decode(), list(r.qual)] for r in keep],
        "snps": snps,
        "profile": prof,
        "pileup": {
            "clusters": clusters,
            "num_reads_processed": st.num_reads_processed, "skipped_due_indel": st.skipped_due_indel,
            "double_stranded": st.double_stranded, "snp_hit": st.snp_hit, "high_frequent_error": st.high_frequent_error,
            "num_crosslinked_clusters": st.num_crosslinked_clusters, "num_allele_positions": st.num_allele_positions,
            "allele_positions": st.allele_positions, "allele_frequency_information": st.allele_frequency_information,
            "open_cluster_start": st.open_cluster.start if st.open_cluster else None,
        },
    }


if __name__ == "__main__":
    out = os.path.join(HERE, "parasuite_golden_v1.json")
    with open(out, "w") as f:
        json.dump(build(), f, separators=(",", ":"))
    print("wrote", out, os.path.getsize(out), "bytes")
