"""CPU: the native cluster flush (csrc/flush.cpp) against the Python oracle's restatement of
PileupClusters.java:178-260 / :529-545, fed with the oracle's own cluster and site records (so this isolates the
flush: SNP filter, HashMap-order anchor tie-break, sorted fractions, allele statistics)."""
import os
import random

import numpy as np
import pytest

import py_oracle as po
from helpers import kat_records, random_genome, random_records, to_py
from kat_vectors import PILEUP_READS, PILEUP_REF
from parasuite_b200 import PackedReference, ReadBatch, Record, abi
from parasuite_b200.flush import Flush, java_double

pytestmark = pytest.mark.skipif(not os.path.exists(abi.lib_path()), reason="library not built")


def native_records(oracle, contigs, recs):
    ref = PackedReference.from_contigs(contigs)
    batch = ReadBatch.from_records(recs, ref)
    return oracle.pileup(ref, batch), ref


def check(oracle, contigs, recs, snps, min_cov):
    res, ref = native_records(oracle, contigs, recs)
    st = po.pileup(to_py(recs), po.Genome(dict(contigs)), po.SnpDb(snps), min_cov)
    fl = Flush(ref.names, min_cov, snps=snps)
    # flush in two chunks: the running state (map capacity, allele sums) must carry over
    cl, si = res["clusters"], res["sites"]
    cut = len(cl) // 2
    rows = np.concatenate([fl.clusters(cl[:cut], si), fl.clusters(cl[cut:], si)])
    assert len(rows) == len(st.clusters)
    for k, (row, exp) in enumerate(zip(rows, st.clusters)):
        assert bool(row["emitted"]) == exp.emitted, k
        if not exp.emitted:
            continue
        assert int(row["num_t2c_sites"]) == exp.num_t2c_sites, k
        assert float(row["fraction"]) == exp.fraction, (k, row["fraction"], exp.fraction)          # bit-exact doubles
        assert int(row["best_pos"]) == exp.best_pos, (k, row, exp.best_pos, exp.sites)
        assert float(row["best_value"]) == exp.best_value, k
        if exp.best_pos > 0:
            assert int(row["best_count"]) == exp.best_count and row["has_ccr"] == 1, k
    t = fl.totals()
    assert t["snp_hit"] == st.snp_hit and t["high_frequent_error"] == st.high_frequent_error
    assert t["num_crosslinked_clusters"] == st.num_crosslinked_clusters
    assert t["num_allele_positions"] == st.num_allele_positions
    assert t["allele_positions"] == st.allele_positions
    assert list(t["allele_frequency_information"]) == st.allele_frequency_information                  # bit-exact doubles
    return rows, st, fl


def test_kat(oracle):
    contigs = [("chr1", PILEUP_REF.encode())]
    rows, st, _ = check(oracle, contigs, kat_records(PILEUP_READS), [], 1)
    assert len(rows) == 3 and int(rows[0]["num_t2c_sites"]) == 2


@pytest.mark.parametrize("seed,min_cov,with_snps", [(1, 1, False), (2, 2, True), (3, 1, True), (4, 5, False)])
def test_random(oracle, seed, min_cov, with_snps):
    rng = random.Random(seed)
    contigs = random_genome(rng, n_contigs=3, length=60000, n_frac=0.003, lower_frac=0.05)
    recs = [r for r in random_records(rng, contigs, 6000, kinds=("M",), Lrange=(18, 30)) if r.pos > 0]
    snps = []
    if with_snps:
        res, _ = native_records(oracle, contigs, recs)
        cl, si = res["clusters"], res["sites"]
        names = [n for n, _ in contigs]
        for c in cl[:: 3]:
            for s in si[int(c["site_begin"]):int(c["site_end"])][:1]:
                chrom = names[int(c["contig"])][3:]                 # the query strips "chr"
                snps.append((chrom, int(s["pos"]), "T", "C"))
        snps.append((names[0], 5, "T", "C"))                         # "chr1" never matches the stripped query
        snps.append((names[0][3:], int(si[0]["pos"]) if len(si) else 7, "G", "C"))   # REF without T: ignored
    rows, st, _ = check(oracle, contigs, recs, snps, min_cov)
    assert sum(int(r["emitted"]) for r in rows) > 20
    if with_snps:
        assert st.snp_hit > 0


def test_hashmap_tie_break(oracle):
    """Two sites with equal count/coverage whose bucket order inverts numeric order: keys 17 and 32 at capacity 16
    iterate as 32, 17 -> the LAST maximum wins -> anchor 17."""
    ref = bytearray(b"A" * 80)
    ref[16] = ord("T")      # position 17
    ref[31] = ord("T")      # position 32
    contigs = [("chr1", bytes(ref))]
    read = bytearray(ref[10:40])
    read[6] = ord("C")      # pos 17
    read[21] = ord("C")     # pos 32
    recs = [Record(0, "chr1", 11, "30M", bytes(read), bytes([30] * 30)),
            Record(0, "chr1", 70, "5M", b"AAAAA", bytes([30] * 5))]          # closes the first cluster
    rows, st, _ = check(oracle, contigs, recs, [], 1)
    assert int(rows[0]["best_pos"]) == 17 and st.clusters[0].best_pos == 17
    assert float(rows[0]["best_value"]) == 1.0 and float(rows[0]["fraction"]) == 2.0


def test_first_cluster_doubles_allele_sums(oracle):
    """alleleFrequencyInformation: addAll at k = 0, then set(k, get(k) + v) for k >= 1 -> entries 1.. doubled (:233-248)."""
    ref = bytearray(b"A" * 60)
    for p in (12, 15):
        ref[p] = ord("T")
    contigs = [("chr1", bytes(ref))]
    r1 = bytearray(ref[10:30]); r1[2] = ord("C"); r1[5] = ord("C")
    r2 = bytearray(ref[10:30]); r2[2] = ord("C")
    recs = [Record(0, "chr1", 11, "20M", bytes(r1), bytes([30] * 20)), Record(0, "chr1", 11, "20M", bytes(r2), bytes([30] * 20)),
            Record(0, "chr1", 50, "5M", b"AAAAA", bytes([30] * 5))]
    rows, st, fl = check(oracle, contigs, recs, [], 1)
    assert list(fl.totals()["allele_frequency_information"]) == [1.0, 1.0]      # sorted = [1.0, 0.5]; 0.5 added twice
    assert fl.sitefrequency_lines() == ["1.0", "1.0"]
    assert len(fl.sitepositions_lines()) == 51


def test_vcf_loader(tmp_path, oracle):
    import gzip
    vcf = tmp_path / "snp.vcf.gz"
    with gzip.open(vcf, "wt") as f:
        f.write("##fileformat=VCFv4.1\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n1\t17\t.\tT\tC,G\t.\t.\t.\n1\t32\t.\tT\tG\t.\t.\t.\n")
    ref = bytearray(b"A" * 80); ref[16] = ord("T"); ref[31] = ord("T")
    contigs = [("chr1", bytes(ref))]
    read = bytearray(ref[10:40]); read[6] = ord("C"); read[21] = ord("C")
    recs = [Record(0, "chr1", 11, "30M", bytes(read), bytes([30] * 30)), Record(0, "chr1", 70, "5M", b"AAAAA", bytes([30] * 5))]
    res, refp = native_records(oracle, contigs, recs)
    fl = Flush(refp.names, 1, vcf=str(vcf))
    rows = fl.clusters(res["clusters"], res["sites"])
    assert fl.totals()["snp_hit"] == 1 and int(rows[0]["best_pos"]) == 32 and int(rows[0]["num_t2c_sites"]) == 2
    exp = po.pileup(to_py(recs), po.Genome(dict(contigs)), po.SnpDb([("1", 17, "T", "C"), ("1", 32, "T", "G")]), 1)
    assert exp.clusters[0].best_pos == 32 and exp.snp_hit == 1


def test_java_double_formatting():
    for x, e in [(0.5, "0.5"), (1.0, "1.0"), (1e-4, "1.0E-4"), (1e7, "1.0E7"), (1 / 3, "0.3333333333333333"),
                 (0.0, "0.0"), (float("nan"), "NaN"), (123456.789, "123456.789"), (2.5e-5, "2.5E-5")]:
        assert java_double(x) == e


def test_thin_cluster_still_grows_the_map(oracle):
    """A cluster below minReadCoverage is not flushed, but the record loop has put its T>C keys into the run's one
    mutationMap (PileupClusters.java:126, :353, :651-661): 13 keys grow the table from 16 to 32 buckets and clear() keeps
    that capacity, so the tied sites 80 and 97 of the next cluster iterate as 97, 80 (buckets 1, 16) instead of 80, 97
    (buckets 0, 1) and the anchor -- the LAST maximum -- is 80."""
    ref = bytearray(b"A" * 140)
    t_pos = list(range(12, 12 + 26, 2))            # 13 T positions (1-based) inside the first read
    for p in t_pos + [80, 97]:
        ref[p - 1] = ord("T")
    contigs = [("chr1", bytes(ref))]
    r1 = bytearray(ref[10:50])                      # start 11, 40M
    for p in t_pos:
        r1[p - 11] = ord("C")
    r2 = bytearray(ref[69:109])                     # start 70, 40M: covers 80 and 97
    r2[80 - 70] = ord("C"); r2[97 - 70] = ord("C")
    r3 = bytearray(ref[70:110])                     # start 71, 40M: the same two conversions -> both sites at 2/2
    r3[80 - 71] = ord("C"); r3[97 - 71] = ord("C")
    recs = [Record(0, "chr1", 11, "40M", bytes(r1), bytes([30] * 40)),
            Record(0, "chr1", 70, "40M", bytes(r2), bytes([30] * 40)),
            Record(0, "chr1", 71, "40M", bytes(r3), bytes([30] * 40)),
            Record(0, "chr1", 130, "5M", b"AAAAA", bytes([30] * 5))]      # closes the second cluster
    rows, st, _ = check(oracle, contigs, recs, [], 2)
    assert not rows[0]["emitted"] and rows[1]["emitted"]
    assert st.clusters[1].best_pos == 80 and int(rows[1]["best_pos"]) == 80
