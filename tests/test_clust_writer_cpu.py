"""CPU: the native writer of the `clust` output files (csrc/clust_writer.cpp: cluster sequence assembly, cluster rows, CCR
FASTA / TSV, .report, sitefrequency, sitepositions) against the literal Python restatement of PileupClusters.java:62-545
(oracle/py_oracle.py: clust_files), byte for byte.  Cluster / site records come from the C++ oracle here (the GPU tests
feed the writer with the kernels' records); the writer itself never decides a cluster boundary."""
import os
import random

import numpy as np
import pytest

import py_oracle as po
from helpers import kat_records, random_genome, random_records, to_py
from kat_vectors import PILEUP_READS, PILEUP_REF
from parasuite_b200 import PackedReference, ReadBatch, Record, abi
from parasuite_b200.bamio import write_fasta
from parasuite_b200.flush import ClustWriter, Flush
from parasuite_b200.sharding import slice_batch

pytestmark = pytest.mark.skipif(not os.path.exists(abi.lib_path()), reason="library not built")

FILES = {"pileup": "{out}", "ccr.fasta": "{out}.ccr.fasta", "ccr.tsv": "{out}.ccr.tsv", "report": "{out}.report",
         "sitefrequency": "{bam}.sitefrequency.tsv", "sitepositions": "{bam}.sitepositions.tsv"}


def run_native(tmp_path, oracle, contigs, recs, snps, min_cov, windows=1, line_width=60):
    fa, out, bam = str(tmp_path / "ref.fa"), str(tmp_path / "clusters.tsv"), str(tmp_path / "reads.bam")
    write_fasta(fa, contigs, line_width=line_width)
    ref = PackedReference.from_contigs(contigs)
    batch = ReadBatch.from_records(recs, ref)
    res = oracle.pileup(ref, batch)
    fl = Flush(ref.names, min_cov, snps=snps)
    w = ClustWriter(fl, fa, out, bam)
    cl, si = res["clusters"], res["sites"]
    open_first = None if res["open_cluster"] is None else int(res["open_cluster"]["first_read"])
    if windows == 1:
        w.feed(batch, 0, cl, si, open_first)
    else:
        # the same stream in several feeds: every feed brings the clusters that closed inside its reads and names the
        # cluster still open behind it
        n = batch.n_reads
        cuts = sorted(set([0, n] + [n * k // windows + 3 for k in range(1, windows)]))
        starts = [int(x) for x in cl["first_read"]] + ([open_first] if open_first is not None else [])
        done = 0
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            # clusters closed by a read of [lo, hi): those whose successor starts in [lo, hi)
            k1 = done
            while k1 < len(cl) and starts[k1 + 1] < hi:
                k1 += 1
            sub = cl[done:k1].copy()
            s0 = int(cl["site_begin"][done]) if k1 > done else 0
            s1 = int(cl["site_end"][k1 - 1]) if k1 > done else 0
            sub["site_begin"] -= np.uint64(s0)
            sub["site_end"] -= np.uint64(s0)
            nxt = starts[k1] if k1 < len(starts) else None
            w.feed(slice_batch(batch, lo, hi), lo, sub, si[s0:s1], nxt)
            done = k1
    stats = w.finish(res["counters"])
    got = {k: open(v.format(out=out, bam=bam)).read() for k, v in FILES.items()}
    w.close()
    return got, stats


def check(tmp_path, oracle, contigs, recs, snps, min_cov, **kw):
    exp = po.clust_files(to_py(recs), po.Genome(dict(contigs)), po.SnpDb(snps), min_cov)
    got, stats = run_native(tmp_path, oracle, contigs, recs, snps, min_cov, **kw)
    for k in FILES:
        assert got[k] == exp[k], (k, got[k][:400], exp[k][:400])
    return got, stats


def test_kat(tmp_path, oracle):
    contigs = [("chr1", PILEUP_REF.encode())]
    got, stats = check(tmp_path, oracle, contigs, kat_records(PILEUP_READS), [], 1)
    rows = got["pileup"].splitlines()
    assert len(rows) == 4 and rows[1].startswith("cl_2_chr1\tchr1\t3\t13\t+\t3\t3\t2\t")
    assert stats["rows"] == 3


@pytest.mark.parametrize("threads", [None, 5])
@pytest.mark.parametrize("seed,kinds,min_cov,windows", [(1, ("M",), 1, 1), (2, ("M", "clip", "indel"), 2, 1),
                                                        (3, ("M", "clip", "indel", "splice"), 1, 4), (4, ("M",), 3, 7)])
def test_random(tmp_path, oracle, monkeypatch, seed, kinds, min_cov, windows, threads):
    """threads: the record loop of a feed is walked in that many ranges, each beginning at a cluster start (the default
    takes one range for inputs this small)."""
    if threads:
        monkeypatch.setenv("PARASUITE_B200_WRITER_THREADS", str(threads))
    rng = random.Random(seed)
    contigs = random_genome(rng, n_contigs=3, length=30000, n_frac=0.003, lower_frac=0.3)
    recs = [r for r in random_records(rng, contigs, 5000, kinds=kinds, Lrange=(18, 30)) if r.pos > 0]
    g = po.Genome(dict(contigs))
    ok = []
    for r in recs:                     # records on which the JVM dies in the record loop are tested elsewhere
        try:
            po.pileup(to_py([r]), g, po.SnpDb([]), 1)
            ok.append(r)
        except po.ReferenceWouldThrow:
            pass
    try:
        po.clust_files(to_py(ok), g, po.SnpDb([]), min_cov)
    except po.JvmWouldDie:             # a soft-clipped read shifted past a contig end: drop reads near contig ends
        ok = [r for r in ok if r.pos + 80 < 30000]
    snps = []
    got, stats = check(tmp_path, oracle, contigs, ok, snps, min_cov, windows=windows, line_width=61)
    assert stats["rows"] > 30 and got["ccr.fasta"].count(">") == stats["ccr_rows"]
    if "clip" in kinds:                # the soft-clip shift and the lower-case reference must show up in the sequences
        assert any(c.islower() for c in got["pileup"])


def test_minus_first_cluster_is_reverse_complemented(tmp_path, oracle):
    ref = b"ACGTTGCAAGGCTTACGATCGGATCCTTAGacgtacgtTTGACCATTTGCATGCATTTACG"
    contigs = [("chr1", ref)]
    read = bytearray(ref[4:24].upper())
    recs = [Record(16, "chr1", 5, "20M", bytes(read), bytes([30] * 20)),
            Record(0, "chr1", 8, "20M", ref[7:27].upper(), bytes([30] * 20)),
            Record(0, "chr1", 50, "5M", ref[49:54].upper(), bytes([30] * 5))]
    got, _ = check(tmp_path, oracle, contigs, recs, [], 1)
    row = got["pileup"].splitlines()[1].split("\t")
    assert row[4] == "-" and row[10] == "-" and int(row[11]) == 23      # 5..27
    fwd = ref[4:27].decode()
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "a": "t", "c": "g", "g": "c", "t": "a"}
    assert row[9] == "".join(comp[c] for c in reversed(fwd))


def test_fetch_past_contig_end_is_reported(tmp_path, oracle):
    """A leading soft clip shifts the fetch window of the cluster sequence (every non-I element advances the cursor,
    :402-404): 10S20M at 35 of a 60-base contig fetches 45..64 -> SAMException outside any try block."""
    ref = b"ACGT" * 15
    contigs = [("chr1", ref)]
    recs = [Record(0, "chr1", 35, "10S20M", b"A" * 30, bytes([30] * 30))]
    with pytest.raises(po.JvmWouldDie):
        po.clust_files(to_py(recs), po.Genome(dict(contigs)), po.SnpDb([]), 1)
    with pytest.raises(abi.ReferenceWouldThrow) as e:
        run_native(tmp_path, oracle, contigs, recs, [], 1)
    assert e.value.fault == (abi.PS_THROW_REF_RANGE, 0)


def test_fault_in_the_middle_of_a_feed(tmp_path, oracle, monkeypatch):
    """The record the JVM dies on sits in the middle of the stream: the same fault is reported however many ranges the
    feed is walked in, and the rows of the clusters that ended in front of it -- and only those -- are in the file."""
    rng = random.Random(17)
    contigs = random_genome(rng, n_contigs=2, length=20000, n_frac=0.0, lower_frac=0.1)
    recs = [r for r in random_records(rng, contigs, 3000, kinds=("M",), Lrange=(18, 30)) if 0 < r.pos < 19900]
    L0 = len(contigs[0][1])
    bad = Record(0, contigs[0][0], L0 - 24, "10S20M", b"A" * 30, bytes([30] * 30))     # fetches up to L0 + 5
    order = {n: i for i, (n, _) in enumerate(contigs)}
    recs = sorted(recs + [bad], key=lambda r: (order[r.rname], r.pos))
    at = recs.index(bad)
    assert 500 < at < len(recs) - 500
    fa = str(tmp_path / "ref.fa")
    write_fasta(fa, contigs)
    ref = PackedReference.from_contigs(contigs)
    batch = ReadBatch.from_records(recs, ref)
    res = oracle.pileup(ref, batch)

    def run(threads):
        monkeypatch.setenv("PARASUITE_B200_WRITER_THREADS", str(threads))
        out = str(tmp_path / f"clusters{threads}.tsv")
        w = ClustWriter(Flush(ref.names, 1, snps=[]), fa, out, str(tmp_path / "reads.bam"))
        try:
            with pytest.raises(abi.ReferenceWouldThrow) as e:
                w.feed(batch, 0, res["clusters"], res["sites"], int(res["open_cluster"]["first_read"]))
            assert e.value.fault == (abi.PS_THROW_REF_RANGE, at)
        finally:
            w.close()                                       # writes out what it holds
        return open(out).read()
    got = run(1)
    # the stream cut in front of the bad record: the same rows, plus the cluster that the bad record ended if it began one
    # (the end of a run never writes the last cluster, :502)
    exp = po.clust_files(to_py(recs[:at]), po.Genome(dict(contigs)), po.SnpDb([]), 1)["pileup"]
    assert got.startswith(exp) and got.count("\n") - exp.count("\n") <= 1 and got.count("\n") > 100
    for threads in (2, 6):
        assert run(threads) == got


def test_large_feed_default_threads_equal_one_range(tmp_path, oracle, monkeypatch):
    """300 000 synthetic reads in one feed: the default (several ranges, rows formatted and the three row files written
    side by side) writes the same bytes as one range on one thread, which the small cases above pin on the oracle."""
    from parasuite_b200 import synth
    ref = synth.synth_reference(7, [2_000_000], n_run=1000)
    batch = synth.synth_reads(ref, 300_000, 36, seed=8)
    codes = np.zeros(ref.n_bases, dtype=np.uint8)
    for k in range(16):
        codes[k::16] = ((ref.seq2[: (ref.n_bases + 15) // 16] >> (2 * k)) & 3)[: len(codes[k::16])]
    asc = np.frombuffer(b"ACGT", dtype=np.uint8)[codes].copy()
    asc[np.unpackbits(ref.inv.view(np.uint8), bitorder="little")[: ref.n_bases].astype(bool)] = ord("N")
    fa = str(tmp_path / "ref.fa")
    write_fasta(fa, [(ref.names[0], asc.tobytes())])
    res = oracle.pileup(ref, batch)
    assert len(res["clusters"]) > 8192

    def run(tag, threads):
        if threads:
            monkeypatch.setenv("PARASUITE_B200_WRITER_THREADS", str(threads))
        else:
            monkeypatch.delenv("PARASUITE_B200_WRITER_THREADS", raising=False)
        out, bam = str(tmp_path / f"{tag}.tsv"), str(tmp_path / f"{tag}.bam")
        w = ClustWriter(Flush(ref.names, 1, snps=[]), fa, out, bam)
        w.feed(batch, 0, res["clusters"], res["sites"], int(res["open_cluster"]["first_read"]))
        stats = w.finish(res["counters"])
        w.close()
        return {k: open(v.format(out=out, bam=bam)).read() for k, v in FILES.items()}, stats
    one, s1 = run("one", 1)
    many, s2 = run("many", None)
    assert s1 == s2 and s1["rows"] == len(res["clusters"])
    for k in FILES:
        assert one[k] == many[k], k
    # spot check of a few rows against the reference text: the cluster sequence is the reference over [start, end]
    for line in one["pileup"].splitlines()[1:2000:97]:
        c = line.split("\t")
        seq = asc[int(c[2]) - 1:int(c[3])].tobytes().decode()
        if c[4] == "-":
            seq = seq[::-1].translate(str.maketrans("ACGTacgt", "TGCAtgca"))
        assert c[9] == seq and int(c[11]) == len(seq)


def test_java_double(oracle):
    from parasuite_b200.flush import java_double
    for x in (0.5, 1.0, 1e-4, 1e7, 1 / 3, 0.0, 123456.789, 2.5e-5, 9999999.999, 0.001, 1e-3 - 1e-12, 2 / 3, 100.0, 1e22):
        assert po.java_double_str(x) == java_double(x), x
