"""GPU: the file loops behind the C ABI, streaming and over several contexts in one process --
  * ps_pileup_bam takes the file in windows (carry-in from the decoded records, halo merge of the boundary cluster in
    C++): any window size gives the records of the whole stream (PileupClusters.java:137 streams);
  * ps_create_multi: two contexts (both on cuda:0 here), batches / windows round-robin, host sums -- bit-identical;
  * ps_clust_bam: the six output files of the `clust` tool, byte for byte against the literal Python restatement."""
import random

import numpy as np
import pytest

import py_oracle as po
from helpers import assert_profile_equal, random_genome, random_records, to_py
from parasuite_b200 import PackedReference, ReadBatch, abi
from parasuite_b200.bamio import write_bam, write_fasta
from test_clust_writer_cpu import FILES
from test_gpu_pileup import assert_pileup_equal

pytestmark = pytest.mark.gpu


def _records(seed, kinds, n=4000, length=20000):
    rng = random.Random(seed)
    contigs = random_genome(rng, n_contigs=3, length=length, n_frac=0.004, lower_frac=0.2)
    recs = [r for r in random_records(rng, contigs, n, kinds=kinds, Lrange=(18, 32), flags_special=0.03) if r.pos > 0]
    g = po.Genome(dict(contigs))
    ok = []
    for r in recs:
        try:
            po.pileup(to_py([r]), g, po.SnpDb([]), 1)
            po.profile(to_py([r]), g, 64)
            ok.append(r)
        except po.ReferenceWouldThrow:
            pass
    return contigs, ok


def _files(tmp_path, contigs, recs):
    fa, bam = str(tmp_path / "ref.fa"), str(tmp_path / "reads.bam")
    write_fasta(fa, contigs)
    write_bam(bam, [(n, len(s)) for n, s in contigs], recs)
    return fa, bam


@pytest.mark.parametrize("window", [257, 1000, 10 ** 9])
def test_windowed_pileup_bam_equals_whole_stream(oracle, tmp_path, monkeypatch, window):
    from parasuite_b200.runtime import Context
    contigs, recs = _records(7, ("M", "M", "clip", "indel"))
    fa, bam = _files(tmp_path, contigs, recs)
    monkeypatch.setenv("PARASUITE_B200_WINDOW_READS", str(window))
    ctx = Context(0)
    try:
        ctx.load_fasta(fa)
        with ctx.pileup_bam(bam) as res:
            got = res.fetch(boundary=True)
    finally:
        ctx.close()
    ref = PackedReference.from_contigs(contigs)
    exp = oracle.pileup(ref, ReadBatch.from_records(recs, ref))
    assert_pileup_equal(got, exp, f"window {window}")
    assert len(got["clusters"]) > 50 and got["open_cluster"] is not None


def test_two_contexts_in_one_process(oracle, tmp_path, monkeypatch):
    from parasuite_b200.runtime import MultiContext
    contigs, recs = _records(8, ("M", "clip", "indel"), n=6000)
    fa, bam = _files(tmp_path, contigs, recs)
    monkeypatch.setenv("PARASUITE_B200_WINDOW_READS", "700")
    monkeypatch.setenv("PARASUITE_B200_BATCH_READS", "500")
    m = MultiContext([0, 0])
    try:
        assert m.n_devices == 2
        m.load_fasta(fa)
        prof = m.profile_bam(bam, 64)
        with m.pileup_bam(bam) as res:
            pile = res.fetch(boundary=True)
    finally:
        m.close()
    ref = PackedReference.from_contigs(contigs)
    batch = ReadBatch.from_records(recs, ref)
    assert_profile_equal(prof, oracle.profile(ref, batch, 64), "two contexts, profile")
    assert_pileup_equal(pile, oracle.pileup(ref, batch), "two contexts, pileup")


def test_devices_from_the_environment(monkeypatch):
    from parasuite_b200.runtime import MultiContext
    monkeypatch.setenv("PARASUITE_B200_DEVICES", "0,0,0")
    m = MultiContext()
    try:
        assert m.n_devices == 3
    finally:
        m.close()


@pytest.mark.parametrize("window,min_cov", [(10 ** 9, 1), (400, 2)])
def test_clust_tool_files(oracle, tmp_path, monkeypatch, window, min_cov):
    import gzip
    from parasuite_b200.runtime import Context
    contigs, recs = _records(9, ("M", "M", "clip", "indel", "splice"), n=5000)
    g = po.Genome(dict(contigs))
    try:
        po.clust_files(to_py(recs), g, po.SnpDb([]), min_cov)
    except po.JvmWouldDie:
        recs = [r for r in recs if r.pos + 90 < 20000]
    fa, bam = _files(tmp_path, contigs, recs)
    # a SNP file that hits some T>C sites (chromosome names without "chr", SNPCalling.java:51-54)
    st = po.pileup(to_py(recs), g, po.SnpDb([]), min_cov)
    snps = []
    for c in st.clusters[::4]:
        for pos, _, _ in c.sites[:1]:
            snps.append((c.chrom[3:], pos, "T", "C"))
    vcf = str(tmp_path / "snp.vcf.gz")
    with gzip.open(vcf, "wt") as f:
        f.write("##fileformat=VCFv4.1\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n")
        for ch, pos, r, a in sorted(set(snps)):
            f.write(f"{ch}\t{pos}\t.\t{r}\t{a}\t.\t.\t.\n")
    exp = po.clust_files(to_py(recs), g, po.SnpDb(sorted(set(snps))), min_cov)
    monkeypatch.setenv("PARASUITE_B200_WINDOW_READS", str(window))
    out = str(tmp_path / "clusters.tsv")
    ctx = Context(0)
    try:
        ctx.load_fasta(fa)
        ctr = ctx.clust_bam(bam, out, vcf, min_cov)
    finally:
        ctx.close()
    assert ctr["num_reads_processed"] == len(recs)
    for k, path in FILES.items():
        got = open(path.format(out=out, bam=bam)).read()
        assert got == exp[k], (k, got[:300], exp[k][:300])
    assert exp["pileup"].count("\n") > 40 and "T-C mutations identified as SNPs: 0" not in exp["report"]
