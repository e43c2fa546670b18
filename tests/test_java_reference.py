"""Pins the oracle on the reference itself whenever a JVM is present: runs bin/parasuite.jar (`error`, `clust`) on BAM +
FASTA files written here and compares its output files, byte for byte, with what the Python restatement
(oracle/py_oracle.py) says they must be.  No JRE is in the build image or on the GPU box, so these tests skip there
with the reason; on a machine with `java` and the jar (PARASUITE_JAR=...) they are the missing link of SURVEY 8c.

`qualities` (the -q file) is compared on the mean column only: the reference sums the standard deviation over a linked
list in arrival order (ErrorProfiling.java:575-589), whose last digits depend on that order."""
import gzip
import json
import os
import random

import pytest

import java_ref
import py_oracle as po
from helpers import kat_records, random_genome, random_records, to_py
from kat_vectors import KAT_MAXLEN, KAT_REF, PILEUP_READS, PILEUP_REF, PROFILE_KATS
from parasuite_b200 import Record
from parasuite_b200.bamio import write_bam, write_fasta
from parasuite_b200.flush import java_double

JAVA, JAR = java_ref.find()
pytestmark = pytest.mark.skipif(JAVA is None, reason=f"reference jar not runnable here: {JAR}")

EXACT = ("errorprofile", "errorprofile.vcf", "qualityPerMismatch", "indels", "indelprofile")


def survivors(recs, genome, max_len, pileup=False):
    ok = []
    for r in recs:
        try:
            po.profile(to_py([r]), genome, max_len)
            if pileup:
                po.pileup(to_py([r]), genome, po.SnpDb([]), 1)
            ok.append(r)
        except po.ReferenceWouldThrow:
            pass
    return ok


def files_for(tmp_path, contigs, recs, tag):
    fa, bam = str(tmp_path / f"{tag}.fa"), str(tmp_path / f"{tag}.bam")
    write_fasta(fa, contigs)
    write_bam(bam, [(n, len(s)) for n, s in contigs], recs)
    return fa, bam


def check_error(tmp_path, contigs, recs, max_len, infer_q, tag):
    fa, bam = files_for(tmp_path, contigs, recs, tag)
    got, _, proc = java_ref.run_error(JAVA, JAR, bam, fa, max_len, infer_q)
    assert proc.returncode == 0, proc.stderr[-2000:]
    st = po.profile(to_py(recs), po.Genome(dict(contigs)), max_len, infer_qual=infer_q)
    exp = po.profile_outputs(st, infer_q, java_double)
    for k in EXACT:
        assert got[k] == exp[k], k
    if infer_q:
        gm = [l.split("\t")[0] for l in got["qualities"].splitlines()]
        em = [l.split("\t")[0] for l in exp["qualities"].splitlines()]
        assert gm == em


def test_error_tool_on_known_answer_reads(tmp_path):
    recs = []
    for kid in sorted(PROFILE_KATS):
        recs += kat_records(PROFILE_KATS[kid]["reads"])
    contigs = [("chr1", KAT_REF.encode())]
    recs = sorted(survivors(recs, po.Genome(dict(contigs)), KAT_MAXLEN), key=lambda r: r.pos)
    check_error(tmp_path, contigs, recs, KAT_MAXLEN, False, "kat")


@pytest.mark.parametrize("seed,kinds,infer_q", [(11, ("M",), False), (12, ("M", "clip", "indel", "splice"), False), (13, ("M", "M", "ins"), True)])
def test_error_tool_on_random_records(tmp_path, seed, kinds, infer_q):
    rng = random.Random(seed)
    contigs = random_genome(rng, n_contigs=3, length=5000, n_frac=0.01, lower_frac=0.1)
    g = po.Genome(dict(contigs))
    kinds = tuple(k for k in kinds if k in ("M", "clip", "indel"))
    recs = survivors(random_records(rng, contigs, 4000, kinds=kinds, Lrange=(15, 50), flags_special=0.03), g, 51)
    if infer_q:
        recs = [r for r in recs if "D" not in r.cigar and "N" not in r.cigar]
    check_error(tmp_path, contigs, recs, 51, infer_q, f"rand{seed}")


SHIPPED_VCF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config1", "snp_db.vcf.gz")
SHIPPED_SNPS = [("1", 17700, "T", "C"), ("1", 21050, "G", "A"), ("1", 29000, "T", "C")]      # its three rows


def check_clust(tmp_path, contigs, recs, min_cov, tag):
    """The SNP file is the reference's own examples/references/snp_db.vcf.gz with its .tbi (SNPCalling opens a TabixReader,
    which needs the index; the fixture directory keeps both)."""
    fa, bam = files_for(tmp_path, contigs, recs, tag)
    out = str(tmp_path / f"{tag}.clusters")
    got, _, proc = java_ref.run_clust(JAVA, JAR, bam, fa, out, SHIPPED_VCF, min_cov)
    assert proc.returncode == 0, proc.stderr[-2000:]
    exp = po.clust_files(to_py(recs), po.Genome(dict(contigs)), po.SnpDb(SHIPPED_SNPS), min_cov)
    for k in java_ref.CLUST_FILES:
        assert got[k] == exp[k], k
    return exp


def test_clust_tool_on_known_answer_reads(tmp_path):
    recs = kat_records(PILEUP_READS)
    check_clust(tmp_path, [("chr1", PILEUP_REF.encode())], recs, 1, "pkat")


@pytest.mark.parametrize("seed,min_cov", [(21, 1), (22, 3)])
def test_clust_tool_on_random_records(tmp_path, seed, min_cov):
    rng = random.Random(seed)
    contigs = random_genome(rng, n_contigs=2, length=60000, n_frac=0.0, lower_frac=0.2)
    seq = bytearray(contigs[0][1])
    for p in (17700, 29000):                      # a T under the shipped T>C rows, and reads that show C there
        seq[p - 1] = ord("T")
    contigs[0] = (contigs[0][0], bytes(seq))
    g = po.Genome(dict(contigs))
    recs = random_records(rng, contigs, 3000, kinds=("M", "M", "clip"), Lrange=(20, 40), err=0.08)
    for p in (17700, 29000):
        for k in range(4):
            s0 = p - 5 - k
            b = bytearray(seq[s0 - 1:s0 - 1 + 30].upper())
            b[p - s0] = ord("C")
            recs.append(Record(0, contigs[0][0], s0, "30M", bytes(b), bytes([30] * 30)))
    order = {n: i for i, (n, _) in enumerate(contigs)}
    recs = sorted(survivors(recs, g, 64, pileup=True), key=lambda r: (order[r.rname], r.pos))
    exp = check_clust(tmp_path, contigs, recs, min_cov, f"prand{seed}")
    assert "T-C mutations identified as SNPs: 0" not in exp["report"]


def test_config1_fixture(tmp_path):
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config1")
    text = gzip.open(os.path.join(d, "reference_chr1.fa.gz"), "rb").read()
    name = text[1:text.index(b"\n")].split()[0].decode()
    seq = b"".join(text[text.index(b"\n") + 1:].split())
    recs = [Record(f, rn, p, c, s.encode(), bytes(q)) for f, rn, p, c, s, q in
            json.loads(gzip.open(os.path.join(d, "reads.json.gz"), "rb").read())]
    exp = json.loads(gzip.open(os.path.join(d, "expected.json.gz"), "rb").read())
    fa, bam = files_for(tmp_path, [(name, seq)], recs, "config1")
    out = str(tmp_path / "config1.clusters")
    got, _, proc = java_ref.run_clust(JAVA, JAR, bam, fa, out, SHIPPED_VCF, 1)
    assert proc.returncode == 0, proc.stderr[-2000:]
    for k in java_ref.CLUST_FILES:
        assert got[k] == exp["clust_files"][k], k


def test_comb_tool(tmp_path):
    """`comb` (CombineGenomeTranscript): the jar's combined BAM against the Python restatement, record by record."""
    from parasuite_b200.bamio import read_bam_records
    from test_liftover_cpu import random_hit_cigar, random_transcript
    rng = random.Random(3)
    genome = [("chr1", 200000), ("chr2", 150000)]
    transcripts = list(dict.fromkeys(random_transcript(rng, chrom=rng.choice(["1", "2", "7"])) for _ in range(20)))
    g_recs = [Record(rng.choice([0, 16]), "chr1", 100 + 37 * k, "30M", bytes(rng.choice(b"ACGT") for _ in range(30)), bytes([30] * 30))
              for k in range(200)]
    g_names = [b"g%d" % k for k in range(len(g_recs))]
    t_recs, t_names = [], []
    for k in range(300):
        tr = rng.choice(transcripts)
        L = rng.randint(18, 45)
        cigar, R = random_hit_cigar(rng, L)
        t_recs.append(Record(rng.choice([0, 16]), tr[0], rng.randint(1, max(1, tr[1] - R + 3)), cigar,
                             bytes(rng.choice(b"ACGT") for _ in range(L)), bytes(rng.randint(2, 40) for _ in range(L))))
        t_names.append(b"t%04d" % k)
    gb, tb, ob = str(tmp_path / "g.bam"), str(tmp_path / "t.bam"), str(tmp_path / "o.bam")
    write_bam(gb, genome, g_recs, names=g_names)
    write_bam(tb, [(t[0], t[1]) for t in transcripts], t_recs, sort_order="queryname", names=t_names)
    proc, _ = java_ref.run_comb(JAVA, JAR, gb, tb, ob)
    assert proc.returncode == 0, proc.stderr[-2000:]
    _, _, got = read_bam_records(ob)

    def as_dicts(recs, names):
        return [{"name": n.decode(), "flag": r.flag, "rname": r.rname, "pos": r.pos, "cigar": r.cigar, "seq": r.seq, "qual": bytes(r.qual),
                 "mapq": 255} for r, n in zip(recs, names)]
    want, _ = po.combine([n for n, _ in genome], "coordinate", as_dicts(g_recs, g_names), as_dicts(t_recs, t_names))
    assert len(got) == len(want)
    for a, b in zip(got, want):
        for k in ("name", "flag", "rname", "pos", "mapq", "seq", "qual"):
            assert a[k] == b[k], (k, a, b)
        assert a["cigar"] == (b["cigar"] or "*")
