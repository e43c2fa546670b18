"""GPU parity through the file-level entry points (ps_reference_load_fasta, ps_profile_bam, ps_pileup_bam): the two
tool loops from a coordinate-sorted BAM + FASTA, against the CPU oracle on the same records."""
import random

import numpy as np
import pytest

import py_oracle as po
from helpers import assert_profile_equal, random_genome, random_records, to_py
from parasuite_b200 import PackedReference, ReadBatch, abi
from parasuite_b200.bamio import write_bam, write_fasta
from test_gpu_pileup import assert_pileup_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from parasuite_b200.runtime import Context
    c = Context(0)
    yield c
    c.close()


def _files(tmp_path, contigs, recs):
    fa, bam = str(tmp_path / "ref.fa"), str(tmp_path / "reads.bam")
    write_fasta(fa, contigs)
    write_bam(bam, [(n, len(s)) for n, s in contigs], recs)
    return fa, bam


@pytest.mark.parametrize("batch_reads", [0, 300])
def test_profile_bam(ctx, oracle, tmp_path, monkeypatch, batch_reads):
    rng = random.Random(41)
    contigs = random_genome(rng, n_contigs=3, length=6000, n_frac=0.01, lower_frac=0.1)
    recs = random_records(rng, contigs, 3000, kinds=("M", "M", "clip", "indel"), Lrange=(15, 40), flags_special=0.05)
    g = po.Genome(dict(contigs))
    ok = []
    for r in recs:                      # keep the records the JVM survives
        try:
            po.profile(to_py([r]), g, 64)
            ok.append(r)
        except po.ReferenceWouldThrow:
            pass
    fa, bam = _files(tmp_path, contigs, ok)
    if batch_reads:
        monkeypatch.setenv("PARASUITE_B200_BATCH_READS", str(batch_reads))     # many batches: slab / staging reuse
    ctx.load_fasta(fa)
    got = ctx.profile_bam(bam, 64)
    ref = PackedReference.from_contigs(contigs)
    exp = oracle.profile(ref, ReadBatch.from_records(ok, ref), 64)
    assert_profile_equal(got, exp, f"profile_bam batch_reads={batch_reads}")
    assert int(got["counters"][0]) > 2000


def test_pileup_bam(ctx, oracle, tmp_path):
    rng = random.Random(42)
    contigs = random_genome(rng, n_contigs=2, length=5000, n_frac=0.005, lower_frac=0.1)
    recs = [r for r in random_records(rng, contigs, 2500, kinds=("M", "clip", "indel"), Lrange=(15, 30)) if r.pos > 0]
    g = po.Genome(dict(contigs))
    ok = []
    for r in recs:
        try:
            po.pileup(to_py([r]), g, po.SnpDb([]), 1)
            ok.append(r)
        except po.ReferenceWouldThrow:
            pass
    fa, bam = _files(tmp_path, contigs, ok)
    ctx.load_fasta(fa)
    with ctx.pileup_bam(bam) as res:
        got = res.fetch(boundary=False)
    ref = PackedReference.from_contigs(contigs)
    exp = oracle.pileup(ref, ReadBatch.from_records(ok, ref))
    exp2 = {k: exp[k] for k in ("clusters", "sites", "counters")}
    exp2["open_cluster"] = None
    got["open_cluster"] = None
    assert_pileup_equal(got, exp2, "pileup_bam")
    assert len(got["clusters"]) > 20


def test_unsorted_bam_is_refused(ctx, tmp_path):
    contigs = [("chr1", b"ACGT" * 100)]
    fa, bam = str(tmp_path / "r.fa"), str(tmp_path / "u.bam")
    write_fasta(fa, contigs)
    write_bam(bam, [("chr1", 400)], [], sort_order="queryname")
    ctx.load_fasta(fa)
    with pytest.raises(abi.PsError) as e:
        ctx.profile_bam(bam, 51)
    assert e.value.status == abi.PS_ERR_UNSORTED


def test_bam_without_fasta_is_a_state_error(tmp_path):
    from parasuite_b200.runtime import Context
    c = Context(0)
    try:
        with pytest.raises(abi.PsError) as e:
            c.profile_bam(str(tmp_path / "x.bam"), 51)
        assert e.value.status == abi.PS_ERR_STATE
    finally:
        c.close()


def test_error_tool_from_files(tmp_path, oracle):
    """ps_error_bam: the whole `error` tool -- record loop on the GPU, the six output files natively -- against the Python
    writer fed with the oracle's arrays."""
    from parasuite_b200.profile_files import profile_file_texts
    from parasuite_b200.runtime import Context
    rng = random.Random(31)
    contigs = random_genome(rng, n_contigs=2, length=4000, n_frac=0.01)
    g = po.Genome(dict(contigs))
    recs = []
    for r in random_records(rng, contigs, 3000, kinds=("M", "M", "M", "indel"), Lrange=(18, 45), flags_special=0.03):
        try:
            po.profile(to_py([r]), g, 51)
            recs.append(r)
        except po.ReferenceWouldThrow:
            pass
    fa, bam = str(tmp_path / "r.fa"), str(tmp_path / "r.bam")
    write_fasta(fa, contigs)
    write_bam(bam, [(n, len(s)) for n, s in contigs], recs)
    ref = PackedReference.from_contigs(contigs)
    exp = oracle.profile(ref, ReadBatch.from_records(recs, ref), 51)
    want = profile_file_texts(exp, False)
    ctx = Context(0)
    try:
        ctx.load_fasta(fa)
        ctr = ctx.error_bam(bam, 51)
    finally:
        ctx.close()
    assert list(ctr) == [int(x) for x in exp["counters"]]
    for k in ("errorprofile", "errorprofile.vcf", "qualityPerMismatch", "indels", "indelprofile", "qualities"):
        assert open(f"{bam}.{k}").read() == want[k], k
