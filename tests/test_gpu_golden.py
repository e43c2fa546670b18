"""GPU: the CUDA path through the C ABI against the committed golden vectors (no oracle at run time), from SoA batches
and from the BAM + FASTA files written from the same records."""
import pytest

import golden_util as gu

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from parasuite_b200.runtime import Context
    c = Context(0)
    yield c
    c.close()


def test_batches(ctx):
    from parasuite_b200.flush import Flush
    g, contigs, recs = gu.load()
    ref, batch, pbatch = gu.batches(contigs, recs)
    ctx.upload_reference(ref)
    gu.check_profile(ctx.profile(batch, g["max_read_length"]), g, "CUDA profile")
    res = ctx.pileup(pbatch)
    gu.check_pileup(res, g, ref.names, "CUDA pileup")
    fl = Flush(ref.names, g["min_read_coverage"], snps=[tuple(s) for s in g["snps"]])
    gu.check_flush(fl.clusters(res["clusters"], res["sites"]), fl.totals(), g, "CUDA pileup + native flush")


def test_files(ctx, tmp_path):
    from parasuite_b200.bamio import write_bam, write_fasta
    g, contigs, recs = gu.load()
    fa, bam, pbam = str(tmp_path / "g.fa"), str(tmp_path / "g.bam"), str(tmp_path / "p.bam")
    write_fasta(fa, contigs)
    sq = [(n, len(s)) for n, s in contigs]
    write_bam(bam, sq, recs)
    write_bam(pbam, sq, [r for r in recs if r.pos > 0])
    ctx.load_fasta(fa)
    gu.check_profile(ctx.profile_bam(bam, g["max_read_length"]), g, "ps_profile_bam")
    with ctx.pileup_bam(pbam) as h:
        res = h.fetch(boundary=True)
    gu.check_pileup(res, g, [n for n, _ in contigs], "ps_pileup_bam")
