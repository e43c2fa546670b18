"""BASELINE config 1: the reference's own example inputs (examples/references/reference_chr1.fa, the transcripts,
snp_db.vcf.gz; bin/examples.sh:51,65 -- `error <bam> <fasta> 51`, `clust <bam> <fasta> <out> snp_db.vcf.gz 1`) as the
committed fixture tests/golden/config1/ (made by tests/golden/make_config1.py).
CPU: the fixture loads, the C++ oracle agrees with the committed expectations, the native flush + writer reproduce the
six clust files from the oracle's records.  GPU: ps_profile_bam and ps_clust_bam on the BAM + FASTA written from it."""
import gzip
import json
import os

import numpy as np
import pytest

from parasuite_b200 import PackedReference, ReadBatch, Record, abi
from parasuite_b200.bamio import write_bam, write_fasta
from test_clust_writer_cpu import FILES

DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config1")
pytestmark = pytest.mark.skipif(not os.path.exists(abi.lib_path()), reason="library not built")


@pytest.fixture(scope="module")
def cfg(tmp_path_factory):
    d = tmp_path_factory.mktemp("config1")
    text = gzip.open(os.path.join(DIR, "reference_chr1.fa.gz"), "rb").read()
    name = text[1:text.index(b"\n")].split()[0].decode()
    seq = b"".join(text[text.index(b"\n") + 1:].split())
    contigs = [(name, seq)]
    recs = [Record(f, rn, p, c, s.encode(), bytes(q)) for f, rn, p, c, s, q in
            json.loads(gzip.open(os.path.join(DIR, "reads.json.gz"), "rb").read())]
    exp = json.loads(gzip.open(os.path.join(DIR, "expected.json.gz"), "rb").read())
    fa, bam = str(d / "reference_chr1.fa"), str(d / "reads.bam")
    write_fasta(fa, contigs)
    write_bam(bam, [(name, len(seq))], recs)
    return {"dir": d, "contigs": contigs, "recs": recs, "exp": exp, "fa": fa, "bam": bam,
            "vcf": os.path.join(DIR, "snp_db.vcf.gz")}


def check_profile(got, exp):
    p = exp["profile"]
    assert np.array_equal(np.asarray(got["position_conversions"]), np.asarray(p["pos_conv"], dtype=np.int32))
    assert np.array_equal(np.asarray(got["quality_per_mismatch"]), np.asarray(p["qual_mm"], dtype=np.int32))
    assert np.array_equal(np.asarray(got["quality_per_mismatch_counts"]), np.asarray(p["qual_mm_cnt"], dtype=np.int32))
    assert np.array_equal(np.asarray(got["insertions_per_pos"]), np.asarray(p["ins_per_pos"]))
    assert np.array_equal(np.asarray(got["deletions_per_pos"]), np.asarray(p["del_per_pos"]))
    assert list(np.asarray(got["counters"])) == p["counters"]


def test_fixture_shape(cfg):
    assert len(cfg["contigs"][0][1]) == 483300 and cfg["exp"]["n_records"] == len(cfg["recs"]) > 2000
    assert sum("N" in r.cigar for r in cfg["recs"]) > 100 and sum(r.flag & 16 > 0 for r in cfg["recs"]) > 500
    assert cfg["exp"]["n_clusters"] > 50 and cfg["exp"]["clust_files"]["pileup"].count("\n") == cfg["exp"]["n_clusters"] + 1


def test_cpu_oracle_and_native_writer(cfg, oracle):
    from parasuite_b200.flush import ClustWriter, Flush
    ref = PackedReference.from_contigs(cfg["contigs"])
    batch = ReadBatch.from_records(cfg["recs"], ref)
    check_profile(oracle.profile(ref, batch, cfg["exp"]["max_len"]), cfg["exp"])
    res = oracle.pileup(ref, batch)
    out = str(cfg["dir"] / "cpu_clusters.tsv")
    fl = Flush(ref.names, 1, vcf=cfg["vcf"])
    w = ClustWriter(fl, cfg["fa"], out, cfg["bam"])
    w.feed(batch, 0, res["clusters"], res["sites"], None if res["open_cluster"] is None else int(res["open_cluster"]["first_read"]))
    w.finish(res["counters"])
    w.close()
    for k, path in FILES.items():
        assert open(path.format(out=out, bam=cfg["bam"])).read() == cfg["exp"]["clust_files"][k], k


@pytest.mark.gpu
@pytest.mark.parametrize("window", [10 ** 9, 500])
def test_gpu_tools_from_files(cfg, monkeypatch, window):
    from parasuite_b200.runtime import Context
    monkeypatch.setenv("PARASUITE_B200_WINDOW_READS", str(window))
    monkeypatch.setenv("PARASUITE_B200_BATCH_READS", str(min(window, 1 << 22)))
    ctx = Context(0)
    try:
        ctx.load_fasta(cfg["fa"])
        check_profile(ctx.profile_bam(cfg["bam"], cfg["exp"]["max_len"]), cfg["exp"])          # examples.sh:51
        out = str(cfg["dir"] / f"gpu_clusters_{window}.tsv")
        ctr = ctx.clust_bam(cfg["bam"], out, cfg["vcf"], 1)                                     # examples.sh:65
    finally:
        ctx.close()
    assert ctr["double_stranded"] == cfg["exp"]["double_stranded"] and ctr["skipped_due_indel"] == cfg["exp"]["skipped_due_indel"]
    for k, path in FILES.items():
        assert open(path.format(out=out, bam=cfg["bam"])).read() == cfg["exp"]["clust_files"][k], k


# ---- the example pipeline between the aligner and the two tools: comb (examples.sh:58), then error (:51) and clust (:65) ----
@pytest.fixture(scope="module")
def comb(tmp_path_factory, cfg):
    from parasuite_b200.comb import comb_bam
    d = tmp_path_factory.mktemp("config1_comb")
    doc = json.loads(gzip.open(os.path.join(DIR, "comb.json.gz"), "rb").read())

    def recs(rows):
        return ([Record(f, rn, p, "" if c == "*" else c, s.encode(), bytes(q)) for _, f, rn, p, c, s, q, _ in rows],
                [n.encode() for n, *_ in rows])
    g, gn = recs(doc["genomic"])
    t, tn = recs(doc["transcript"])
    name, seq = cfg["contigs"][0]
    gb, tb, ob = str(d / "genomic.bam"), str(d / "transcript.bam"), str(d / "combined.bam")
    write_bam(gb, [(name, len(seq))], g, names=gn)
    write_bam(tb, [(h, ln) for h, ln in doc["transcripts"]], t, sort_order="queryname", names=tn)
    stats = comb_bam(gb, tb, ob)
    return {"doc": doc, "stats": stats, "bam": ob, "dir": d}


def test_comb_lifts_the_transcript_hits_like_the_restatement(comb):
    from parasuite_b200.bamio import read_bam_records
    doc, stats = comb["doc"], comb["stats"]
    assert stats["mapped_reads"] == doc["stats"]["mapped_reads"] and stats["spliced_reads"] == doc["stats"]["spliced_reads"]
    assert stats["missed_transcript_alignments"] == doc["stats"]["missed_transcript_alignments"]
    assert stats["lifted_records"] == doc["stats"]["lifted"] > 1000 and stats["spliced_reads"] > 100
    _, _, got = read_bam_records(comb["bam"])
    assert len(got) == len(doc["combined"])
    for a, (n, f, rn, p, c, s, q, mq) in zip(got, doc["combined"]):
        assert (a["name"], a["flag"], a["rname"], a["pos"], a["cigar"], a["seq"].decode(), list(a["qual"]), a["mapq"]) == \
               (n, f, rn, p, c or "*", s, q, mq if n.startswith("read") else 255)


def test_cpu_tools_on_the_combined_records(comb, cfg, oracle):
    from parasuite_b200.bamio import read_bam_records
    from parasuite_b200.flush import ClustWriter, Flush
    _, _, got = read_bam_records(comb["bam"])
    recs = [Record(a["flag"], a["rname"], a["pos"], "" if a["cigar"] == "*" else a["cigar"], a["seq"], a["qual"]) for a in got]
    ref = PackedReference.from_contigs(cfg["contigs"])
    batch = ReadBatch.from_records(recs, ref)
    check_profile(oracle.profile(ref, batch, 51), {"profile": comb["doc"]["profile"]})
    res = oracle.pileup(ref, batch)
    out = str(comb["dir"] / "cpu_clusters.tsv")
    fl = Flush(ref.names, 1, vcf=cfg["vcf"])
    w = ClustWriter(fl, cfg["fa"], out, comb["bam"])
    w.feed(batch, 0, res["clusters"], res["sites"], None if res["open_cluster"] is None else int(res["open_cluster"]["first_read"]))
    w.finish(res["counters"])
    w.close()
    for k, path in FILES.items():
        assert open(path.format(out=out, bam=comb["bam"])).read() == comb["doc"]["clust_files"][k], k


@pytest.mark.gpu
def test_gpu_tools_on_the_combined_bam(comb, cfg):
    from parasuite_b200.runtime import Context
    ctx = Context(0)
    try:
        ctx.load_fasta(cfg["fa"])
        check_profile(ctx.profile_bam(comb["bam"], 51), {"profile": comb["doc"]["profile"]})           # examples.sh:51
        out = str(comb["dir"] / "gpu_clusters.tsv")
        ctx.clust_bam(comb["bam"], out, cfg["vcf"], 1)                                                 # examples.sh:65
    finally:
        ctx.close()
    for k, path in FILES.items():
        assert open(path.format(out=out, bam=comb["bam"])).read() == comb["doc"]["clust_files"][k], k
