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
