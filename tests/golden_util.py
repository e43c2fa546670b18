"""Loading the committed golden fixture (tests/golden/parasuite_golden_v1.json, made by tests/golden/make_golden.py)."""
import json
import os

import numpy as np

from parasuite_b200 import PackedReference, ReadBatch, Record

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "parasuite_golden_v1.json")


def load():
    g = json.load(open(PATH))
    contigs = [(n, s.encode()) for n, s in g["contigs"]]
    recs = [Record(f, rn, p, c, s.encode(), bytes(q)) for f, rn, p, c, s, q in g["records"]]
    return g, contigs, recs


def batches(contigs, recs):
    ref = PackedReference.from_contigs(contigs)
    return ref, ReadBatch.from_records(recs, ref), ReadBatch.from_records([r for r in recs if r.pos > 0], ref)


def check_profile(got: dict, g: dict, what: str):
    p = g["profile"]
    assert np.array_equal(np.asarray(got["position_conversions"]), np.asarray(p["pos_conv"], dtype=np.int32)), what
    assert np.array_equal(np.asarray(got["quality_per_mismatch"]), np.asarray(p["qual_mm"], dtype=np.int32)), what
    assert np.array_equal(np.asarray(got["quality_per_mismatch_counts"]), np.asarray(p["qual_mm_cnt"], dtype=np.int32)), what
    assert np.array_equal(np.asarray(got["insertions_per_pos"]), np.asarray(p["ins_per_pos"])), what
    assert np.array_equal(np.asarray(got["deletions_per_pos"]), np.asarray(p["del_per_pos"])), what
    assert list(np.asarray(got["counters"])) == p["counters"], what


STRAND = {"+": 0, "-": 1, "+/-": 2}


def check_pileup(got: dict, g: dict, names, what: str):
    """got = Context.pileup / oracle_lib.pileup result (cluster + site records); g = golden."""
    exp = g["pileup"]
    cl, si = got["clusters"], got["sites"]
    assert len(cl) == len(exp["clusters"]), what
    assert got["counters"]["num_reads_processed"] == exp["num_reads_processed"], what
    assert got["counters"]["skipped_due_indel"] == exp["skipped_due_indel"], what
    assert got["counters"]["double_stranded"] == exp["double_stranded"], what
    for k, e in enumerate(exp["clusters"]):
        c = cl[k]
        assert int(c["running_id"]) == e["running_id"] and names[int(c["contig"])] == e["chrom"], (what, k)
        assert (int(c["start"]), int(c["end"])) == (e["start"], e["end"]), (what, k)
        assert bool(c["first_reverse"]) == e["first_reverse"] and int(c["combined_strand"]) == STRAND[e["combined_strand"]], (what, k)
        assert (int(c["num_reads"]), int(c["num_t2c"])) == (e["num_reads"], e["num_t2c"]), (what, k)
        assert [j for j in range(51) if (int(c["mask51"]) >> j) & 1] == e["mask51"], (what, k)
        s = si[int(c["site_begin"]):int(c["site_end"])]
        order = np.argsort(s["order_key"], kind="stable")          # mutationMap insertion order
        assert [[int(s["pos"][j]), int(s["t2c"][j]), int(s["cov"][j])] for j in order] == e["sites_in_insertion_order"], (what, k)
    oc = got["open_cluster"]
    assert (None if oc is None else int(oc["start"])) == exp["open_cluster_start"], what


def check_flush(rows, totals: dict, g: dict, what: str):
    exp = g["pileup"]
    for k, e in enumerate(exp["clusters"]):
        r = rows[k]
        assert bool(r["emitted"]) == e["emitted"], (what, k)
        if not e["emitted"]:
            continue
        assert int(r["num_t2c_sites"]) == e["num_t2c_sites"] and int(r["best_pos"]) == e["best_pos"], (what, k)
        assert float(r["fraction"]) == e["fraction"] and float(r["best_value"]) == e["best_value"], (what, k)
        if e["best_pos"] > 0:
            assert int(r["best_count"]) == e["best_count"], (what, k)
    for key in ("snp_hit", "high_frequent_error", "num_crosslinked_clusters", "num_allele_positions", "allele_positions"):
        assert totals[key] == exp[key], (what, key)
    assert list(totals["allele_frequency_information"]) == exp["allele_frequency_information"], what
