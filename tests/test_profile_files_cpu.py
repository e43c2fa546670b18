"""CPU: the output files of the `error` tool (ErrorProfiling.java:410-591).  The vectorised writer of the product
(parasuite_b200/profile_files.py) against the statement-by-statement restatement in the oracle, both fed from the
Python oracle's loop state; plus hand-checked lines on the SURVEY known-answer reads."""
import math
import random

import numpy as np
import pytest

import py_oracle as po
from helpers import kat_records, py_profile_dict, random_genome, random_records, to_py
from kat_vectors import KAT_MAXLEN, KAT_REF, PROFILE_KATS
from parasuite_b200.flush import java_double
from parasuite_b200.profile_files import profile_file_texts, write_profile_files, write_profile_files_native

EXACT = ("errorprofile", "errorprofile.vcf", "qualityPerMismatch", "indels", "indelprofile", "averaged_t2c_epr")


def state_to_result(st: po.ProfileState, infer_q: bool):
    res = py_profile_dict(st)
    if infer_q:
        h = np.zeros((st.max_len, 256), dtype=np.int64)
        for i, d in enumerate(st.qual_hist):
            for q, c in d.items():
                h[i, q & 0xFF] = c
        res["quality_hist"] = h
    return res


def survivors(recs, genome, max_len):
    ok = []
    for r in recs:
        try:
            po.profile(to_py([r]), genome, max_len)
            ok.append(r)
        except po.ReferenceWouldThrow:
            pass
    return ok


@pytest.mark.parametrize("seed,kinds,infer_q", [(1, ("M",), False), (2, ("M", "clip", "indel"), False), (3, ("M", "M", "indel"), True)])
def test_writer_matches_literal_restatement(seed, kinds, infer_q):
    rng = random.Random(seed)
    contigs = random_genome(rng, n_contigs=2, length=3000, n_frac=0.01, lower_frac=0.1)
    g = po.Genome(dict(contigs))
    recs = survivors(random_records(rng, contigs, 1500, kinds=kinds, Lrange=(15, 40), flags_special=0.03), g, 64)
    if infer_q:      # -q touches qual[i] for every i < ml: a D read would kill the JVM (Q7); keep M and I reads
        recs = [r for r in recs if "D" not in r.cigar and "N" not in r.cigar]
    st = po.profile(to_py(recs), g, 64, infer_qual=infer_q)
    exp = po.profile_outputs(st, infer_q, java_double)
    got = profile_file_texts(state_to_result(st, infer_q), infer_q)
    for k in EXACT:
        assert got[k] == exp[k], k
    if infer_q:
        gl, el = got["qualities"].splitlines(), exp["qualities"].splitlines()
        assert len(gl) == len(el) == 64
        for a, b in zip(gl, el):
            (am, asd), (bm, bsd) = a.split("\t"), b.split("\t")
            assert am == bm                                             # the mean is exact
            if bsd == "NaN":
                assert asd == "NaN"
            else:                                                       # the SD is a sequential FP sum in Java (tolerance 1e-12)
                assert math.isclose(float(asd), float(bsd), rel_tol=1e-12, abs_tol=1e-12)


def test_kat_lines(tmp_path):
    """All SURVEY 8(c) profile reads in one run: hand-checkable totals."""
    recs = []
    for kid in sorted(PROFILE_KATS):
        recs += kat_records(PROFILE_KATS[kid]["reads"])
    g = po.Genome({"chr1": KAT_REF.encode()})
    st = po.profile(to_py(recs), g, KAT_MAXLEN)
    res = state_to_result(st, False)
    t = write_profile_files(str(tmp_path / "x.bam"), res)
    for suffix in ("errorprofile", "errorprofile.vcf", "qualityPerMismatch", "indels", "indelprofile", "qualities"):
        assert (tmp_path / f"x.bam.{suffix}").read_text() == t[suffix]
    pc = np.asarray(res["position_conversions"], dtype=np.int64)
    rows = t["errorprofile"].splitlines()
    assert len(rows) == 4 and all(len(r.split("\t")) == 5 for r in rows)           # 4 values + trailing tab
    tot = pc.sum(axis=0)
    for j in range(4):
        vals = [float(x) for x in rows[j].split("\t")[:4]]
        assert math.isclose(sum(vals), 1.0, rel_tol=1e-12)
        assert vals[0] == tot[j, 0] / tot[j].sum()
    vcf = t["errorprofile.vcf"].split("\n")
    assert vcf[0] == f"A\tA\t{java_double(float(tot[0, 0]))}" and vcf[4] == ""    # blank line after each reference base
    assert t["indelprofile"].count("\t") == 1 and not t["indelprofile"].endswith("\n")
    assert len(t["indels"].splitlines()) == KAT_MAXLEN


def test_empty_run_prints_nan():
    res = {"position_conversions": np.zeros((5, 4, 4), np.int32), "quality_per_mismatch": np.zeros((4, 4), np.int32),
           "quality_per_mismatch_counts": np.zeros((4, 4), np.int32), "insertions_per_pos": np.zeros(5),
           "deletions_per_pos": np.zeros(5), "counters": np.zeros(8, np.int32)}
    t = profile_file_texts(res)
    assert t["errorprofile"] == ("NaN\t" * 4 + "\n") * 4 and t["indelprofile"] == "0.0\t0.0"
    st = po.ProfileState(5)
    assert po.profile_outputs(st, False, java_double)["errorprofile"] == t["errorprofile"]


@pytest.mark.parametrize("seed,infer_q", [(5, False), (6, True)])
def test_native_writer_writes_the_same_files(tmp_path, seed, infer_q):
    """csrc/profile_writer.cpp (ps_profile_write_files) against the Python writer: the five exact files byte for byte,
    .qualities with the exact mean and the SD to 1e-12 (the two sum the histogram in different orders)."""
    rng = random.Random(seed)
    contigs = random_genome(rng, n_contigs=2, length=3000, n_frac=0.01, lower_frac=0.1)
    g = po.Genome(dict(contigs))
    recs = survivors(random_records(rng, contigs, 1500, kinds=("M", "M", "indel") if not infer_q else ("M", "M"), Lrange=(15, 40),
                                    flags_special=0.03), g, 64)
    st = po.profile(to_py(recs), g, 64, infer_qual=infer_q)
    res = state_to_result(st, infer_q)
    want = profile_file_texts(res, infer_q)
    bam = str(tmp_path / "x.bam")
    avg = write_profile_files_native(bam, res, infer_q)
    assert java_double(avg) == want["averaged_t2c_epr"]
    for k in EXACT:
        if k == "averaged_t2c_epr":
            continue
        assert open(f"{bam}.{k}").read() == want[k], k
    got_q = open(f"{bam}.qualities").read()
    if not infer_q:
        assert got_q == ""
    else:
        gl, wl = got_q.splitlines(), want["qualities"].splitlines()
        assert len(gl) == len(wl) == 64
        for a, b in zip(gl, wl):
            (am, asd), (bm, bsd) = a.split("\t"), b.split("\t")
            assert am == bm
            if bsd != "NaN":
                assert abs(float(asd) - float(bsd)) <= 1e-12 * max(1.0, abs(float(bsd)))
            else:
                assert asd == "NaN"
