"""CPU: the `comb` tool (CombineGenomeTranscript.java:36-666) -- transcript hits lifted to genomic coordinates (cigars
with N across introns), merged with the genomic hits.  Hand-derived known answers for the per-hit arithmetic (worked
from the Java by hand, including its quirks), the native implementation (csrc/liftover.cpp) against the literal Python
restatement (oracle/py_oracle.py) on random transcripts, and the whole tool on BAM files."""
import os
import random

import pytest

import py_oracle as po
from parasuite_b200 import Record, abi
from parasuite_b200.bamio import read_bam_records, write_bam
from parasuite_b200.comb import comb_bam, liftover_hit

pytestmark = pytest.mark.skipif(not os.path.exists(abi.lib_path()), reason="library not built")

T_PLUS = "GENE|TX|1|100;300;500|149;349;549|1"          # three exons of 50 nt, introns of 150 nt
T_MINUS = "GENE|TX|1|100;300;500|149;349;549|-1"
# (transcript, alignment start, alignment end, read length, cigar) -> (new start, new cigar, missed)
KATS = [
    ((T_PLUS, 10, 45, 36, "36M"), (109, "36M", 0)),                     # inside exon 1: the hit's own cigar (:232-238)
    ((T_PLUS, 40, 75, 36, "36M"), (139, "11M150N25M", 0)),              # across the first junction
    ((T_PLUS, 45, 150, 106, "106M"), (144, "6M150N50M150N50M", 0)),     # a whole exon covered (:344-352)
    ((T_PLUS, 45, 155, 111, "111M"), (144, "6M150N50M150N50M", 0)),     # a hit running past the last exon: the walk just ends
    ((T_MINUS, 10, 45, 36, "36M"), (505, "36M", 0)),                    # minus strand: transcript 1 = genomic 549
    ((T_MINUS, 40, 75, 36, "36M"), (325, "25M150N11M", 0)),
    ((T_PLUS, 40, 75, 35, "20M1D15M"), (139, "", 1)),                   # indel + junction: start kept, cigar empty (:294)
    ((T_MINUS, 40, 75, 35, "20M1D15M"), (-1, "", 1)),                   # minus strand: nothing kept (:447)
    ((T_PLUS, 10, 46, 36, "20M1D16M"), (109, "20M1D16M", 0)),           # indel inside one exon is fine
    (("G|T|1|900;1000|950;1100|1", 5, 40, 36, "36M"), (1004, "36M", 0)),   # exon lists sorted as STRINGS (:172-175)
    ((T_PLUS, 200, 235, 36, "36M"), (-1, "", 0)),                       # behind the transcript's end
    (("G|T|1|100;150|149;199|1", 40, 75, 36, "36M"), (139, "11M", 0)),  # adjacent exons: intron length 0 stops the walk (:381)
    (("G|T|1|100|149|0", 10, 45, 36, "36M"), (-1, "", 0)),              # strand neither 1 nor -1
]


@pytest.mark.parametrize("args,want", KATS)
def test_known_answers(args, want):
    assert po.liftover_hit(*args) == want
    assert liftover_hit(*args) == want


def test_malformed_names_kill_the_tool():
    for name in ("G|T|1|100|149", "G|T|1|100;x|149;200|1", "G|T|1|100;300|149|1"):
        with pytest.raises(po.ReferenceWouldThrow):
            po.liftover_hit(name, 40, 75, 36, "36M")
        with pytest.raises(abi.ReferenceWouldThrow):
            liftover_hit(name, 40, 75, 36, "36M")


def random_transcript(rng, chrom="1"):
    n = rng.randint(1, 6)
    pos = rng.randint(1, 20000)
    starts, ends = [], []
    for _ in range(n):
        ln = rng.randint(5, 120)
        starts.append(pos)
        ends.append(pos + ln - 1)
        pos += ln + rng.choice([0, 1, 1, 30, 500, 5000])
    return f"G{rng.randint(0, 5)}|T{rng.randint(0, 99)}|{chrom}|{';'.join(map(str, starts))}|{';'.join(map(str, ends))}|{rng.choice(['1', '-1'])}", \
        sum(e - s + 1 for s, e in zip(starts, ends))


def random_hit_cigar(rng, L):
    kind = rng.random()
    if kind < 0.7 or L < 12:
        return f"{L}M", L
    a = rng.randint(3, L - 6)
    if kind < 0.85:
        return f"{a}M1I{L - a - 1}M", L - 1
    return f"{a}M2D{L - a}M", L + 2


def test_random_hits_native_equals_restatement():
    rng = random.Random(42)
    spliced = 0
    for _ in range(4000):
        name, tlen = random_transcript(rng)
        L = rng.randint(15, 60)
        cigar, R = random_hit_cigar(rng, L)
        start = rng.randint(1, max(1, tlen + 5))
        args = (name, start, start + R - 1, L, cigar)
        want = po.liftover_hit(*args)
        assert liftover_hit(*args) == want, args
        spliced += "N" in want[1]
    assert spliced > 300


def comp_rev(seq: bytes) -> bytes:
    return seq.translate(bytes.maketrans(b"ACGT", b"TGCA"))[::-1]


def test_whole_tool_on_bam_files(tmp_path):
    """Genomic BAM (coordinate sorted) + transcript BAM (queryname sorted, multi-hit reads, secondary hits, hits that do not
    lift, a contig the genome lacks, MT) -> combined BAM: the records of the Python restatement, in its order."""
    rng = random.Random(7)
    genome = [("chr1", 200000), ("chr2", 150000), ("chrMT", 16000), ("chrM", 16000)]
    transcripts = [random_transcript(rng, chrom=rng.choice(["1", "1", "2", "MT", "7"])) for _ in range(30)]
    g_recs, g_names = [], []
    for k in range(300):
        L = rng.randint(20, 40)
        c = rng.choice(genome[:2])
        g_recs.append(Record(rng.choice([0, 16]), c[0], rng.randint(1, c[1] - 50), f"{L}M", bytes(rng.choice(b"ACGT") for _ in range(L)),
                             bytes(rng.randint(2, 40) for _ in range(L))))
    order = {n: i for i, (n, _) in enumerate(genome)}
    g_recs.sort(key=lambda r: (order[r.rname], r.pos))
    g_names = [b"g%d" % k for k in range(len(g_recs))]
    t_recs, t_names = [], []
    for k in range(400):
        name = b"t%04d" % k
        n_hits = rng.choice([1, 1, 1, 2, 3])
        same = rng.random() < 0.5                     # several hits of one transcript position (they lift to one start)
        tr0 = rng.choice(transcripts)
        L = rng.randint(18, 45)
        cigar, R = random_hit_cigar(rng, L)
        start0 = rng.randint(1, max(1, tr0[1] - R + 3))
        seq = bytes(rng.choice(b"ACGTN") for _ in range(L))
        qual = bytes(rng.randint(2, 40) for _ in range(L))
        for h in range(n_hits):
            tr, start = (tr0, start0) if (same or h == 0) else (rng.choice(transcripts), rng.randint(1, 50))
            flag = (0 if h == 0 else 0x100) | rng.choice([0, 16])
            t_recs.append(Record(flag, tr[0], start, cigar, seq, qual))
            t_names.append(name)
        if rng.random() < 0.05:                       # an unplaced hit of the same read
            t_recs.append(Record(4, "*", 0, "*", seq, qual))
            t_names.append(name)
    gb, tb, ob = str(tmp_path / "genomic.bam"), str(tmp_path / "transcript.bam"), str(tmp_path / "combined.bam")
    write_bam(gb, genome, g_recs, names=g_names)
    write_bam(tb, [(t[0], t[1]) for t in dict.fromkeys(transcripts)], t_recs, sort_order="queryname", names=t_names)
    stats = comb_bam(gb, tb, ob)
    text, refs, got = read_bam_records(ob)
    assert refs == genome and "SO:coordinate" in text

    def as_dicts(recs, names):
        return [{"name": n.decode(), "flag": r.flag, "rname": r.rname, "pos": r.pos, "cigar": r.cigar, "seq": r.seq, "qual": bytes(r.qual),
                 "mapq": 255} for r, n in zip(recs, names)]
    want, wstats = po.combine([n for n, _ in genome], "coordinate", as_dicts(g_recs, g_names), as_dicts(t_recs, t_names))
    assert stats["mapped_reads"] == wstats["mapped_reads"] and stats["spliced_reads"] == wstats["spliced_reads"]
    assert stats["missed_transcript_alignments"] == wstats["missed_transcript_alignments"] and stats["lifted_records"] == wstats["lifted"]
    assert stats["lifted_records"] > 100 and stats["spliced_reads"] > 20
    assert len(got) == len(want)
    for a, b in zip(got, want):
        for k in ("name", "flag", "rname", "pos", "mapq", "seq", "qual"):
            assert a[k] == b[k], (k, a, b)
        assert a["cigar"] == (b["cigar"] or "*")
    # the lifted records of minus-strand transcripts: strand flipped, bases reverse-complemented, qualities as they were
    src = {n.decode(): r for r, n in zip(t_recs, t_names) if not (r.flag & 0x100) and r.rname != "*"}
    flipped = 0
    for a in got:
        if a["name"].startswith("t") and src[a["name"]].rname.endswith("|-1"):
            s = src[a["name"]]
            assert a["seq"] == comp_rev(s.seq) and a["qual"] == bytes(s.qual) and (a["flag"] ^ s.flag) & 16
            flipped += 1
    assert flipped > 20
    # the lifted file goes straight into the batcher (N cigars and all)
    from parasuite_b200.bamio import BamBatcher                        # noqa: F401  (import check only: no GPU here)


def test_unsorted_transcript_bam_is_refused(tmp_path):
    gb, tb, ob = str(tmp_path / "g.bam"), str(tmp_path / "t.bam"), str(tmp_path / "o.bam")
    write_bam(gb, [("chr1", 1000)], [Record(0, "chr1", 5, "10M", b"ACGTACGTAC", bytes([30] * 10))])
    write_bam(tb, [(T_PLUS, 150)], [Record(0, T_PLUS, 5, "10M", b"ACGTACGTAC", bytes([30] * 10))], sort_order="coordinate")
    with pytest.raises(abi.PsError) as e:
        comb_bam(gb, tb, ob)
    assert e.value.status == abi.PS_ERR_UNSORTED
