"""GPU parity: batches whose reads differ in length or cigar (adapter-trimmed PAR-CLIP reads as an aligner leaves
them) go through the repack kernel + the per-read-length instantiation of the fast profile kernel (profile.cu,
profile_fast.cuh RG).  Bit-exact against the CPU oracle; the debug word says how many reads really took that path."""
import ctypes as C
import random

import numpy as np
import pytest

import py_oracle as po
from helpers import assert_profile_equal, random_genome, random_records, to_py
from parasuite_b200 import PackedReference, ReadBatch, Record, abi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from parasuite_b200.runtime import Context
    c = Context(0)
    c.lib.ps_debug_word.restype = C.c_ulonglong
    c.lib.ps_debug_word.argtypes = [C.c_void_p, C.c_int]
    yield c
    c.close()


def survivors(recs, g, max_len):
    ok, bad = [], []
    for r in recs:
        try:
            po.profile(to_py([r]), g, max_len)
            ok.append(r)
        except po.ReferenceWouldThrow:
            bad.append(r)
    return ok, bad


def profile_counting_fast(ctx, ref, batch, max_len):
    ctx.upload_reference(ref)
    ctx.profile_begin(max_len)
    ctx.profile_batch(batch)
    fast = int(ctx.lib.ps_debug_word(ctx.h, 1))
    return ctx.profile_end(), fast


@pytest.mark.parametrize("max_len", [9, 16, 17, 31, 32, 33, 48, 49, 62, 64])
def test_every_row_geometry(ctx, oracle, max_len):
    """Lengths 1 .. max_len (+ a few longer ones the JVM would die on, filtered), both strands, N calls, special flags,
    reads touching both ends of short contigs, some clipped / gapped reads for the deferred kernel."""
    rng = random.Random(1000 + max_len)
    contigs = random_genome(rng, n_contigs=3, length=300, n_frac=0.03)
    g = po.Genome(dict(contigs))
    recs = random_records(rng, contigs, 2500, kinds=("M",) * 8 + ("clip", "indel"), Lrange=(1, max_len), flags_special=0.05)
    for name, seq in contigs:                        # reads flush with the contig ends
        for L in (1, max_len // 2 + 1, max_len):
            for pos in (1, len(seq) - L + 1):
                b = bytes(seq[pos - 1:pos - 1 + L]).upper().replace(b"N", b"A")
                recs.append(Record(rng.choice([0, 16]), name, pos, f"{L}M", b, bytes(rng.randint(2, 40) for _ in range(L))))
    order = {n: i for i, (n, _) in enumerate(contigs)}
    recs.sort(key=lambda r: (order[r.rname], r.pos))
    ok, _ = survivors(recs, g, max_len)
    ref = PackedReference.from_contigs(contigs)
    batch = ReadBatch.from_records(ok, ref)
    assert batch.uniform_len == 0
    got, fast = profile_counting_fast(ctx, ref, batch, max_len)
    assert_profile_equal(got, oracle.profile(ref, batch, max_len), f"max_len {max_len}")
    single_m = sum(1 for r in ok if r.cigar == f"{len(r.seq)}M" and not (r.flag & 0x404))
    assert fast >= single_m * 0.8 and fast <= len(ok), (fast, single_m, len(ok))


def test_faults_keep_their_order(ctx, oracle):
    rng = random.Random(77)
    contigs = random_genome(rng, n_contigs=2, length=400)
    g = po.Genome(dict(contigs))
    recs = random_records(rng, contigs, 1500, kinds=("M", "M", "indel"), Lrange=(10, 60), flags_special=0.05)
    ok, bad = survivors(recs, g, 48)                 # reads longer than 48 are what dies here
    assert len(bad) > 5
    mixed = ok[:300] + [bad[0]] + ok[300:600] + bad[1:3] + ok[600:]
    ref = PackedReference.from_contigs(contigs)
    mb = ReadBatch.from_records(mixed, ref)
    with pytest.raises(oracle.OracleFault) as eo:
        oracle.profile(ref, mb, 48)
    ctx.upload_reference(ref)
    with pytest.raises(abi.ReferenceWouldThrow) as eg:
        ctx.profile(mb, 48)
    assert eg.value.fault == (eo.value.code, eo.value.ordinal)


def trimmed_reads(rng, contigs, n, lo, hi, indel_every=0):
    """Single-M reads of lengths lo..hi inside the contigs (what survives adapter trimming), sorted."""
    recs = []
    for k in range(n):
        ci = rng.randrange(len(contigs))
        name, seq = contigs[ci]
        L = rng.randint(lo, hi)
        pos = rng.randint(1, len(seq) - L - 2)
        b = bytearray(bytes(seq[pos - 1:pos - 1 + L]).upper())
        for j in range(L):
            if b[j] not in b"ACGT":
                b[j] = ord("A")
            x = rng.random()
            if x < 0.02:
                b[j] = rng.choice(b"ACGT")
            elif x < 0.023:
                b[j] = ord("N")
        cigar = f"{L}M"
        if indel_every and k % indel_every == 0 and L > 8:
            cigar = f"{L // 2}M1I{L - L // 2 - 1}M"
        recs.append(Record(16 if rng.random() < 0.5 else 0, name, pos, cigar, bytes(b), bytes(rng.randint(2, 41) for _ in range(L))))
    order = {nm: i for i, (nm, _) in enumerate(contigs)}
    recs.sort(key=lambda r: (order[r.rname], r.pos))
    return recs


@pytest.mark.parametrize("chunk", [None, 1024])
def test_trimmed_reads_in_chunks(ctx, oracle, monkeypatch, chunk):
    """40 000 trimmed reads (18..44 nt, an insertion in every 50th); with a small chunk the batch takes several repack +
    kernel passes and one deferred pass at the end."""
    if chunk:
        monkeypatch.setenv("PARASUITE_B200_RAGGED_CHUNK", str(chunk))
    rng = random.Random(5)
    contigs = random_genome(rng, n_contigs=2, length=200_000, n_frac=0.002)
    recs = trimmed_reads(rng, contigs, 40_000, 18, 44, indel_every=50)
    ref = PackedReference.from_contigs(contigs)
    batch = ReadBatch.from_records(recs, ref)
    got, fast = profile_counting_fast(ctx, ref, batch, 51)
    assert_profile_equal(got, oracle.profile(ref, batch, 51, threads=4), f"chunk {chunk}")
    want = sum(1 for r in recs if "I" not in r.cigar)
    assert want - 64 <= fast <= want       # the warp-tile that straddles the two contigs defers the second contig's reads


def test_several_batches_of_a_run_and_the_switch(ctx, oracle, monkeypatch):
    """Ragged and uniform batches in one run (the deferred-read counters take turns across both kernels), and the same
    ragged batch with the path switched off (warp-per-read kernel) gives the same vector."""
    from parasuite_b200 import synth
    rng = random.Random(9)
    ref = synth.synth_reference(3, [300_000, 200_000], n_run=500)
    uni = synth.synth_reads(ref, 30_000, 36, seed=4, special_ppm=20_000)
    from parasuite_b200.bamio import batch_to_records
    base = batch_to_records(synth.synth_reads(ref, 6_000, 40, seed=8, special_ppm=0), ref)
    rag = []
    for r in base:                                     # trim every read to its own length, keep the alignment start
        L = rng.randint(15, 40)
        rag.append(Record(r.flag, r.rname, r.pos, f"{L}M", r.seq[:L], r.qual[:L]))
    rb = ReadBatch.from_records(rag, ref)
    ctx.upload_reference(ref)
    ctx.profile_begin(51)
    acc = None
    for b in (rb, uni, rb, uni):
        ctx.profile_batch(b)
        acc = oracle.profile_acc(ref, b, 51, threads=4, acc=acc) if acc is not None else oracle.profile_acc(ref, b, 51, threads=4)
    got = ctx.profile_end()
    assert np.array_equal(got["wide"], acc)
    monkeypatch.setenv("PARASUITE_B200_NO_RAGGED_FAST", "1")
    ctx.profile_begin(51)
    ctx.profile_batch(rb)
    assert int(ctx.lib.ps_debug_word(ctx.h, 1)) == 0
    off = ctx.profile_end()
    ctx.profile_begin(51)
    monkeypatch.delenv("PARASUITE_B200_NO_RAGGED_FAST")
    ctx.profile_batch(rb)
    assert int(ctx.lib.ps_debug_word(ctx.h, 1)) > 5000
    on = ctx.profile_end()
    assert np.array_equal(on["wide"], off["wide"])


def test_trimmed_synthetic_at_scale(ctx, oracle):
    """600 000 trimmed reads (20..50 nt) with special flags and N calls, against the multi-threaded oracle."""
    from parasuite_b200 import synth
    ref = synth.synth_reference(21, [3_000_000, 1_000_000], n_run=2000)
    batch = synth.trim_uniform(synth.synth_reads(ref, 600_000, 50, seed=5, special_ppm=2000, n_ppm=5000), 20, seed=6)
    got, fast = profile_counting_fast(ctx, ref, batch, 51)
    assert_profile_equal(got, oracle.profile(ref, batch, 51, threads=8), "trimmed synthetic")
    assert fast > 590_000


@pytest.mark.parametrize("trim", [0, 20])
def test_quality_histogram_on_the_fast_path(ctx, oracle, trim):
    """-q (ErrorProfiling.java:402-407) with the fast kernels: their ok-map + the histogram kernel for the reads they took,
    the deferred kernel for the others; uniform 36-nt batch and the same batch trimmed to 20..36 nt."""
    from parasuite_b200 import synth
    ref = synth.synth_reference(33, [2_000_000, 500_000], n_run=1500)
    batch = synth.synth_reads(ref, 300_001, 36, seed=12, special_ppm=3000, n_ppm=4000)
    if trim:
        batch = synth.trim_uniform(batch, trim, seed=13)
    ctx.upload_reference(ref)
    ctx.profile_begin(51, infer_qualities=True)
    ctx.profile_batch(batch)
    fast = int(ctx.lib.ps_debug_word(ctx.h, 1))
    got = ctx.profile_end()
    exp = oracle.profile(ref, batch, 51, True, threads=8)
    assert_profile_equal(got, exp, f"-q trim {trim}")
    assert np.array_equal(got["quality_hist"], exp["quality_hist"])
    assert fast > 290_000 and int(got["quality_hist"].sum()) > 290_000 * 20
