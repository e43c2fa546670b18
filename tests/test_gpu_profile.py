"""GPU parity (through the C ABI): error-profile kernel vs the CPU oracle.  Bit-exact (integer counting)."""
import random

import numpy as np
import pytest

import py_oracle as po
from helpers import assert_profile_equal, kat_records, random_genome, random_records, to_py
from kat_vectors import KAT_MAXLEN, KAT_REF, PROFILE_KATS
from parasuite_b200 import PackedReference, ReadBatch, Record, abi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from parasuite_b200.runtime import Context
    c = Context(0)
    yield c
    c.close()


def run_gpu(ctx, ref, batch, max_len, infer_q=False):
    ctx.upload_reference(ref)
    return ctx.profile(batch, max_len, infer_q)


@pytest.mark.parametrize("kid", sorted(PROFILE_KATS))
def test_kat(ctx, oracle, kid):
    ref = PackedReference.from_contigs([("chr1", KAT_REF.encode())])
    batch = ReadBatch.from_records(kat_records(PROFILE_KATS[kid]["reads"]), ref)
    got = run_gpu(ctx, ref, batch, KAT_MAXLEN)
    assert_profile_equal(got, oracle.profile(ref, batch, KAT_MAXLEN), kid)


def test_kat_all(ctx, oracle):
    ref = PackedReference.from_contigs([("chr1", KAT_REF.encode())])
    recs = []
    for kid in sorted(PROFILE_KATS):
        recs += kat_records(PROFILE_KATS[kid]["reads"])
    batch = ReadBatch.from_records(recs * 40, ref)          # > 1 tile, ragged last tile
    got = run_gpu(ctx, ref, batch, KAT_MAXLEN)
    assert_profile_equal(got, oracle.profile(ref, batch, KAT_MAXLEN), "all KATs x40")


def test_empty_batch(ctx):
    ref = PackedReference.from_contigs([("chr1", KAT_REF.encode())])
    batch = ReadBatch.from_records([], ref)
    got = run_gpu(ctx, ref, batch, KAT_MAXLEN)
    assert got["position_conversions"].sum() == 0 and got["counters"].sum() == 0


@pytest.mark.parametrize("seed,kinds", [(1, ("M",)), (2, ("M", "clip")), (3, ("indel",)), (4, ("splice",)),
                                        (5, ("wild", "indel", "clip", "M", "splice")), (6, ("wild",))])
def test_random_records(ctx, oracle, seed, kinds):
    rng = random.Random(seed)
    contigs = random_genome(rng, n_contigs=3, length=500)
    recs = random_records(rng, contigs, 1500, kinds=kinds, flags_special=0.1)
    ref = PackedReference.from_contigs(contigs)
    g = po.Genome(dict(contigs))
    ok, bad = [], []
    for r in recs:
        try:
            po.profile(to_py([r]), g, 48)
            ok.append(r)
        except po.ReferenceWouldThrow:
            bad.append(r)
    batch = ReadBatch.from_records(ok, ref)
    for infer_q in (False, True):
        if infer_q:   # -q also touches every i < mappingLength: keep reads on which that does not throw
            ok_q = []
            for r in ok:
                try:
                    po.profile(to_py([r]), g, 48, infer_qual=True)
                    ok_q.append(r)
                except po.ReferenceWouldThrow:
                    pass
            batch = ReadBatch.from_records(ok_q, ref)
        got = run_gpu(ctx, ref, batch, 48, infer_q)
        exp = oracle.profile(ref, batch, 48, infer_q)
        assert_profile_equal(got, exp, f"seed {seed} q={infer_q}")
        if infer_q:
            assert np.array_equal(got["quality_hist"], exp["quality_hist"])
    # faults: same first ordinal and same reason as the oracle
    if bad:
        mixed = ok[:300] + [bad[0]] + ok[300:600] + bad[1:3] + ok[600:]
        mb = ReadBatch.from_records(mixed, ref)
        with pytest.raises(oracle.OracleFault) as eo:
            oracle.profile(ref, mb, 48)
        ctx.upload_reference(ref)
        with pytest.raises(abi.ReferenceWouldThrow) as eg:
            ctx.profile(mb, 48)
        assert eg.value.fault == (eo.value.code, eo.value.ordinal)


@pytest.mark.parametrize("mode,L,max_len,n", [(0, 36, 51, 300_000), (0, 50, 51, 200_000), (1, 150, 176, 100_000),
                                              (0, 17, 20, 50_001), (0, 40, 51, 100_003), (0, 51, 51, 100_003),
                                              (0, 32, 40, 100_003), (0, 64, 64, 60_001), (0, 47, 51, 60_001)])
def test_synthetic(ctx, oracle, mode, L, max_len, n):
    from parasuite_b200 import synth
    ref = synth.synth_reference(11 + L, [3_000_000, 2_000_000], n_run=2000)
    batch = synth.synth_reads(ref, n, L, seed=100 + L, mode=mode, special_ppm=500)
    got = run_gpu(ctx, ref, batch, max_len)
    assert_profile_equal(got, oracle.profile(ref, batch, max_len, threads=8), f"synthetic mode {mode} L {L}")
    assert got["position_conversions"].sum() == got["counters"][7]


@pytest.mark.parametrize("period", [1, 40, 5000])
def test_warp_sums_flushed_in_the_middle_of_a_launch(ctx, oracle, monkeypatch, period):
    """The warp-per-read kernel keeps its per-lane sums in 32 bits and empties them into the block's 64-bit cells before
    2^23 columns per lane have gone in -- never reached at test sizes, so the period is shortened here (every read, every
    few reads, a few times per warp): same counts."""
    from parasuite_b200 import synth
    monkeypatch.setenv("PARASUITE_B200_GENERIC_FLUSH_READS", str(period))
    ref = synth.synth_reference(161, [3_000_000, 2_000_000], n_run=2000)
    batch = synth.synth_reads(ref, 100_000, 150, seed=250, mode=1, special_ppm=500)
    got = run_gpu(ctx, ref, batch, 176)
    assert_profile_equal(got, oracle.profile(ref, batch, 176, threads=8), f"flush period {period}")


def test_device_resident_and_multibatch(ctx, oracle):
    """Device-resident entry point, several batches per run (ordinals continue), host path equality."""
    import torch
    from parasuite_b200 import synth
    from parasuite_b200.runtime import DeviceBatch
    ref = synth.synth_reference(5, [2_000_000], n_run=1000)
    b1 = synth.synth_reads(ref, 120_000, 36, seed=1)
    b2 = synth.synth_reads(ref, 70_000, 36, seed=2, mode=1)
    ctx.upload_reference(ref)
    ctx.profile_begin(51)
    d1 = DeviceBatch(b1, "cuda:0")
    d2 = DeviceBatch(b2, "cuda:0")
    ctx.profile_batch_device(d1, torch.cuda.current_stream().cuda_stream)
    ctx.profile_batch_device(d2, torch.cuda.current_stream().cuda_stream)
    ctx.profile_batch(b1)
    acc_t = ctx.profile_acc_tensor()
    torch.cuda.synchronize()
    got = ctx.profile_end()
    acc = oracle.profile_acc(ref, b1, 51, threads=4)
    acc = oracle.profile_acc(ref, b2, 51, threads=4, acc=acc)
    acc = oracle.profile_acc(ref, b1, 51, threads=4, acc=acc)
    assert_profile_equal(got, oracle.split_acc(acc, 51), "multi-batch")
    assert np.array_equal(got["wide"], acc)
    assert acc_t.numel() == acc.size


def test_deferred_reads_in_every_batch_of_a_run(ctx, oracle):
    """Reads the fast kernel hands over (special flags, contig edges) in each of four batches of one run: the two
    deferred-read counters take turns, each batch's deferred kernel clearing the next batch's."""
    import torch
    from parasuite_b200 import synth
    from parasuite_b200.runtime import DeviceBatch
    ref = synth.synth_reference(6, [1_000_000, 400_000], n_run=1000)
    batches = [synth.synth_reads(ref, 40_000 + 7_001 * k, 36, seed=30 + k, special_ppm=30_000) for k in range(4)]
    ctx.upload_reference(ref)
    ctx.profile_begin(51)
    acc = None
    for k, b in enumerate(batches):
        if k % 2:
            ctx.profile_batch(b)
        else:
            ctx.profile_batch_device(DeviceBatch(b, "cuda:0"), torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
        acc = oracle.profile_acc(ref, b, 51, threads=4, acc=acc) if acc is not None else oracle.profile_acc(ref, b, 51, threads=4)
    got = ctx.profile_end()
    assert np.array_equal(got["wide"], acc)
    assert int(got["counters"][1]) + int(got["counters"][2]) > 1000      # unmapped + duplicates went through the deferred list


def test_int32_wrap(ctx, oracle):
    """Java int wrap-around: quality sums exceed 2^31 within a few million reads."""
    from parasuite_b200 import synth
    ref = synth.synth_reference(3, [20_000_000], n_run=0)
    batch = synth.synth_reads(ref, 9_000_000, 36, seed=3)
    got = run_gpu(ctx, ref, batch, 51)
    exp = oracle.profile(ref, batch, 51, threads=8)
    assert_profile_equal(got, exp, "wrap")
    assert (got["wide"][16 * 51:16 * 51 + 16] > 2 ** 31).any()
    assert (got["quality_per_mismatch"] < 0).any()


def test_full_size_config2(ctx, oracle):
    """BASELINE config 2 at full size (10M x 36 nt, 100 Mb reference) against the multi-threaded oracle,
    plus the size-independent invariants: sum(cells) == totalBasesChecked, counts <= reads per position."""
    from parasuite_b200 import synth
    ref = synth.synth_reference(0x5EED0001, [100_000_000])
    batch = synth.synth_reads(ref, 10_000_000, 36, seed=0x5EED0002)
    got = run_gpu(ctx, ref, batch, 51)
    c = got["counters"]
    assert got["position_conversions"].astype(np.int64).sum() == int(c[7])
    assert (got["position_conversions"].reshape(51, 16).sum(axis=1) <= batch.n_reads).all()
    assert got["position_conversions"][36:].sum() == 0
    exp = oracle.profile(ref, batch, 51, threads=16)
    assert_profile_equal(got, exp, "config 2 full size")


@pytest.mark.parametrize("infer_q", [False, True])
def test_long_reads_dense_cigar(ctx, oracle, infer_q):
    """Maximum sizes: reads of up to 2500 bases with soft clips, indels, splices and `=`/`X` blocks at the largest
    maxReadLength the library accepts (3000): many 32-column passes per block in the general kernel; with -q also the
    per-position quality histogram."""
    rng = random.Random(77)
    contigs = random_genome(rng, n_contigs=2, length=30000, n_frac=0.002, lower_frac=0.05)
    recs = random_records(rng, contigs, 600, kinds=("M", "clip", "indel", "wild"), Lrange=(800, 2500), flags_special=0.02)
    ref = PackedReference.from_contigs(contigs)
    ok = []
    for r in recs:      # keep what the JVM survives (C++ oracle, one record at a time)
        try:
            oracle.profile(ref, ReadBatch.from_records([r], ref), 3000, infer_q)
            ok.append(r)
        except oracle.OracleFault:
            pass
    assert len(ok) > 200
    batch = ReadBatch.from_records(ok, ref)
    ctx.upload_reference(ref)
    got = ctx.profile(batch, 3000, infer_q)
    exp = oracle.profile(ref, batch, 3000, infer_q)
    assert_profile_equal(got, exp, "long reads")
    if infer_q:
        assert np.array_equal(got["quality_hist"], exp["quality_hist"])


def test_more_than_255_cigar_ops_is_flagged(ctx, oracle):
    """A record with > 255 cigar ops cannot be held by the SoA (PS_RF_CIGAR_OVERFLOW; htsjdk takes any number,
    ErrorProfiling.java:206-207): both tools refuse the batch by name (PS_ERR_UNSUPPORTED) instead of inventing a result."""
    contigs = [("chr1", b"ACGT" * 400)]
    cig = "".join("1M1I" for _ in range(130)) + "1M"          # 261 ops, 261 read bases... 131 M + 130 I
    seq = b"A" * 261
    rec = Record(0, "chr1", 10, cig, seq, bytes([30] * 261))
    ref = PackedReference.from_contigs(contigs)
    batch = ReadBatch.from_records([rec], ref)
    assert (int(batch.meta[0]) >> 24) & abi.PS_RF_CIGAR_OVERFLOW
    ctx.upload_reference(ref)
    with pytest.raises(abi.PsError) as e:
        ctx.profile(batch, 600)
    assert e.value.status == abi.PS_ERR_UNSUPPORTED and not isinstance(e.value, abi.ReferenceWouldThrow)
    with pytest.raises(abi.PsError) as e:
        ctx.pileup(batch)
    assert e.value.status == abi.PS_ERR_UNSUPPORTED
    # a good read in front of it: the fault names the offending record, earlier records are not affected
    good = Record(0, "chr1", 5, "20M", b"ACGT" * 5, bytes([30] * 20))
    with pytest.raises(abi.PsError) as e:
        ctx.profile(ReadBatch.from_records([good, rec], ref), 600)
    assert e.value.status == abi.PS_ERR_UNSUPPORTED and "record 1 " in str(e.value)


@pytest.mark.parametrize("L,ragged", [(36, False), (50, False), (40, True)])
def test_tiles_with_more_events_than_the_queue_holds(ctx, oracle, L, ragged):
    """Reads unrelated to the reference (three mismatches in four bases): a warp-tile of 64 such reads holds far more
    mismatch events than the per-warp event queue of the fast kernel (FAST_QCAP), which then drains lane by lane; mixed
    with clean reads so that both ways run in one launch, with and without T>C mask words."""
    rng = random.Random(4242 + L)
    contigs = random_genome(rng, n_contigs=2, length=30_000, n_frac=0.002)
    recs = []
    for k in range(6000):
        name, seq = contigs[k % 2]
        Lr = rng.randint(12, L) if ragged else L
        pos = rng.randint(1, len(seq) - Lr - 1)
        noisy = (k // 300) % 2 == 0                     # runs of 300 noisy reads, then 300 clean ones
        if noisy:
            b = bytes(rng.choice(b"ACGT") for _ in range(Lr))
        else:
            b = bytes(c if c in b"ACGT" else ord("A") for c in bytes(seq[pos - 1:pos - 1 + Lr]).upper())
        recs.append(Record(16 if rng.random() < 0.5 else 0, name, pos, f"{Lr}M", b, bytes(rng.randint(2, 41) for _ in range(Lr))))
    order = {n: i for i, (n, _) in enumerate(contigs)}
    recs.sort(key=lambda r: (order[r.rname], r.pos))
    ref = PackedReference.from_contigs(contigs)
    batch = ReadBatch.from_records(recs, ref)
    exp = oracle.profile(ref, batch, 51)
    assert_profile_equal(run_gpu(ctx, ref, batch, 51), exp, f"noisy tiles L {L}")
    if not ragged:          # the mask-emitting variant of the same kernel, and the pileup fed by its masks
        import torch
        from parasuite_b200.runtime import DeviceBatch
        d = DeviceBatch(batch, "cuda:0")
        st = torch.cuda.current_stream().cuda_stream
        ctx.profile_begin(51, emit_t2c_masks=True)
        ctx.profile_batch_device(d, st)
        masks = ctx.profile_masks()
        with ctx.pileup_run(d, stream=st, masks=masks) as h:
            got_m = h.fetch(boundary=False)
        assert_profile_equal(ctx.profile_end(), exp, "noisy tiles, masks")
        with ctx.pileup_run(d, stream=st) as h:
            got_d = h.fetch(boundary=False)
        assert np.array_equal(got_m["clusters"], got_d["clusters"]) and np.array_equal(got_m["sites"], got_d["sites"])
