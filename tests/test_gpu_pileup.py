"""GPU parity (through the C ABI): T>C pileup kernels vs the CPU oracle.  Bit-exact."""
import random

import numpy as np
import pytest

import py_oracle as po
from helpers import kat_records, random_genome, random_records, to_py
from kat_vectors import PILEUP_READS, PILEUP_REF
from parasuite_b200 import PackedReference, ReadBatch, abi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from parasuite_b200.runtime import Context
    c = Context(0)
    yield c
    c.close()


CL_FIELDS = ("first_read", "running_id", "contig", "start", "end", "num_reads", "num_t2c", "minus_after_first",
             "first_reverse", "combined_strand", "mask51", "site_begin", "site_end")
SITE_FIELDS = ("pos", "t2c", "cov", "order_key")


def assert_pileup_equal(got, exp, what=""):
    assert got["counters"] == exp["counters"], (what, got["counters"], exp["counters"])
    for f in CL_FIELDS:
        if not np.array_equal(got["clusters"][f], exp["clusters"][f]):
            k = int(np.argwhere(got["clusters"][f] != exp["clusters"][f])[0][0])
            raise AssertionError(f"{what}: cluster field {f} differs first at {k}: "
                                 f"{got['clusters'][k]} vs {exp['clusters'][k]}")
    for f in SITE_FIELDS:
        if not np.array_equal(got["sites"][f], exp["sites"][f]):
            k = int(np.argwhere(got["sites"][f] != exp["sites"][f])[0][0])
            raise AssertionError(f"{what}: site field {f} differs first at {k}: {got['sites'][k]} vs {exp['sites'][k]}")
    if exp["open_cluster"] is None:
        assert got["open_cluster"] is None
    else:
        for f in CL_FIELDS:
            assert got["open_cluster"][f] == exp["open_cluster"][f], (what, "open", f)
        for f in SITE_FIELDS:
            assert np.array_equal(got["open_sites"][f], exp["open_sites"][f]), (what, "open sites", f)


def test_kat(ctx, oracle):
    ref = PackedReference.from_contigs([("chr1", PILEUP_REF.encode())])
    batch = ReadBatch.from_records(kat_records(PILEUP_READS), ref)
    ctx.upload_reference(ref)
    got = ctx.pileup(batch)
    assert_pileup_equal(got, oracle.pileup(ref, batch), "KAT")
    assert got["counters"]["double_stranded"] == 2 and len(got["clusters"]) == 3
    assert int(got["clusters"][0]["running_id"]) == 2          # first cluster is cl_2 (PileupClusters.java:355)


def test_empty_and_single(ctx, oracle):
    ref = PackedReference.from_contigs([("chr1", PILEUP_REF.encode())])
    ctx.upload_reference(ref)
    got = ctx.pileup(ReadBatch.from_records([], ref))
    assert len(got["clusters"]) == 0 and got["open_cluster"] is None
    one = ReadBatch.from_records(kat_records(PILEUP_READS[:1]), ref)
    assert_pileup_equal(ctx.pileup(one), oracle.pileup(ref, one), "single read")


@pytest.mark.parametrize("seed,kinds", [(11, ("M",)), (12, ("M", "clip", "indel")), (13, ("wild", "splice", "M")),
                                        (14, ("indel", "clip"))])
def test_random_records(ctx, oracle, seed, kinds):
    rng = random.Random(seed)
    contigs = random_genome(rng, n_contigs=3, length=8000, n_frac=0.003, lower_frac=0.05)
    recs = random_records(rng, contigs, 2500, kinds=kinds, Lrange=(15, 30), flags_special=0.05)
    recs = [r for r in recs if r.pos > 0]
    ref = PackedReference.from_contigs(contigs)
    g = po.Genome(dict(contigs))
    ok, bad = [], []
    for r in recs:
        try:
            po.pileup(to_py([r]), g, po.SnpDb([]), 1)
            ok.append(r)
        except po.ReferenceWouldThrow:
            bad.append(r)
    batch = ReadBatch.from_records(ok, ref)
    ctx.upload_reference(ref)
    got = ctx.pileup(batch)
    exp = oracle.pileup(ref, batch)
    assert_pileup_equal(got, exp, f"seed {seed}")
    assert len(got["clusters"]) > 50
    if bad:
        mixed = ok[:500] + [bad[0]] + ok[500:]
        mb = ReadBatch.from_records(mixed, ref)
        with pytest.raises(oracle.OracleFault) as eo:
            oracle.pileup(ref, mb)
        with pytest.raises(abi.ReferenceWouldThrow) as eg:
            ctx.pileup(mb)
        assert eg.value.fault == (eo.value.code, eo.value.ordinal)


@pytest.mark.parametrize("L,n,ppm", [(36, 400_000, 0), (50, 300_000, 2000), (21, 100_003, 0)])
def test_synthetic(ctx, oracle, L, n, ppm):
    from parasuite_b200 import synth
    from parasuite_b200.runtime import DeviceBatch
    ref = synth.synth_reference(21 + L, [4_000_000, 3_000_000], n_run=3000)
    batch = synth.synth_reads(ref, n, L, seed=200 + L, special_ppm=ppm)
    if ppm:   # POS==0 on a mapped record kills the JVM in the pileup loop: drop those for the parity run
        keep = ((batch.meta >> 24) & abi.PS_RF_POS_ZERO) == 0
        batch.meta[~keep] |= np.uint32(abi.PS_RF_UNMAPPED << 24)
    ctx.upload_reference(ref)
    exp = oracle.pileup(ref, batch)
    assert_pileup_equal(ctx.pileup(batch), exp, f"synthetic L {L} host")
    assert_pileup_equal(ctx.pileup(DeviceBatch(batch, "cuda:0")), exp, f"synthetic L {L} device")


def test_dense_overlapping_clusters(ctx, oracle):
    """Reads packed so densely that clusters chain and overlap by < 5 positions; deep pileups."""
    from parasuite_b200 import synth
    ref = synth.synth_reference(77, [200_000], n_run=0)
    batch = synth.synth_reads(ref, 150_000, 36, seed=5)      # ~27 reads per 36 bp: long chained clusters
    ctx.upload_reference(ref)
    assert_pileup_equal(ctx.pileup(batch), oracle.pileup(ref, batch), "dense")


def test_sparse_reads_capacity_retry(ctx, oracle):
    """Nearly every read is its own cluster: more cluster slots than the first-guess capacity (n/8) -> the scan
    kernel is re-run once with exact capacities; results must not depend on that."""
    from parasuite_b200 import synth
    from parasuite_b200.runtime import Context
    from parasuite_b200.sharding import take_uniform
    ref = synth.synth_reference(78, [30_000_000], n_run=0)
    dense = synth.synth_reads(ref, 640_000, 36, seed=9, n_ppm=0)
    batch = take_uniform(dense, np.arange(0, dense.n_reads, 16))      # ~1 read per cluster, 40 000 reads
    fresh = Context(0)                     # capacities are learnt per context: use one that has seen nothing
    try:
        fresh.upload_reference(ref)
        exp = oracle.pileup(ref, batch)
        got = fresh.pileup(batch)
        assert_pileup_equal(got, exp, "sparse")
        assert len(got["clusters"]) > 40_000 // 8
        assert_pileup_equal(fresh.pileup(batch), exp, "sparse, second call")
    finally:
        fresh.close()


def test_shared_upload(ctx, oracle):
    """ps_batch_upload: one H2D copy, both tools on the device view."""
    from helpers import assert_profile_equal
    from parasuite_b200 import synth
    ref = synth.synth_reference(79, [2_000_000], n_run=1000)
    batch = synth.synth_reads(ref, 120_000, 36, seed=10)
    ctx.upload_reference(ref)
    view = ctx.upload(batch)
    ctx.profile_begin(51)
    ctx.profile_batch_device(view)
    assert_profile_equal(ctx.profile_end(), oracle.profile(ref, batch, 51, threads=4), "shared upload profile")
    assert_pileup_equal(ctx.pileup(view), oracle.pileup(ref, batch), "shared upload pileup")


def test_region_sharding_halo_merge(ctx, oracle):
    """Two shards cut at an arbitrary read index; carry-in + head partial merge equals the whole-stream result."""
    from parasuite_b200 import synth
    from parasuite_b200.sharding import merge_pileup_shards, slice_batch
    ref = synth.synth_reference(31, [3_000_000, 1_000_000], n_run=2000)
    batch = synth.synth_reads(ref, 200_000, 36, seed=6)
    ctx.upload_reference(ref)
    whole = oracle.pileup(ref, batch)
    for cut in (100_000, 100_007, 65_536, 199_999, 1):
        parts = [slice_batch(batch, 0, cut), slice_batch(batch, cut, batch.n_reads)]
        shard_results = []
        carry = None
        for p in parts:
            res = ctx.pileup(p, carry=carry)
            shard_results.append(res)
            carry = merge_pileup_shards.carry_after(shard_results, carry)
        merged = merge_pileup_shards(shard_results, [0, cut])
        assert_pileup_equal(merged, whole, f"cut {cut}")


def test_deep_pileup_one_cluster(ctx, oracle):
    """A cluster far larger than the shared-memory chunk (60 000 reads on 3 kb, one chain): the streaming sweep of the
    warp routine (sliding ring of positions) against the oracle."""
    from parasuite_b200 import synth
    ref = synth.synth_reference(91, [3_000], n_run=0)
    batch = synth.synth_reads(ref, 60_000, 36, seed=13)
    assert np.all(np.diff(batch.ref_start.astype(np.int64)) >= 0)
    ctx.upload_reference(ref)
    got, exp = ctx.pileup(batch), oracle.pileup(ref, batch)
    assert_pileup_equal(got, exp, "deep pileup")
    assert exp["open_cluster"] is not None and int(exp["open_cluster"]["num_reads"]) > 10_000


def test_records_out_of_order_inside_a_contig(ctx, oracle):
    """The tools only check the header's sort order and take records in FILE order.  Records whose starts go backwards
    inside a contig must give the file-order result too (the sweep gives up, the window loop takes over)."""
    from parasuite_b200 import synth
    from parasuite_b200.sharding import take_uniform
    ref = synth.synth_reference(92, [40_000], n_run=0)
    batch = synth.synth_reads(ref, 30_000, 36, seed=14, n_ppm=0)
    rng = np.random.default_rng(5)
    idx = np.arange(batch.n_reads)
    for k in range(0, batch.n_reads - 8, 8):                # shuffle inside windows of 8 reads
        rng.shuffle(idx[k:k + 8])
    shuffled = take_uniform(batch, idx)
    assert np.any(np.diff(shuffled.ref_start.astype(np.int64)) < 0)
    ctx.upload_reference(ref)
    assert_pileup_equal(ctx.pileup(shuffled), oracle.pileup(ref, shuffled), "out of order inside a contig")


def test_device_resident_carry_keys(ctx, oracle):
    """Region sharding without a host round trip: the shards' maximum (contig, end) keys stay on the device (as after an
    all-gather) and the flag kernel takes the prefix-max itself; same result as the host carry and as the whole stream."""
    import torch
    from parasuite_b200 import synth
    from parasuite_b200.runtime import DeviceBatch
    from parasuite_b200.sharding import merge_pileup_shards, slice_batch
    ref = synth.synth_reference(33, [2_000_000, 1_500_000], n_run=1000)
    batch = synth.synth_reads(ref, 150_000, 36, seed=16)
    ctx.upload_reference(ref)
    whole = oracle.pileup(ref, batch)
    cuts = [0, 40_003, 90_000, 90_001, batch.n_reads]
    shards = [DeviceBatch(slice_batch(batch, lo, hi), "cuda:0") for lo, hi in zip(cuts[:-1], cuts[1:])]
    keys = torch.cat([ctx.pileup_max_key_tensor(s).clone() for s in shards])       # what the all-gather leaves on every rank
    torch.cuda.synchronize()
    host_keys = [ctx.pileup_max_key(s) for s in shards]
    for k, hk in zip(keys.tolist(), host_keys):
        assert (None if k == 0 else ((k >> 32) - 1, k & 0xFFFFFFFF)) == hk
    results = []
    for r, s in enumerate(shards):
        with ctx.pileup_run(s, carry_keys=(keys.data_ptr(), r)) as h:
            results.append(h.fetch())
    assert_pileup_equal(merge_pileup_shards(results, cuts[:-1]), whole, "device carry keys")


def test_record_reaching_over_the_halo_falls_back_to_exact_flags(oracle):
    """The boundary-flag pass of the vector path assumes that the running maximum of (contig, end) in front of a tile of
    2048 reads is the maximum over the 128 reads in front of it, and checks that afterwards.  One record with a 6 kb
    reference span in front of some 600 successors breaks the assumption: the call must notice, repeat with the exact
    look-back pass and give the file-order result; the context then stays in exact mode."""
    from parasuite_b200 import synth
    from parasuite_b200.runtime import Context
    ref = synth.synth_reference(93, [2_000_000], n_run=0)
    batch = synth.synth_reads(ref, 200_000, 36, seed=15)
    k = 2048 * 3 - 200                                  # its reach crosses a tile boundary by more than the halo
    span = 6_000
    assert int(batch.ref_start[k]) + span < 2_000_000
    reached = int(np.searchsorted(batch.ref_start, batch.ref_start[k] + span)) - k
    assert reached > 400
    batch.cigar[k] = (span << 4) | 3                    # a lone N op: no aligned block, alignment end 6 kb downstream
    c = Context(0)
    try:
        c.upload_reference(ref)
        assert c.lib.ps_pileup_flag_mode(c.h) == 0
        exp = oracle.pileup(ref, batch)
        got = c.pileup(batch)
        assert c.lib.ps_pileup_flag_mode(c.h) == 1
        assert_pileup_equal(got, exp, "record reaching over the halo")
        assert int(exp["clusters"]["num_reads"].max()) >= reached
        assert_pileup_equal(c.pileup(batch), exp, "exact mode, second call")
    finally:
        c.close()


def test_site_compaction_by_look_back(oracle, monkeypatch):
    """Past 2 M cluster slots per batch the site runs are ordered with a decoupled look-back over the compaction tiles
    instead of per-tile sums; PARASUITE_B200_COMPACT_LOOKBACK=1 (read by ps_create) forces that path."""
    from parasuite_b200 import synth
    from parasuite_b200.runtime import Context
    monkeypatch.setenv("PARASUITE_B200_COMPACT_LOOKBACK", "1")
    ref = synth.synth_reference(94, [3_000_000], n_run=500)
    batch = synth.synth_reads(ref, 250_000, 36, seed=17)
    c = Context(0)
    try:
        c.upload_reference(ref)
        assert_pileup_equal(c.pileup(batch), oracle.pileup(ref, batch), "compaction by look-back")
    finally:
        c.close()


def test_submit_and_wait(ctx, oracle):
    """ps_pileup_submit_device / ps_pileup_wait: the call in two halves gives the result of the synchronous call; until
    the wait the handle answers PS_ERR_STATE, the context refuses a second pileup call, and closing a submitted handle
    without waiting is safe."""
    import ctypes as C
    from parasuite_b200 import synth
    from parasuite_b200.runtime import DeviceBatch
    ref = synth.synth_reference(95, [1_500_000], n_run=300)
    batch = synth.synth_reads(ref, 120_000, 36, seed=18)
    ctx.upload_reference(ref)
    exp = oracle.pileup(ref, batch)
    d = DeviceBatch(batch, "cuda:0")
    h = ctx.pileup_run(d, defer=True)
    ctr = abi.ps_pileup_counters()
    assert ctx.lib.ps_pileup_counters_get(h.h, C.byref(ctr)) == abi.PS_ERR_STATE
    with pytest.raises(abi.PsError):
        ctx.pileup_run(d)                                  # one submitted call per context
    prof = ctx.profile(batch, 51)                          # other work of the context goes on in between
    got = h.wait().fetch()
    h.close()
    assert_pileup_equal(got, exp, "submit + wait")
    assert np.array_equal(prof["wide"], oracle.profile_acc(ref, batch, 51))
    h2 = ctx.pileup_run(d, defer=True)
    h2.close()                                             # waits, then frees
    assert_pileup_equal(ctx.pileup(d), exp, "after closing a submitted handle")


@pytest.mark.parametrize("n", [1, 127, 128, 129, 2047, 2048, 2049, 4096, 6145])
def test_flag_tiles_and_halo_edges(ctx, oracle, n):
    """Batch sizes around the tile (2048 reads) and halo (128 reads) of the speculative boundary-flag pass, two contigs so
    that a contig change falls next to a tile edge for some of them."""
    from parasuite_b200 import synth
    from parasuite_b200.sharding import slice_batch
    ref = synth.synth_reference(96, [30_000, 30_000], n_run=0)
    whole = synth.synth_reads(ref, 8_192, 36, seed=19, n_ppm=0)
    cut = int(np.searchsorted(whole.ref_start, 30_000))          # first read of the second contig
    lo = max(0, min(cut - n // 2, whole.n_reads - n))            # the contig change lies inside the slice where it can
    batch = slice_batch(whole, lo, lo + n)
    ctx.upload_reference(ref)
    assert_pileup_equal(ctx.pileup(batch), oracle.pileup(ref, batch), f"n={n}")
    assert ctx.lib.ps_pileup_flag_mode(ctx.h) == 0


def test_contig_order_backwards_is_refused(ctx, oracle):
    """Records of an earlier contig behind records of a later one: PS_ERR_UNSORTED from the speculative pass too."""
    from parasuite_b200 import synth
    ref = synth.synth_reference(97, [40_000, 40_000], n_run=0)
    whole = synth.synth_reads(ref, 9_000, 36, seed=20, n_ppm=0)
    cut = int(np.searchsorted(whole.ref_start, 40_000))
    assert 0 < cut < whole.n_reads
    idx = np.concatenate([np.arange(cut, whole.n_reads), np.arange(0, cut)])     # second contig first
    ctx.upload_reference(ref)
    with pytest.raises(abi.PsError) as e:
        ctx.pileup(_reorder(whole, idx))
    assert e.value.status == abi.PS_ERR_UNSORTED


def _reorder(batch, idx):
    """Reads of a uniform batch in the order `idx` (any order)."""
    import copy
    assert batch.uniform_len and batch.uniform_ncigar == 1 and not batch.exc_count
    b = copy.copy(batch)
    n, L = batch.n_reads, batch.uniform_len
    bpr = (L + 3) // 4
    b.meta = batch.meta[idx].copy(); b.ref_start = batch.ref_start[idx].copy(); b.cigar = batch.cigar[idx].copy()
    bases, qual = batch.bases2.copy(), batch.qual.copy()
    bases[:n * bpr] = batch.bases2[:n * bpr].reshape(n, bpr)[idx].reshape(-1)
    qual[:n * L] = batch.qual[:n * L].reshape(n, L)[idx].reshape(-1)
    b.bases2, b.qual = bases, qual
    return b


def test_flag_prefix_by_scan_kernel(oracle, monkeypatch):
    """Past 8192 tiles (16.7 M reads) the tile table of the speculative flag pass is prefixed and checked by the
    one-block scan kernel instead of inside the expansion kernel; PARASUITE_B200_FLAG_SCAN_KERNEL=1 (read by ps_create)
    forces that path: same results, and a record reaching over the halo is still noticed."""
    from parasuite_b200 import synth
    from parasuite_b200.runtime import Context
    monkeypatch.setenv("PARASUITE_B200_FLAG_SCAN_KERNEL", "1")
    ref = synth.synth_reference(98, [2_000_000, 500_000], n_run=200)
    batch = synth.synth_reads(ref, 150_000, 36, seed=21)
    c = Context(0)
    try:
        c.upload_reference(ref)
        assert_pileup_equal(c.pileup(batch), oracle.pileup(ref, batch), "scan kernel")
        assert c.lib.ps_pileup_flag_mode(c.h) == 0
        k, span = 2048 * 5 - 300, 7_000
        batch.cigar[k] = (span << 4) | 3
        assert_pileup_equal(c.pileup(batch), oracle.pileup(ref, batch), "scan kernel, record over the halo")
        assert c.lib.ps_pileup_flag_mode(c.h) == 1
    finally:
        c.close()
