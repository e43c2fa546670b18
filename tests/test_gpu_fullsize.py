"""GPU, BASELINE.json's full single-GPU size (config 2: 10M x 36-nt reads vs a 100 Mb reference): bit-exact parity of
the error profile against the multi-threaded oracle, plus size-independent properties of both tools --
  * linearity: profile(A) + profile(B) == profile(A ++ B) on the 64-bit accumulators (what makes read-batch sharding
    + all-reduce exact);
  * checksum of checksums: sum of all positionConversions cells == totalBasesChecked == sum of qualityPerMismatchCounts;
  * pileup conservation: reads in clusters (+ open cluster) == records kept; T>C events in sites == events in clusters;
    coverage at a site >= its T>C count; sites sorted by (cluster, position);
  * region sharding: two shards + halo merge == the whole stream.
"""
import numpy as np
import pytest

from parasuite_b200 import abi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big():
    from parasuite_b200 import synth
    from parasuite_b200.runtime import Context, DeviceBatch
    ref = synth.synth_reference(0x5EED0001, [100_000_000])
    batch = synth.synth_reads(ref, 10_000_000, 36, seed=0x5EED0002)
    ctx = Context(0)
    ctx.upload_reference(ref)
    yield ctx, ref, batch, DeviceBatch(batch, "cuda:0")
    ctx.close()


def test_profile_parity_and_checksums(big, oracle):
    ctx, ref, batch, dbatch = big
    ctx.profile_begin(51)
    ctx.profile_batch_device(dbatch)
    got = ctx.profile_end()
    exp = oracle.profile_acc(ref, batch, 51, threads=16)
    assert np.array_equal(got["wide"], exp)
    w = got["wide"]
    conv, qcnt, ctr = w[:16 * 51], w[16 * 51 + 16:16 * 51 + 32], w[16 * 51 + 32 + 2 * 51:]
    assert conv.sum() == ctr[7] == qcnt.sum()                       # every counted base lands in exactly one cell
    assert ctr[0] == batch.n_reads                                   # config 2 holds no filtered records


def test_profile_linearity(big):
    from parasuite_b200.sharding import slice_batch
    ctx, ref, batch, dbatch = big
    cut = 4_999_936
    parts = []
    for lo, hi in ((0, cut), (cut, batch.n_reads)):
        ctx.profile_begin(51)
        ctx.profile_batch(slice_batch(batch, lo, hi))
        parts.append(ctx.profile_end()["wide"])
    ctx.profile_begin(51)
    ctx.profile_batch_device(dbatch)
    whole = ctx.profile_end()["wide"]
    assert np.array_equal(parts[0] + parts[1], whole)
    # two batches inside one begin/end accumulate the same way
    ctx.profile_begin(51)
    ctx.profile_batch(slice_batch(batch, 0, cut))
    ctx.profile_batch(slice_batch(batch, cut, batch.n_reads))
    assert np.array_equal(ctx.profile_end()["wide"], whole)


def test_pileup_conservation_and_sharding(big, oracle):
    from parasuite_b200.sharding import merge_pileup_shards, slice_batch
    ctx, ref, batch, dbatch = big
    res = ctx.pileup(dbatch)
    cl, si = res["clusters"], res["sites"]
    kept = int(((batch.meta >> 24) & abi.PS_RF_UNMAPPED == 0).sum())
    oc = res["open_cluster"]
    assert int(cl["num_reads"].sum()) + int(oc["num_reads"]) == kept == res["counters"]["num_reads_processed"]
    assert int(si["t2c"].sum()) == int(cl["num_t2c"].sum())
    assert (si["cov"] >= si["t2c"]).all() and (si["t2c"] >= 1).all()
    assert (np.diff(cl["site_begin"].astype(np.int64)) >= 0).all() and int(cl["site_end"][-1]) == len(si)
    owner = np.repeat(np.arange(len(cl)), (cl["site_end"] - cl["site_begin"]).astype(np.int64))
    same = owner[1:] == owner[:-1]
    assert (si["pos"][1:][same] > si["pos"][:-1][same]).all()        # position order inside a cluster
    assert (cl["running_id"] == np.arange(2, 2 + len(cl))).all()    # cl_2, cl_3, ... (PileupClusters.java:355)
    # oracle on a 1M-read prefix (single-threaded restatement)
    head = slice_batch(batch, 0, 1_000_000)
    exp = oracle.pileup(ref, head)
    got = ctx.pileup(head)
    for f in ("start", "end", "num_reads", "num_t2c", "mask51", "site_begin", "site_end", "combined_strand"):
        assert np.array_equal(got["clusters"][f], exp["clusters"][f]), f
    for f in ("pos", "t2c", "cov", "order_key"):
        assert np.array_equal(got["sites"][f], exp["sites"][f]), f
    # region sharding at an arbitrary cut
    cut = 6_123_457
    shards, carry = [], None
    for lo, hi in ((0, cut), (cut, batch.n_reads)):
        r = ctx.pileup(slice_batch(batch, lo, hi), carry=carry)
        shards.append(r)
        carry = merge_pileup_shards.carry_after(shards, carry)
    merged = merge_pileup_shards(shards, [0, cut])
    for f in ("first_read", "running_id", "start", "end", "num_reads", "num_t2c", "mask51", "minus_after_first"):
        assert np.array_equal(merged["clusters"][f], cl[f]), f
    for f in ("pos", "t2c", "cov", "order_key"):
        assert np.array_equal(merged["sites"][f], si[f]), f
