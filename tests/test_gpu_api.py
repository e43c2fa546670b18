"""GPU: behaviour of the C ABI around the kernels -- call order, argument checks, error recovery, several contexts."""
import ctypes as C

import numpy as np
import pytest

from helpers import assert_profile_equal, kat_records
from kat_vectors import KAT_MAXLEN, KAT_REF, PROFILE_KATS
from parasuite_b200 import PackedReference, ReadBatch, Record, abi

pytestmark = pytest.mark.gpu


def small():
    ref = PackedReference.from_contigs([("chr1", KAT_REF.encode())])
    recs = []
    for kid in sorted(PROFILE_KATS):
        recs += kat_records(PROFILE_KATS[kid]["reads"])
    return ref, ReadBatch.from_records(recs, ref)


def test_call_order_and_arguments():
    from parasuite_b200.runtime import Context
    ref, batch = small()
    c = Context(0)
    try:
        with pytest.raises(abi.PsError) as e:                 # no reference yet
            c.profile_begin(KAT_MAXLEN)
        assert e.value.status == abi.PS_ERR_STATE
        with pytest.raises(abi.PsError) as e:
            c.pileup(batch)
        assert e.value.status == abi.PS_ERR_STATE
        c.upload_reference(ref)
        with pytest.raises(abi.PsError) as e:                 # batch before begin
            c.profile_batch(batch)
        assert e.value.status == abi.PS_ERR_STATE
        for bad in (0, 3001):
            with pytest.raises(abi.PsError) as e:
                c.profile_begin(bad)
            assert e.value.status == abi.PS_ERR_INVALID_ARG
        assert c.lib.ps_profile_begin(c.h, None) == abi.PS_ERR_INVALID_ARG
        assert c.lib.ps_create(None, 0) == abi.PS_ERR_INVALID_ARG
        h = C.c_void_p()
        assert c.lib.ps_create(C.byref(h), 9999) == abi.PS_ERR_NO_DEVICE and not h.value
        assert b"sorted" in c.lib.ps_strerror(abi.PS_ERR_UNSORTED)
        n0 = c.kernel_launches()
        c.profile(batch, KAT_MAXLEN)
        assert c.kernel_launches() > n0
    finally:
        c.close()


def test_recovers_after_a_fault(oracle):
    """A record the JVM would die on makes the run fail with its ordinal; the next run on the same context is clean."""
    from parasuite_b200.runtime import Context
    ref, batch = small()
    bad = Record(0, "chr1", len(KAT_REF) - 2, "8M", b"ACGTACGT", bytes([30] * 8))     # runs past the contig end
    recs = kat_records(PROFILE_KATS["A"]["reads"]) + [bad]
    bb = ReadBatch.from_records(recs, ref)
    c = Context(0)
    try:
        c.upload_reference(ref)
        with pytest.raises(abi.ReferenceWouldThrow) as e:
            c.profile(bb, KAT_MAXLEN)
        assert e.value.fault == (abi.PS_THROW_REF_RANGE, len(recs) - 1)
        with pytest.raises(abi.ReferenceWouldThrow) as e:
            c.pileup(bb)
        assert e.value.fault[1] == len(recs) - 1
        assert_profile_equal(c.profile(batch, KAT_MAXLEN), oracle.profile(ref, batch, KAT_MAXLEN), "after a fault")
        got = c.pileup(ReadBatch.from_records(kat_records(PROFILE_KATS["A"]["reads"]), ref))
        assert got["counters"]["num_reads_processed"] == 1
    finally:
        c.close()


def test_two_contexts_interleaved(oracle):
    from parasuite_b200 import synth
    from parasuite_b200.runtime import Context
    ref = synth.synth_reference(61, [1_000_000], n_run=500)
    b1 = synth.synth_reads(ref, 60_000, 36, seed=21)
    b2 = synth.synth_reads(ref, 50_000, 50, seed=22)
    c1, c2 = Context(0), Context(0)
    try:
        c1.upload_reference(ref)
        c2.upload_reference(ref)
        c1.profile_begin(51)
        c2.profile_begin(51)
        c1.profile_batch(b1)
        c2.profile_batch(b2)
        p2 = c2.pileup(b2)
        r1, r2 = c1.profile_end(), c2.profile_end()
        assert np.array_equal(r1["wide"], oracle.profile_acc(ref, b1, 51, threads=4))
        assert np.array_equal(r2["wide"], oracle.profile_acc(ref, b2, 51, threads=4))
        e2 = oracle.pileup(ref, b2)
        assert np.array_equal(p2["clusters"]["num_t2c"], e2["clusters"]["num_t2c"])
        assert np.array_equal(p2["sites"]["cov"], e2["sites"]["cov"])
    finally:
        c1.close()
        c2.close()


def test_handle_outlives_next_call(oracle):
    """Cluster and site records belong to the handle: a later pileup call on the same context does not disturb them."""
    from parasuite_b200 import synth
    from parasuite_b200.runtime import Context
    ref = synth.synth_reference(62, [1_000_000], n_run=500)
    b1 = synth.synth_reads(ref, 40_000, 36, seed=23)
    b2 = synth.synth_reads(ref, 30_000, 36, seed=24)
    c = Context(0)
    try:
        c.upload_reference(ref)
        h1 = c.pileup_run(b1)
        h2 = c.pileup_run(b2)
        g1, g2 = h1.fetch(boundary=False), h2.fetch(boundary=False)
        h1.close(); h2.close()
        for g, b in ((g1, b1), (g2, b2)):
            e = oracle.pileup(ref, b)
            assert np.array_equal(g["clusters"]["end"], e["clusters"]["end"])
            assert np.array_equal(g["sites"]["order_key"], e["sites"]["order_key"])
    finally:
        c.close()


def test_compact_host_form_equals_the_full_one(oracle):
    """ps_read_batch's compact host form (one flag byte per read instead of the meta word, no cigar stream when every
    read has the same single op): the upload expands it on the device; profile, pileup and the returned device view are
    the same as with the full arrays."""
    from parasuite_b200 import synth
    from parasuite_b200.runtime import Context, PinnedBatch
    ref = synth.synth_reference(91, [1_500_000, 700_000], n_run=800)
    batch = synth.synth_reads(ref, 200_003, 36, seed=14, n_ppm=3000)      # (special flags would kill the JVM in the pileup)
    full, comp = PinnedBatch(batch), PinnedBatch(batch, compact=True)
    nt = (batch.n_reads + 255) // 256
    assert comp.compact and comp.packed_qual and comp.compact_start
    assert comp.h2d_bytes == full.h2d_bytes - (7 + 9 + 2) * batch.n_reads + 4 * nt
    ctx = Context(0)
    try:
        ctx.upload_reference(ref)
        out = {}
        for name, pb in (("full", full), ("compact", comp)):
            view = ctx.upload(pb)
            with ctx.pileup_run(view) as h:
                pile = h.fetch(boundary=True)
            ctx.profile_begin(51)
            ctx.profile_batch_device(view)
            out[name] = (ctx.profile_end(), pile)
            ctx.profile_begin(51)                      # and through the calls that upload by themselves
            ctx.profile_batch(pb)
            assert np.array_equal(ctx.profile_end()["wide"], out[name][0]["wide"])
            with ctx.pileup_run(pb) as h:
                p2 = h.fetch(boundary=True)
            assert np.array_equal(p2["clusters"], pile["clusters"]) and np.array_equal(p2["sites"], pile["sites"])
        assert np.array_equal(out["full"][0]["wide"], out["compact"][0]["wide"])
        assert np.array_equal(out["full"][1]["clusters"], out["compact"][1]["clusters"])
        assert np.array_equal(out["full"][1]["sites"], out["compact"][1]["sites"])
        assert np.array_equal(out["compact"][0]["wide"], oracle.profile_acc(ref, batch, 51, threads=4))
        # reads without a defined start (unmapped, POS 0) travel with the tile's base as their start: the profile's
        # filters never look at it
        b2 = synth.synth_reads(ref, 60_000, 36, seed=17, special_ppm=20_000, n_ppm=3000)
        c2 = PinnedBatch(b2, compact=True)
        assert c2.compact_start
        ctx.profile_begin(51)
        ctx.profile_batch(c2)
        assert np.array_equal(ctx.profile_end()["wide"], oracle.profile_acc(ref, b2, 51, threads=4))
        # a ragged batch has no compact form
        rb = synth.trim_uniform(synth.synth_reads(ref, 10_000, 36, seed=15), 20)
        assert not PinnedBatch(rb, compact=True).compact
    finally:
        ctx.close()


@pytest.mark.parametrize("L", [50, 17])
def test_packed_qualities_for_lengths_that_are_no_multiple_of_four(oracle, L):
    from parasuite_b200 import synth
    from parasuite_b200.runtime import Context, PinnedBatch
    ref = synth.synth_reference(92, [900_000], n_run=800)
    batch = synth.synth_reads(ref, 50_001, L, seed=16, n_ppm=2000)
    comp = PinnedBatch(batch, compact=True)
    assert comp.packed_qual
    ctx = Context(0)
    try:
        ctx.upload_reference(ref)
        ctx.profile_begin(51)
        ctx.profile_batch(comp)
        assert np.array_equal(ctx.profile_end()["wide"], oracle.profile_acc(ref, batch, 51, threads=4))
    finally:
        ctx.close()
