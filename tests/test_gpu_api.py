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
