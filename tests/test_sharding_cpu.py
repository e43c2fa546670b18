"""CPU: host-side multi-GPU logic -- read-range sharding, profile all-reduce (gloo, world_size 2) and the pileup
halo merge, checked with the oracle standing in for the per-rank kernels."""
import os
import socket

import numpy as np
import pytest

from parasuite_b200 import abi
from parasuite_b200.sharding import merge_pileup_shards, shard_ranges, slice_batch

CL_FIELDS = ("first_read", "running_id", "contig", "start", "end", "num_reads", "num_t2c", "minus_after_first",
             "first_reverse", "combined_strand", "mask51", "site_begin", "site_end")
SITE_FIELDS = ("pos", "t2c", "cov", "order_key")


def have_synth():
    return os.path.exists(os.path.join(abi.LIB_DIR, "libps_synth.so"))


pytestmark = pytest.mark.skipif(not have_synth(), reason="libps_synth.so not built")


def assert_same(got, exp, what):
    assert got["counters"] == exp["counters"], (what, got["counters"], exp["counters"])
    for f in CL_FIELDS:
        assert np.array_equal(got["clusters"][f], exp["clusters"][f]), (what, f)
    for f in SITE_FIELDS:
        assert np.array_equal(got["sites"][f], exp["sites"][f]), (what, f)
    for f in CL_FIELDS:
        assert got["open_cluster"][f] == exp["open_cluster"][f], (what, "open", f)
    for f in SITE_FIELDS:
        assert np.array_equal(got["open_sites"][f], exp["open_sites"][f]), (what, "open sites", f)


def sharded_pileup(oracle, ref, batch, cuts):
    results, carry = [], None
    bounds = [0] + list(cuts) + [batch.n_reads]
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        results.append(oracle.pileup(ref, slice_batch(batch, lo, hi), carry=carry))
        carry = merge_pileup_shards.carry_after(results, carry)
    return merge_pileup_shards(results, bounds[:-1])


@pytest.mark.parametrize("mode,L", [(0, 36), (1, 150)])
def test_slice_batch_roundtrip(oracle, mode, L):
    from parasuite_b200 import synth
    ref = synth.synth_reference(9, [1_500_000], n_run=500)
    batch = synth.synth_reads(ref, 30_000, L, seed=4, mode=mode, threads=2)
    whole = oracle.profile_acc(ref, batch, 176)
    acc = np.zeros_like(whole)
    for lo, hi in ((0, 7), (7, 10_001), (10_001, 30_000)):
        oracle.profile_acc(ref, slice_batch(batch, lo, hi), 176, acc=acc)
    assert np.array_equal(acc, whole)


def test_pileup_halo_merge(oracle):
    from parasuite_b200 import synth
    ref = synth.synth_reference(31, [2_000_000, 700_000], n_run=1000)
    batch = synth.synth_reads(ref, 60_000, 36, seed=6, threads=2)
    whole = oracle.pileup(ref, batch)
    for cuts in ([30_000], [30_007], [1], [59_999], [256, 512, 40_000], list(range(1000, 1010))):
        assert_same(sharded_pileup(oracle, ref, batch, cuts), whole, str(cuts))
    # dense data: one cluster spans several shards
    ref2 = synth.synth_reference(32, [60_000], n_run=0)
    dense = synth.synth_reads(ref2, 40_000, 36, seed=7, threads=2)
    whole2 = oracle.pileup(ref2, dense)
    assert_same(sharded_pileup(oracle, ref2, dense, [100, 150, 20_000]), whole2, "dense")


def test_shard_ranges():
    for n, w in ((10_000_000, 8), (1000, 8), (257, 2), (0, 4)):
        r = shard_ranges(n, w)
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
        assert all(lo % abi.PS_TILE_READS == 0 for lo, _ in r)


def _worker(rank, world, port, q):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    repo = os.path.dirname(here)
    for p in (os.path.join(repo, "para-suite_b200"), os.path.join(repo, "oracle")):
        sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import oracle_lib
    from parasuite_b200 import synth
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    ref = synth.synth_reference(41, [1_000_000], n_run=500)
    batch = synth.synth_reads(ref, 50_000, 36, seed=8, threads=1)
    lo, hi = shard_ranges(batch.n_reads, world)[rank]
    # profile: per-rank partial count vector, one sum all-reduce (the data-path collective of the profile)
    acc = oracle_lib.profile_acc(ref, batch, 51, first=lo, count=hi - lo, ordinal0=0)
    t = torch.from_numpy(acc)
    dist.all_reduce(t)
    # pileup: exclusive scan of one (contig, end) pair per shard, then gather + merge on rank 0
    shard = slice_batch(batch, lo, hi)
    carry = None
    if rank > 0:
        obj = [None]
        dist.recv_object_list(obj, src=rank - 1)
        carry = obj[0]
    res = oracle_lib.pileup(ref, shard, carry=carry)
    nxt = merge_pileup_shards.carry_after([res], carry)
    if rank + 1 < world:
        dist.send_object_list([nxt], dst=rank + 1)
    gathered = [None] * world
    dist.gather_object(res, gathered if rank == 0 else None, dst=0)
    if rank == 0:
        whole_acc = oracle_lib.profile_acc(ref, batch, 51)
        merged = merge_pileup_shards(gathered, [r[0] for r in shard_ranges(batch.n_reads, world)])
        whole = oracle_lib.pileup(ref, batch)
        ok = bool(np.array_equal(t.numpy(), whole_acc))
        try:
            assert_same(merged, whole, "gloo")
        except AssertionError as e:
            ok = False
            q.put(repr(e))
        q.put(ok)
    dist.destroy_process_group()


def test_world2_gloo(oracle):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    out = []
    while not q.empty():
        out.append(q.get())
    assert out and out[-1] is True, out


# ---- parallel carry: exclusive prefix-max of one (contig, end) pair per shard (parasuite_b200.distributed) ----------
def np_max_key(ref, batch):
    """(contig, end) maximum over the kept records of a single-M-op batch (numpy stand-in for ps_pileup_max_key)."""
    if batch.n_reads == 0:
        return None
    fl = batch.meta >> 24
    kept = (fl & (abi.PS_RF_UNMAPPED | abi.PS_RF_POS_ZERO)) == 0
    if not kept.any():
        return None
    g = batch.ref_start[kept].astype(np.int64)
    contig = np.searchsorted(ref.contig_off.astype(np.int64), g, side="right") - 1
    start = g - ref.contig_off.astype(np.int64)[contig] + 1
    R = (batch.cigar[: batch.n_reads][kept] >> 4).astype(np.int64)
    end = start + R - 1
    k = int(np.lexsort((end, contig))[-1])
    return int(contig[k]), int(end[k])


def test_prefix_max_carry_equals_sequential_carry(oracle):
    from parasuite_b200 import synth
    from parasuite_b200.distributed import exclusive_prefix_max
    assert exclusive_prefix_max([(0, 5), None, (0, 3), (1, 1)]) == [None, (0, 5), (0, 5), (0, 5)]
    for ref, batch in ((synth.synth_reference(31, [2_000_000, 700_000], n_run=1000), None),
                       (synth.synth_reference(32, [60_000], n_run=0), None)):
        batch = synth.synth_reads(ref, 40_000, 36, seed=6, threads=2)
        whole = oracle.pileup(ref, batch)
        for cuts in ([20_000], [1], [39_999], [100, 150, 20_000], [256, 512, 30_000, 30_001]):
            bounds = [0] + cuts + [batch.n_reads]
            shards = [slice_batch(batch, lo, hi) for lo, hi in zip(bounds[:-1], bounds[1:])]
            carries = exclusive_prefix_max([np_max_key(ref, s) for s in shards])
            results = [oracle.pileup(ref, s, carry=c) for s, c in zip(shards, carries)]     # all shards independent
            assert_same(merge_pileup_shards(results, bounds[:-1]), whole, f"parallel carry {cuts}")


def _worker_parallel(rank, world, port, q):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    repo = os.path.dirname(here)
    for p in (os.path.join(repo, "para-suite_b200"), os.path.join(repo, "oracle"), here):
        sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import oracle_lib
    from parasuite_b200 import synth
    from parasuite_b200.distributed import allreduce_profile, sharded_pileup
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    ref = synth.synth_reference(43, [400_000, 300_000], n_run=500)
    batch = synth.synth_reads(ref, 45_000, 36, seed=9, threads=1)
    lo, hi = shard_ranges(batch.n_reads, world)[rank]
    shard = slice_batch(batch, lo, hi)
    acc = allreduce_profile(torch.from_numpy(oracle_lib.profile_acc(ref, batch, 51, first=lo, count=hi - lo, ordinal0=0)))
    res, merged = sharded_pileup(shard, lambda b: np_max_key(ref, b), lambda b, c: oracle_lib.pileup(ref, b, carry=c), lo)
    if rank == 0:
        ok = bool(np.array_equal(acc.numpy(), oracle_lib.profile_acc(ref, batch, 51)))
        try:
            assert_same(merged, oracle_lib.pileup(ref, batch), "gloo parallel carry")
        except AssertionError as e:
            ok = False
            q.put(repr(e))
        q.put(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_world_gloo_parallel_carry(oracle, world):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_parallel, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    out = []
    while not q.empty():
        out.append(q.get())
    assert out and out[-1] is True, out
