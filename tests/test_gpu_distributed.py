"""GPU, multi-process: the N>1 path with the real kernels -- two ranks (one process each; on a single-GPU box both
use cuda:0), read-batch sharded profile + all-reduce, region-sharded pileup with the all-gathered prefix-max carry and
the halo merge, against the oracle on the whole stream."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
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
    from parasuite_b200.runtime import Context, DeviceBatch
    from parasuite_b200.sharding import shard_ranges, slice_batch
    from test_sharding_cpu import assert_same
    dev = rank % torch.cuda.device_count()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    ref = synth.synth_reference(51, [3_000_000, 2_000_000], n_run=1000)
    batch = synth.synth_reads(ref, 300_000, 36, seed=12, threads=2)
    lo, hi = shard_ranges(batch.n_reads, world)[rank]
    if rank == 0:
        hi = lo + (hi - lo) // 2 + 77            # unequal shards, cut in the middle of a tile
    elif rank == 1:
        lo = shard_ranges(batch.n_reads, world)[0][0] + (shard_ranges(batch.n_reads, world)[0][1]) // 2 + 77
    ctx = Context(dev)
    ctx.upload_reference(ref)
    shard = slice_batch(batch, lo, hi)
    d = DeviceBatch(shard, f"cuda:{dev}")
    ctx.profile_begin(51)
    ctx.profile_batch_device(d)
    acc = allreduce_profile(torch.from_numpy(ctx.profile_end()["wide"].copy()))
    res, merged = sharded_pileup(d, ctx.pileup_max_key, lambda b, c: ctx.pileup(b, carry=c), lo)
    if rank == 0:
        ok = bool(np.array_equal(acc.numpy(), oracle_lib.profile_acc(ref, batch, 51, threads=4)))
        try:
            assert_same(merged, oracle_lib.pileup(ref, batch), "2 ranks, real kernels")
        except AssertionError as e:
            ok = False
            q.put(repr(e))
        q.put(ok)
    ctx.close()
    dist.destroy_process_group()


def test_two_ranks_real_kernels(oracle):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    out = []
    while not q.empty():
        out.append(q.get())
    assert out and out[-1] is True, out
