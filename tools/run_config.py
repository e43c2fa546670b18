#!/usr/bin/env python
"""BASELINE.json configs 3, 4 and 5 at their stated size (one process per GPU; `--configs 3,4,5` share one reference):

  python tools/run_config.py --configs 5                                       (1 GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
         tools/run_config.py --configs 3,4,5                                    (8 GPUs: 200 M / 500 M reads in all)

  3  error profile, 50-nt reads vs a 3.1 Gb reference in 25 contigs with GRCh38 lengths, READ-BATCH sharded: every rank
     takes --reads3 reads (25 M at N = 8 -> 200 M) in batches of 4 M, one NCCL all-reduce of the count vector per pass.
     Checks: a >= 10 M-read prefix of every rank's share against the oracle, bit-exact; on the full run the Java-int
     wrap-around (totalBasesChecked > 2^32: Q8), sum of cells == totalBasesChecked == sum of quality counts, per-rank
     partial vectors sum to the all-reduced one.
  4  T>C pileup, 36-nt genome-sorted reads vs the same reference, REGION sharded: rank r holds --reads4 reads (62.5 M at
     N = 8 -> 500 M) of the r-th of N equal slices of the genome (cuts fall inside contigs), carry-in by all-gathered keys
     on the device, head partials exchanged and merged.  Checks: the closed clusters of a >= 10 M-read prefix of every
     rank's shard against the oracle (same carry-in), bit-exact; conservation on the full shard (reads in clusters ==
     records kept, T>C events in sites == events in clusters, coverage >= T>C, position order, consecutive ids).
  5  error profile, 150-nt reads with indels and soft clips (dense CIGAR/MD), per-read-position profile, maxLen 176:
     --reads5 per rank, whole input against the oracle, bit-exact.
One JSON object on rank 0's stdout."""
import argparse
import json
import os
import sys
import time

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(R, "para-suite_b200"), os.path.join(R, "oracle"), R):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--configs", default="5")
ap.add_argument("--reads3", type=int, default=25_000_000)
ap.add_argument("--reads4", type=int, default=62_500_000)
ap.add_argument("--reads5", type=int, default=4_000_000)
ap.add_argument("--prefix", type=int, default=10_000_000)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--small-ref", action="store_true", help="debug: 1/100 of the GRCh38 contig lengths")
args = ap.parse_args()
configs = [int(x) for x in args.configs.split(",")]

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import oracle_lib  # noqa: E402
from bench import BOUNDARY_WORDS, merge_boundary, pack_head, peaks, shard_key, unpack_head  # noqa: E402
from parasuite_b200 import abi, synth  # noqa: E402
from parasuite_b200.distributed import exclusive_prefix_max, gather_keys_device  # noqa: E402
from parasuite_b200.runtime import Context, DeviceBatch  # noqa: E402
from parasuite_b200.sharding import slice_batch  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        dist.init_process_group("nccl", device_id=dev)
        dist.all_reduce(torch.zeros(1, device=dev))
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
oracle_lib.build()
peak, _ = peaks()
cores = max(1, (os.cpu_count() or 1) // world)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
side = torch.cuda.Stream(device=dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def all_true(flag: bool) -> bool:
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item())


def max_over_ranks(x: float) -> float:
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x: int) -> int:
    t = torch.tensor([x], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t)
    return int(t.item())


t_ref = time.perf_counter()
lengths = [max(1000, x // 100) for x in synth.GRCH38_LENGTHS] if args.small_ref else synth.GRCH38_LENGTHS
ref = synth.synth_reference(0x5EED0001, lengths, names=synth.GRCH38_NAMES)
ctx = Context(local_rank)
ctx.upload_reference(ref)
out = {"n_gpus": world, "reference": {"contigs": len(lengths), "bases": ref.n_bases, "lengths": "GRCh38 primary assembly" +
                                        (" / 100 (debug)" if args.small_ref else ""),
                                        "hbm_bytes_per_gpu": int(ref.seq2.nbytes + ref.inv.nbytes),
                                        "synth_and_upload_s": None}, "peak_gbs": peak, "configs": {}}
out["reference"]["synth_and_upload_s"] = time.perf_counter() - t_ref


def timed_profile(dbatches, max_len, steps):
    def one():
        ctx.profile_begin(max_len)
        for d in dbatches:
            ctx.profile_batch_device(d, stream.cuda_stream)
        if world > 1:
            dist.all_reduce(ctx.profile_acc_tensor())
        return ctx.profile_end()
    for _ in range(2):
        res = one()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        res = one()
    e1.record(stream)
    barrier()
    return res, max_over_ranks(e0.elapsed_time(e1)) / steps


def run_profile_config(name, n_reads, L, mode, max_len, batch_reads, prefix):
    t0 = time.perf_counter()
    b = synth.synth_reads(ref, n_reads, L, seed=0x5EED0300 + 16 * mode + rank, mode=mode)
    gen_s = time.perf_counter() - t0
    cuts = list(range(0, b.n_reads, batch_reads)) + [b.n_reads]
    parts = [slice_batch(b, lo, hi) for lo, hi in zip(cuts[:-1], cuts[1:])] if len(cuts) > 2 else [b]
    dbs = [DeviceBatch(p, dev) for p in parts]
    res, ms = timed_profile(dbs, max_len, args.steps)
    total = sum_over_ranks(b.n_reads)
    alg = sum_over_ranks(b.algorithmic_bytes(with_qual=True))
    # ---- local (un-reduced) vector of this rank, and the oracle on a prefix of its share -----------------------------
    ctx.profile_begin(max_len)
    for d in dbs:
        ctx.profile_batch_device(d, stream.cuda_stream)
    local = ctx.profile_end()["wide"]
    n_pre, k_pre = 0, 0
    while k_pre < len(parts) and n_pre < prefix:
        n_pre += parts[k_pre].n_reads
        k_pre += 1
    ctx.profile_begin(max_len)
    for d in dbs[:k_pre]:
        ctx.profile_batch_device(d, stream.cuda_stream)
    got_pre = ctx.profile_end()["wide"]
    exp_pre = oracle_lib.profile_acc(ref, b, max_len, threads=cores, first=0, count=n_pre)
    ok_prefix = all_true(bool(np.array_equal(got_pre, exp_pre)))
    lt = torch.from_numpy(local.copy()).to(dev)
    if world > 1:
        dist.all_reduce(lt)
    ok_partials = all_true(bool(np.array_equal(lt.cpu().numpy(), res["wide"])))
    w = res["wide"]
    conv, qcnt, ctr = w[:16 * max_len], w[16 * max_len + 16:16 * max_len + 32], w[16 * max_len + 32 + 2 * max_len:][:8]
    def wrap(v):        # Java int: two's-complement wrap of the 64-bit sum
        return int(np.array([v], dtype=np.int64).astype(np.uint32).view(np.int32)[0])
    inv = {
        "cells_eq_total_bases_checked": bool(conv.sum() == ctr[7]),
        "reads_processed_eq_reads": bool(ctr[0] + ctr[1] + ctr[2] + ctr[3] == total),
        "int32_outputs_are_wrapped_int64": bool(int(res["counters"][7]) == wrap(ctr[7]) and
                                                all(int(a) == wrap(v) for a, v in zip(res["quality_per_mismatch"].reshape(-1), w[16 * max_len:16 * max_len + 16]))),
        "total_bases_checked": int(ctr[7]), "total_bases_checked_java_int": int(res["counters"][7]),
        "wraps": bool(ctr[7] >= 2 ** 31),
    }
    if mode == 0:       # reads without I / D: every counted base has a quality
        inv["quality_counts_eq_total_bases_checked"] = bool(qcnt.sum() == ctr[7])
    entry = {"what": name, "reads_total": total, "reads_per_gpu": b.n_reads, "read_len": L, "max_read_length": max_len,
             "batches_per_gpu": len(parts), "ms_per_pass": ms, "reads_per_s": total / (ms * 1e-3),
             "algorithmic_bytes": alg, "frac_of_hbm_peak_per_gpu": alg / world / (ms * 1e-3) / 1e9 / peak,
             "parity": {"oracle_prefix_reads_per_gpu": n_pre, "prefix_bit_exact_every_rank": ok_prefix,
                        "partials_sum_to_allreduced": ok_partials, "invariants": inv},
             "generate_s": gen_s}
    entry["parity"]["ok"] = bool(ok_prefix and ok_partials and all(v for k, v in inv.items() if isinstance(v, bool) and k != "wraps"))
    del dbs
    return entry


def run_pileup_config(n_reads, L, prefix):
    t0 = time.perf_counter()
    lo, hi = ref.n_bases * rank // world, ref.n_bases * (rank + 1) // world
    b = synth.synth_reads(ref, n_reads, L, seed=0x5EED0400 + rank, region=(lo, hi))
    gen_s = time.perf_counter() - t0
    d = DeviceBatch(b, dev)

    def one(keep=False):
        keys = None
        if world > 1:
            with torch.cuda.stream(side):
                keys = gather_keys_device(ctx.pileup_max_key_tensor(d, side.cuda_stream))
            stream.wait_stream(side)
        ck = (keys.data_ptr(), rank) if keys is not None else None
        with ctx.pileup_run(d, first_running_id=1, carry_keys=ck, stream=stream.cuda_stream, defer=True) as h:
            c = h.counters
            return (h.fetch(boundary=True), keys) if keep else (c, keys)
    for _ in range(2):
        one()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        one()
    e1.record(stream)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    stage = [float(x) for x in ctx.pileup_stage_ms()]
    res, _ = one(keep=True)
    total = sum_over_ranks(b.n_reads)
    # ---- conservation on the full shard -------------------------------------------------------------------------------
    cl, si = res["clusters"], res["sites"]
    kept = int((((b.meta[:b.n_reads] >> 24) & abi.PS_RF_UNMAPPED) == 0).sum())
    in_clusters = int(cl["num_reads"].sum())
    for k in ("open_cluster", "head_partial"):
        if res[k] is not None:
            in_clusters += int(res[k]["num_reads"])
    owner = np.repeat(np.arange(len(cl)), (cl["site_end"] - cl["site_begin"]).astype(np.int64))
    same = owner[1:] == owner[:-1]
    inv = {
        "reads_in_clusters_eq_kept": bool(in_clusters == kept == res["counters"]["num_reads_processed"]),
        "site_events_eq_cluster_events": bool(int(si["t2c"].sum()) == int(cl["num_t2c"].sum())),
        "coverage_ge_t2c": bool((si["cov"] >= si["t2c"]).all() and (si["t2c"] >= 1).all()),
        "sites_in_position_order": bool((si["pos"][1:][same] > si["pos"][:-1][same]).all()),
        "site_ranges_contiguous": bool(len(cl) == 0 or (int(cl["site_end"][-1]) == len(si) and
                                                        (cl["site_begin"][1:] == cl["site_end"][:-1]).all())),
        "ids_consecutive_from_2": bool((cl["running_id"] == np.arange(2, 2 + len(cl))).all()),
    }
    # ---- oracle on a prefix of the shard, with the shard's true carry-in -------------------------------------------------
    carry = None
    if world > 1:
        keys_all = [None] * world
        dist.all_gather_object(keys_all, shard_key(ref, b))
        carry = exclusive_prefix_max(keys_all)[rank]
    n_pre = min(b.n_reads, prefix)
    exp = oracle_lib.pileup(ref, slice_batch(b, 0, n_pre), carry=carry)
    ne = len(exp["clusters"])
    ok = ne > 0 and ne <= len(cl)
    if ok:
        for f in ("first_read", "running_id", "contig", "start", "end", "num_reads", "num_t2c", "minus_after_first",
                  "first_reverse", "combined_strand", "mask51", "site_begin", "site_end"):
            ok = ok and bool(np.array_equal(cl[f][:ne], exp["clusters"][f]))
        ns = int(exp["clusters"]["site_end"][-1])
        for f in ("pos", "t2c", "cov", "order_key"):
            ok = ok and bool(np.array_equal(si[f][:ns], exp["sites"][f][:ns]))
        ok = ok and (res["head_partial"] is None) == (exp["head_partial"] is None)
        if ok and res["head_partial"] is not None:
            ok = all(res["head_partial"][f] == exp["head_partial"][f] for f in ("end", "num_reads", "num_t2c", "mask51"))
    ok_prefix = all_true(bool(ok))
    # ---- halo merge: every head partial goes to the rank in front of it ----------------------------------------------------
    merged_reads = None
    if world > 1:
        mine = torch.from_numpy(pack_head(res)).to(dev)
        got = torch.empty(world * BOUNDARY_WORDS, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(got, mine)
        g = got.cpu().numpy()
        nxt = unpack_head(g[(rank + 1) * BOUNDARY_WORDS:(rank + 2) * BOUNDARY_WORDS]) if rank + 1 < world else None
        oc, _ = merge_boundary(res, nxt, b.n_reads)
        own_open = 0 if res["open_cluster"] is None else int(res["open_cluster"]["num_reads"])
        merged_reads = 0 if oc is None else int(oc["num_reads"])
        inv["merged_boundary_reads_add_up"] = all_true(merged_reads == own_open + (int(nxt[0]["num_reads"]) if nxt else 0))
        inv["head_partials_in_the_job"] = sum_over_ranks(1 if res["head_partial"] is not None else 0)
    n_cl, n_sites = sum_over_ranks(len(cl)), sum_over_ranks(len(si))
    covered = sum_over_ranks(int((cl["end"].astype(np.int64) - cl["start"].astype(np.int64) + 1).sum()))
    alg = sum_over_ranks(b.algorithmic_bytes(with_qual=False)) + 8 * covered + 32 * n_cl
    inv_ok = all_true(all(v for v in inv.values() if isinstance(v, bool)))
    entry = {"what": "config 4: T>C pileup, genome-sorted 36-nt reads, region sharded (standalone pileup: the kernels decode "
                     "the reads themselves)", "reads_total": total, "reads_per_gpu": b.n_reads, "read_len": L,
             "ms_per_pass": ms, "reads_per_s": total / (ms * 1e-3), "stage_ms_rank0": {"flag": stage[0], "cluster": stage[1], "compact": stage[2]},
             "clusters": n_cl, "sites": n_sites, "algorithmic_bytes": alg,
             "frac_of_hbm_peak_per_gpu": alg / world / (ms * 1e-3) / 1e9 / peak,
             "parity": {"oracle_prefix_reads_per_gpu": n_pre, "prefix_clusters_bit_exact_every_rank": ok_prefix,
                        "invariants_rank0": inv, "invariants_every_rank": inv_ok, "ok": bool(ok_prefix and inv_ok)},
             "generate_s": gen_s}
    del d
    return entry


for c in configs:
    if c == 3:
        e = run_profile_config("config 3: error profile, 50-nt reads (50M), read-batch sharded", args.reads3, 50, 0, 51, 4_000_000, args.prefix)
    elif c == 5:
        e = run_profile_config("config 5: error profile, 150-nt reads with indels and soft clips (dense CIGAR), per-read-position profile",
                               args.reads5, 150, 1, 176, 4_000_000, args.reads5)
    elif c == 4:
        e = run_pileup_config(args.reads4, 36, args.prefix)
    else:
        continue
    out["configs"][str(c)] = e
    torch.cuda.empty_cache()
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
ctx.close()
