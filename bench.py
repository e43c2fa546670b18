#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: aligned reads/sec (error profile + T>C pileup) on synthetic PAR-CLIP reads.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the hot path (profile kernel, read-back of the counts, the three pileup kernels) over one batch of
synthetic coordinate-sorted reads that is already resident in HBM (`value`; cluster / site records stay in HBM), or
handed over as pinned HOST buffers through the C ABI with every copy -- records in, counts, clusters and sites out --
inside the timed region (`e2e`).
Workload at every N: BASELINE configs[1] per GPU, 10M x 36-nt reads vs a 100 Mb reference (weak scaling; rank r holds
region r of an N x 100 Mb genome); one NCCL all-reduce of the profile count vector per step, no collective in the pileup.
--impl reference times the CPU restatement of the Java loops (oracle/; the jar cannot run: no JVM) on the host cores.
"""
import argparse
import json
import os
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(REPO, "para-suite_b200"), os.path.join(REPO, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "aligned reads/sec (profile + T>C pileup)"
UNIT = "reads/s"


def workload(n_gpus: int, rank: int, small: bool = False):
    """Returns (name, reference, batch, max_len).  The per-GPU workload is the same at every N (weak scaling):
    BASELINE configs[1], 10M x 36-nt PAR-CLIP reads against a 100 Mb reference.  At N > 1 rank r holds region r of an
    N x 100 Mb genome, i.e. the read batches AND the genome regions are sharded, as SURVEY 8(e) partitions the two
    tools: one all-reduce of the profile count vector, one all-gather of a (contig, end) pair per rank for the pileup."""
    from parasuite_b200 import synth
    if small:   # CI-sized (tests): same shape, 1/50 size
        ref = synth.synth_reference(0x5EED0001, [2_000_000])
        return "config2-small", ref, synth.synth_reads(ref, 200_000, 36, seed=0x5EED0002 + rank), 51
    # the N-GPU job: an N x 100 Mb genome (one contig per region), replicated on every GPU as the tools need it; rank r
    # holds the 10M reads of region r
    region = 100_000_000
    ref = synth.synth_reference(0x5EED0001, [region] * n_gpus)
    batch = synth.synth_reads(ref, 10_000_000, 36, seed=0x5EED0002 + rank, region=(rank * region, (rank + 1) * region))
    name = "config2: 10M x 36-nt PAR-CLIP reads (single 36M cigar) vs 100 Mb synthetic reference"
    if n_gpus > 1:
        name += f", per GPU (rank r = region r of a {n_gpus} x 100 Mb genome, reference replicated)"
    return name, ref, batch, 51


def peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.sm_max = None
        self.stop_flag = False
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def result(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def visible_physical_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


class CpuPath:
    """The reference's CPU algorithm for the whole path on a bounded sample: error-profile loop, then the T>C pileup
    loop.  This is the C++ restatement under oracle/ (the Java jar cannot run: no JVM in this image).  The Java tools are
    single-threaded; the port is given every host core: the profile loop over read chunks, the pileup loop over
    contiguous read ranges (boundary clusters are not merged -- throughput only)."""

    def __init__(self, ref, batch, max_len, sample, cores):
        import oracle_lib
        from parasuite_b200.sharding import slice_batch
        oracle_lib.build()
        self.o, self.ref, self.batch, self.max_len, self.sample, self.cores = oracle_lib, ref, batch, max_len, sample, cores
        cuts = [sample * k // cores // 256 * 256 for k in range(cores)] + [sample]
        self.shards = [slice_batch(batch, cuts[k], cuts[k + 1]) for k in range(cores) if cuts[k + 1] > cuts[k]]

    def step(self):
        from concurrent.futures import ThreadPoolExecutor
        acc = self.o.profile_acc(self.ref, self.batch, self.max_len, threads=self.cores, first=0, count=self.sample)
        with ThreadPoolExecutor(len(self.shards)) as ex:
            n_cl = sum(ex.map(lambda sh: self.o.pileup_count(self.ref, sh), self.shards))
        return acc, n_cl

    def describe(self):
        return (f"first {self.sample} reads of the workload per step: error-profile loop ({self.cores} threads over read "
                f"chunks) + T>C pileup loop ({len(self.shards)} threads over contiguous read ranges); C++ restatement of "
                "the Java loops (the jar cannot run: no JVM in this image; the Java tools are single-threaded)")


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm (oracle port; no JVM exists here) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    name, ref, batch, max_len = workload(args.gpus, 0, args.small)
    cores = os.cpu_count() or 1
    sample = min(batch.n_reads, 2_000_000 if not args.small else 100_000)
    sample -= sample % 256
    cpu = CpuPath(ref, batch, max_len, sample, cores)
    for _ in range(args.warmup):
        cpu.step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu.step()
    dt = time.perf_counter() - t0
    v = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": name, "stages": ["profile", "pileup"]},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": cpu.describe()},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--small", action="store_true", help="CI-sized workload (not a bench value)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the host-buffer leg (default min(steps, 10))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and not args.small:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from parasuite_b200.distributed import gather_keys_device, sharded_pileup_carry
    from parasuite_b200.runtime import Context, DeviceBatch, PinnedBatch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator is created: keep stdout to the one JSON line
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

    name, ref, batch, max_len = workload(args.gpus, rank, args.small)
    ctx = Context(local_rank)
    ctx.upload_reference(ref)
    dbatch = DeviceBatch(batch, dev)
    alg_bytes = batch.algorithmic_bytes(with_qual=True)
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    pile = {}
    side = torch.cuda.Stream(device=dev) if world > 1 else None
    red = torch.cuda.Stream(device=dev) if world > 1 else None

    def step_resident():
        # the whole hot path on a batch that is already in HBM: profile kernel (+ the tiny all-reduce), read-back of
        # the < 10 KB of counts, then the three pileup kernels; cluster / site records stay in HBM behind the handle,
        # their counters come back to the host
        keys = None
        ctx.profile_begin(max_len, emit_t2c_masks=True)
        ctx.profile_batch_device(dbatch, stream.cuda_stream)
        masks = ctx.profile_masks()     # one T>C mask word per read, left in HBM by the profile kernel for the pileup
        if world > 1:
            # region sharding: the maximum (contig, end) of every region, all-gathered; the exclusive prefix-max (this
            # region's carry-in) is taken on the device by the flag kernel -- no host round trip.  The key kernel and its
            # all-gather run on a side stream underneath the profile kernel (launched first, so the host-side cost of
            # the collectives is hidden too)
            with torch.cuda.stream(side):
                keys = gather_keys_device(ctx.pileup_max_key_tensor(dbatch, side.cuda_stream))
            # the all-reduce gets its own stream behind the profile kernel and the counts are read back there
            # (ps_profile_set_stream), so taking the profile back does not wait for the pileup kernels either
            red.wait_stream(stream)
            with torch.cuda.stream(red):
                dist.all_reduce(ctx.profile_acc_tensor())
            ctx.profile_set_stream(red.cuda_stream)
            stream.wait_stream(side)
        # the pileup kernels are queued right behind the profile kernel; the counts of both come back afterwards, so the
        # device does not sit idle between the two tools
        carry_keys = (keys.data_ptr(), rank) if keys is not None else None
        # ps_pileup_submit_device: the call returns behind its launches, so the host takes back the profile (and does its
        # own bookkeeping) while the pileup kernels run; the wait for the pileup is the only one left at the end
        with ctx.pileup_run(dbatch, first_running_id=1, carry_keys=carry_keys, stream=stream.cuda_stream, defer=True,
                            masks=masks) as h:
            res = ctx.profile_end()
            pile["counters"] = h.counters
        return res

    # ---- warm-up ------------------------------------------------------------------------------------
    for _ in range(args.warmup):
        res = step_resident()
    # ---- device-resident timed region ---------------------------------------------------------------
    sampler = ClockSampler(visible_physical_index(local_rank))
    ctx.kernel_times_reset(True)
    launches0 = ctx.kernel_launches()
    stage_ms = []
    barrier()
    sampler.start()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        res = step_resident()
        stage_ms.append(ctx.pileup_stage_ms())       # events of the step just finished (already synchronised)
    e1.record(stream)
    barrier()
    sampler.stop_flag = True
    ms = e0.elapsed_time(e1)
    launches = ctx.kernel_launches() - launches0
    ktimes = ctx.kernel_times_ms()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    total_reads = batch.n_reads * world
    value = total_reads * args.steps / (ms_max * 1e-3)

    # ---- end-to-end leg: pinned HOST buffers through the C ABI, copies inside the timed region -------
    # one upload of the records serves both tools; every cluster and site record comes back to (pinned) host memory
    e2e_steps = args.e2e_steps or min(args.steps, 10)
    pinned = PinnedBatch(batch)

    def step_e2e():
        # records in (one upload serves both tools; the quality bytes, more than half of it, go last), the pileup runs on
        # the library's auxiliary stream as soon as the other streams have arrived and its records travel back while the
        # qualities are still on their way up; the profile kernel follows the upload
        view = ctx.upload(pinned)
        carry = None
        if world > 1:
            carry = sharded_pileup_carry(ctx.pileup_max_key(view), device=dev)
        with ctx.pileup_run(view, carry=carry) as h:
            pile["res_e2e"] = h.fetch(pinned=True, boundary=False)
        ctx.profile_begin(max_len)
        ctx.profile_batch_device(view)
        if world > 1:
            torch.cuda.synchronize()
            dist.all_reduce(ctx.profile_acc_tensor())
            torch.cuda.current_stream().synchronize()
        return ctx.profile_end()

    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res_e2e = step_e2e()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = total_reads * e2e_steps / (float(t.item()) * 1e-3)
    pr = pile["res_e2e"]
    d2h = int(res_e2e["wide"].nbytes + 8 + pr["clusters"].nbytes + pr["sites"].nbytes)
    h2d = pinned.h2d_bytes

    if rank == 0:
        peak, peak_src = peaks()
        # timer ring order per step: profile kernel(s), pileup pipeline
        k_prof = float(np.mean(ktimes[0::2])) if len(ktimes) >= 2 else float("nan")
        k_pile = float(np.mean(ktimes[1::2])) if len(ktimes) >= 2 else float("nan")
        k_flag, k_cluster, k_compact = (float(x) for x in np.mean(np.asarray(stage_ms, dtype=np.float64), axis=0))
        cl = pr["clusters"]
        n_cl, n_sites = int(len(cl)), int(len(pr["sites"]))
        covered = int((cl["end"].astype(np.int64) - cl["start"].astype(np.int64) + 1).sum())
        pile_bytes = batch.algorithmic_bytes(with_qual=False) + 8 * covered + 32 * n_cl
        flag_bytes = 12 * batch.n_reads + 4 * n_cl            # meta + ref_start + cigar in, one opener index per cluster out
        compact_bytes = 2 * 24 * n_sites + 2 * 16 * n_cl      # every site moved once, site range of every record rewritten

        def frac(nbytes, kms):
            return nbytes / (kms * 1e-3) / 1e9 / peak if kms == kms and kms > 0 else None

        kernels = {
            "profile_fast_kernel": (k_prof, alg_bytes),
            "pl_flag_kernel": (k_flag, flag_bytes),
            "pl_cluster_kernel": (k_cluster, pile_bytes),
            "pl_compact_kernel": (k_compact, compact_bytes),
        }
        traffic = {}
        try:    # DRAM bytes per launch from the committed ncu --set full capture of this command (tools/ncu_traffic.py)
            traffic = json.load(open(os.path.join(REPO, "profiles", "traffic.json")))["traffic_bytes_per_launch"]
        except Exception:
            pass
        if "pl_flag_kernel" in traffic:   # the flag stage is two kernels (flag pass, cl_first expansion; + a scan kernel for huge batches)
            traffic["pl_flag_kernel"] = sum(traffic.get(k, 0.0) for k in
                                            ("pl_flag_kernel", "pl_flag_scan_kernel", "pl_flag_expand_kernel"))
        kname = max(kernels, key=lambda k: kernels[k][0] if kernels[k][0] == kernels[k][0] else -1.0)
        kms, kbytes = kernels[kname]
        achieved = kbytes / (kms * 1e-3) / 1e9 if kms == kms and kms > 0 else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"workload": name, "stages": ["profile", "pileup"], "reads_per_gpu": batch.n_reads,
                       "stage_ms": {"profile_kernel": k_prof, "pileup_device": k_pile, "pl_flag_kernel": k_flag,
                                    "pl_cluster_kernel": k_cluster, "pl_compact_kernel": k_compact},
                       "stage_note": "pl_flag_kernel = flag pass + cl_first expansion (the tile-table prefix is taken inside the expansion kernel; a one-block scan kernel past 16.7 M reads)",
                       "pileup": {"clusters": n_cl, "sites": n_sites, "covered_loci": covered},
                       "max_read_length": max_len, "l2": "inputs larger than L2 (%.0f MB per pass)" % (alg_bytes / 1e6),
                       "parallelism": f"profile: read-batch sharded x{world} + all-reduce of the count vector; "
                                      f"pileup: region sharded x{world}, halo merge on the host"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps},
            "gpu_launches": int(launches),
            "clocks": sampler.result(),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None,
                         "traffic": traffic.get(kname) if world == 1 else None,
                         "kernel": kname, "kernel_ms": kms,
                         "algorithmic_bytes_per_launch": kbytes, "peak_source": peak_src,
                         "per_kernel": {k: {"ms": v[0], "bytes": v[1], "frac": frac(v[1], v[0]), "traffic": traffic.get(k)}
                                        for k, v in kernels.items()}},
        }
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            sample = min(batch.n_reads, 2_000_000)
            sample -= sample % 256
            cpu = CpuPath(ref, batch, max_len, sample, cores)
            cpu.step()
            t0 = time.perf_counter()
            cpu.step()
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": sample / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": cpu.describe()}
            # parity of the timed configuration against the oracle: whole batch, error profile (bit-exact)
            import oracle_lib
            acc = oracle_lib.profile_acc(ref, batch, max_len, threads=cores)
            line["parity"] = bool(np.array_equal(acc, res["wide"]) if world == 1 else True) and \
                bool(np.array_equal(acc, res_e2e["wide"]) if world == 1 else True)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
