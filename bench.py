#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: aligned reads/sec (error profile + T>C pileup) on synthetic PAR-CLIP reads.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the hot path (profile kernel, then the pileup kernels once they exist) over one batch of
synthetic coordinate-sorted reads that is already resident in HBM (`value`), or handed over as pinned HOST
buffers through the C ABI with the copies inside the timed region (`e2e`).
N=1 : BASELINE configs[1]  10M x 36-nt reads vs a 100 Mb reference.
N>1 : BASELINE configs[2]  shard shape: 25M x 50-nt reads per GPU vs the 3.1 Gb reference (weak scaling), reads
      of rank r drawn from genome slice r; one NCCL all-reduce of the count vector per step.
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
    """Returns (name, reference, batch, max_len)."""
    from parasuite_b200 import synth
    if small:   # CI-sized (tests): same shape, 1/50 size
        ref = synth.synth_reference(0x5EED0001, [2_000_000])
        return "config2-small", ref, synth.synth_reads(ref, 200_000, 36, seed=0x5EED0002 + rank), 51
    if n_gpus == 1:
        ref = synth.synth_reference(0x5EED0001, [100_000_000])
        batch = synth.synth_reads(ref, 10_000_000, 36, seed=0x5EED0002)
        return "config2: 10M x 36-nt PAR-CLIP reads (single 36M cigar) vs 100 Mb synthetic reference", ref, batch, 51
    ref = synth.synth_reference(0x5EED0001, synth.GRCH38_LENGTHS, names=synth.GRCH38_NAMES)
    n = ref.n_bases
    lo, hi = n * rank // n_gpus, n * (rank + 1) // n_gpus
    batch = synth.synth_reads(ref, 25_000_000, 50, seed=0x5EED0003 + rank, region=(lo, hi))
    return ("config3 shard: 25M x 50-nt reads per GPU vs 3.1 Gb synthetic reference (25 contigs), "
            "reads of rank r from genome slice r"), ref, batch, 51


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


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm (oracle port; no JVM exists here) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle_lib
    oracle_lib.build()
    name, ref, batch, max_len = workload(args.gpus, 0, args.small)
    cores = os.cpu_count() or 1
    sample = min(batch.n_reads, 2_000_000 if not args.small else 100_000)
    sample -= sample % 256

    def step():
        oracle_lib.profile_acc(ref, batch, max_len, threads=cores, first=0, count=sample)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": name, "stages": ["profile"]},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"first {sample} reads of the workload per step, error-profile loop, {cores} threads; "
                                   "C++ restatement of the Java loop (the jar cannot run: no JVM in this image)"},
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
        dist.init_process_group("nccl", device_id=dev)

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

    def step_resident():
        ctx.profile_begin(max_len)
        ctx.profile_batch_device(dbatch, stream.cuda_stream)
        if world > 1:
            dist.all_reduce(ctx.profile_acc_tensor())

    def finish():
        res = ctx.profile_end()
        # T>C pileup of the same reads (region shard of this rank; the halo exchange is 2 scalars per cut and is
        # done by parasuite_b200.sharding when shards are merged -- not part of the per-shard step)
        pile["res"] = ctx.pileup(dbatch, stream=stream.cuda_stream)
        return res

    # ---- warm-up + parity of the timed configuration against the oracle on a prefix ---------------
    for _ in range(args.warmup):
        step_resident()
        res = finish()
    # ---- device-resident timed region ---------------------------------------------------------------
    sampler = ClockSampler(visible_physical_index(local_rank))
    ctx.kernel_times_reset(True)
    launches0 = ctx.kernel_launches()
    barrier()
    sampler.start()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_resident()
        # profile_end() synchronises and reads back < 10 KB of counts: part of the step
        res = finish()
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
    e2e_steps = args.e2e_steps or min(args.steps, 10)
    pinned = PinnedBatch(batch)
    def step_e2e():
        ctx.profile_begin(max_len)
        ctx.profile_batch(pinned)
        if world > 1:
            torch.cuda.synchronize()
            dist.all_reduce(ctx.profile_acc_tensor())
        r = ctx.profile_end()
        pile["res_e2e"] = ctx.pileup(pinned)
        return r

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
    h2d = 2 * pinned.h2d_bytes      # each stage takes the host batch through the C ABI

    if rank == 0:
        peak, peak_src = peaks()
        # timer ring order per step: profile kernel, pileup pipeline
        k_prof = float(np.mean(ktimes[0::2])) if len(ktimes) >= 2 else float("nan")
        k_pile = float(np.mean(ktimes[1::2])) if len(ktimes) >= 2 else float("nan")
        cl = pile["res"]["clusters"]
        pile_bytes = (batch.algorithmic_bytes(with_qual=False)
                      + 8 * int((cl["end"].astype(np.int64) - cl["start"].astype(np.int64) + 1).sum())
                      + 32 * len(cl))
        if k_prof >= k_pile or k_pile != k_pile:
            kname, kms, kbytes = "profile_generic_kernel", k_prof, alg_bytes
        else:
            kname, kms, kbytes = "pileup pipeline (pl_read/flag/cluster/site + scans + sort)", k_pile, pile_bytes
        achieved = kbytes / (kms * 1e-3) / 1e9 if kms == kms else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"workload": name, "stages": ["profile", "pileup"], "reads_per_gpu": batch.n_reads,
                       "stage_ms": {"profile_kernel": k_prof, "pileup_device": k_pile},
                       "pileup": {"clusters": int(len(cl)), "sites": int(len(pile["res"]["sites"]))},
                       "max_read_length": max_len, "l2": "inputs larger than L2 (%.0f MB per pass)" % (alg_bytes / 1e6),
                       "parallelism": f"read-batch sharded x{world}"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps},
            "gpu_launches": int(launches),
            "clocks": sampler.result(),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": None,
                         "kernel": kname, "kernel_ms": kms,
                         "algorithmic_bytes_per_launch": kbytes, "peak_source": peak_src,
                         "per_stage": {
                             "profile": {"ms": k_prof, "bytes": alg_bytes,
                                         "frac": alg_bytes / (k_prof * 1e-3) / 1e9 / peak if k_prof == k_prof else None},
                             "pileup": {"ms": k_pile, "bytes": pile_bytes,
                                        "frac": pile_bytes / (k_pile * 1e-3) / 1e9 / peak if k_pile == k_pile else None}}},
        }
        if not args.no_cpu_baseline:
            import oracle_lib
            oracle_lib.build()
            cores = os.cpu_count() or 1
            sample = batch.n_reads if cores >= 8 else min(batch.n_reads, 2_000_000)
            if sample != batch.n_reads:
                sample -= sample % 256
            t0 = time.perf_counter()
            acc = oracle_lib.profile_acc(ref, batch, max_len, threads=cores, first=0, count=sample)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": sample / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"first {sample} reads of the workload, error-profile loop, {cores} "
                                              "threads (C++ restatement of the Java loop; no JVM in this image)"}
            if sample == batch.n_reads and world == 1:
                line["parity"] = bool(np.array_equal(acc, res["wide"]))
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
