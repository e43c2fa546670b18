#!/usr/bin/env python
"""Kernel-level timings of one library build (PARASUITE_B200_LIB selects a variant under para-suite_b200/lib/):
profile kernel with / without T>C mask words, pileup stages with / without them, each checked for equality with the other
path and (profile, --check) with the oracle.  One JSON line per run.
  python tools/bench_kernels.py [--reads N] [--len L] [--iters K] [--check]"""
import argparse, json, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(R, "para-suite_b200")); sys.path.insert(0, os.path.join(R, "oracle"))
import numpy as np, torch
from parasuite_b200 import synth
from parasuite_b200.runtime import Context, DeviceBatch

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=10_000_000)
ap.add_argument("--len", type=int, default=36)
ap.add_argument("--ref", type=int, default=100_000_000)
ap.add_argument("--iters", type=int, default=12)
ap.add_argument("--check", action="store_true")
ap.add_argument("--mode", type=int, default=0)
ap.add_argument("--max-len", type=int, default=51)
ap.add_argument("--trim", type=int, default=0, help="cut every read to its own length in [TRIM, len]: ragged batch")
args = ap.parse_args()
peak = 6552.0
try:
    peak = float(json.load(open(os.path.join(R, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
ctx = Context(0)
ref = synth.synth_reference(0x5EED0001, [args.ref])
ctx.upload_reference(ref)
b = synth.synth_reads(ref, args.reads, args.len, seed=0x5EED0002, mode=args.mode)
st = torch.cuda.current_stream().cuda_stream
if args.trim:      # ragged single-M batch: repack + per-read-length fast kernel against the warp-per-read kernel
    b = synth.trim_uniform(b, args.trim, seed=3)
    d = DeviceBatch(b, "cuda:0")
    out = {"reads": args.reads, "len": [args.trim, args.len], "ragged": True}
    by = b.algorithmic_bytes(with_qual=True)
    res = {}
    for name, off in (("ragged_fast", False), ("warp_per_read", True)):
        if off:
            os.environ["PARASUITE_B200_NO_RAGGED_FAST"] = "1"
        else:
            os.environ.pop("PARASUITE_B200_NO_RAGGED_FAST", None)
        ctx.kernel_times_reset(True)
        for _ in range(args.iters if not off else max(3, args.iters // 4)):
            ctx.profile_begin(args.max_len)
            ctx.profile_batch_device(d, st)
            res[name] = ctx.profile_end()
        t = ctx.kernel_times_ms()
        ms = float(np.mean(t[1:] if off else t[3:]))
        out[name] = {"ms": ms, "reads_per_s": args.reads / ms * 1e3, "frac_of_hbm_peak": by / (ms * 1e-3) / 1e9 / peak}
    out["equal"] = bool(np.array_equal(res["ragged_fast"]["wide"], res["warp_per_read"]["wide"]))
    if args.check:
        import oracle_lib
        oracle_lib.build()
        acc = oracle_lib.profile_acc(ref, b, args.max_len, threads=os.cpu_count() or 1)
        out["parity"] = bool(np.array_equal(acc, res["ragged_fast"]["wide"]))
    print(json.dumps(out), flush=True)
    ctx.close()
    sys.exit(0)
d = DeviceBatch(b, "cuda:0")
out = {"lib": os.environ.get("PARASUITE_B200_LIB", "default"), "reads": args.reads, "len": args.len, "mode": args.mode}
by = b.algorithmic_bytes(with_qual=True)
res = {}
for emit in (False, True):
    ctx.kernel_times_reset(True)
    for _ in range(args.iters):
        ctx.profile_begin(args.max_len, emit_t2c_masks=emit)
        ctx.profile_batch_device(d, st)
        res[emit] = ctx.profile_end()
    ms = float(np.mean(ctx.kernel_times_ms()[3:]))
    out["profile_ms_masks" if emit else "profile_ms"] = ms
    out["profile_frac_masks" if emit else "profile_frac"] = by / (ms * 1e-3) / 1e9 / peak
out["profile_masks_equal"] = bool(np.array_equal(res[False]["wide"], res[True]["wide"]))
if args.check:
    import oracle_lib
    oracle_lib.build()
    acc = oracle_lib.profile_acc(ref, b, args.max_len, threads=os.cpu_count() or 1)
    out["profile_parity"] = bool(np.array_equal(acc, res[True]["wide"]))
if args.mode == 0:
    piles = {}
    for use in (False, True):
        stg = []
        for _ in range(max(4, args.iters // 2)):
            masks = None
            if use:
                ctx.profile_begin(args.max_len, emit_t2c_masks=True)
                ctx.profile_batch_device(d, st)
                masks = ctx.profile_masks()
            with ctx.pileup_run(d, stream=st, masks=masks) as h:
                stg.append(ctx.pileup_stage_ms())
                if len(stg) == 1:
                    piles[use] = h.fetch(boundary=False)
            if use:
                ctx.profile_end()
        s = np.mean(np.asarray(stg[1:]), axis=0)
        out["pileup_ms_masks" if use else "pileup_ms"] = {"flag": float(s[0]), "cluster": float(s[1]), "compact": float(s[2]),
                                                          "sum": float(s.sum())}
    out["pileup_masks_equal"] = bool(np.array_equal(piles[False]["clusters"], piles[True]["clusters"]) and
                                     np.array_equal(piles[False]["sites"], piles[True]["sites"]))
print(json.dumps(out), flush=True)
ctx.close()
