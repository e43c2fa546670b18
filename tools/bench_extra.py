#!/usr/bin/env python
"""Secondary measurements (not the bench.py line): other read shapes of BASELINE.json's configs and the file path.
  python tools/bench_extra.py [--bam-reads N]
Prints one JSON object: kernel time / reads/s / fraction of the HBM roofline per shape, and BAM end-to-end reads/s."""
import argparse, json, os, sys, tempfile, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(R, "para-suite_b200")); sys.path.insert(0, os.path.join(R, "oracle"))
import numpy as np, torch
from parasuite_b200 import synth
from parasuite_b200.runtime import Context, DeviceBatch

ap = argparse.ArgumentParser()
ap.add_argument("--bam-reads", type=int, default=300_000)
ap.add_argument("--bam-repeat", type=int, default=10, help="a second file with the records repeated this many times (profile only)")
args = ap.parse_args()
peak = 6552.0
try:
    peak = float(json.load(open(os.path.join(R, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
out = {"peak_gbs": peak, "shapes": {}}
ctx = Context(0)
ref = synth.synth_reference(0x5EED0001, [100_000_000])
ctx.upload_reference(ref)
stream = torch.cuda.current_stream().cuda_stream
for name, n, L, mode, max_len, pile in (("config2 36-nt 36M", 10_000_000, 36, 0, 51, True),
                                        ("config3 50-nt 50M", 10_000_000, 50, 0, 51, True),
                                        ("config5 150-nt dense cigar", 4_000_000, 150, 1, 176, False)):
    b = synth.synth_reads(ref, n, L, seed=0x5EED0002, mode=mode)
    d = DeviceBatch(b, "cuda:0")
    ctx.kernel_times_reset(True)
    for _ in range(6):
        ctx.profile_begin(max_len); ctx.profile_batch_device(d, stream); ctx.profile_end()
    ms = float(np.mean(ctx.kernel_times_ms()[2:]))
    by = b.algorithmic_bytes(with_qual=True)
    e = {"reads": n, "profile_ms": ms, "profile_reads_per_s": n / ms * 1e3, "profile_bytes": by,
         "profile_frac": by / (ms * 1e-3) / 1e9 / peak}
    if pile:
        st = []
        for _ in range(5):
            with ctx.pileup_run(d, stream=stream) as h:
                st.append(ctx.pileup_stage_ms())
        st = np.mean(np.asarray(st[1:]), axis=0)
        e["pileup_ms"] = {"flag": float(st[0]), "cluster": float(st[1]), "compact": float(st[2])}
        e["pileup_reads_per_s"] = n / float(st.sum()) * 1e3
    out["shapes"][name] = e
    del d
# ---- file path: FASTA + BAM -> ps_profile_bam / ps_pileup_bam (host decode inside) ----
from parasuite_b200.bamio import batch_to_records, write_bam, write_fasta
sref = synth.synth_reference(7, [5_000_000], n_run=1000)
sb = synth.synth_reads(sref, args.bam_reads, 36, seed=8)
codes = np.zeros(sref.n_bases, dtype=np.uint8)
for k in range(16):
    codes[k::16] = ((sref.seq2[: (sref.n_bases + 15) // 16] >> (2 * k)) & 3)[: len(codes[k::16])]
asc = np.frombuffer(b"ACGT", dtype=np.uint8)[codes].copy()
asc[np.unpackbits(sref.inv.view(np.uint8), bitorder="little")[: sref.n_bases].astype(bool)] = ord("N")
with tempfile.TemporaryDirectory() as td:
    fa, bam = os.path.join(td, "r.fa"), os.path.join(td, "r.bam")
    write_fasta(fa, [("chr1", asc.tobytes())])
    t0 = time.perf_counter()
    write_bam(bam, [("chr1", sref.n_bases)], batch_to_records(sb, sref), level=6)
    out["bam"] = {"reads": args.bam_reads, "bam_bytes": os.path.getsize(bam), "python_writer_s": time.perf_counter() - t0}
    t0 = time.perf_counter(); ctx.load_fasta(fa); out["bam"]["load_fasta_s"] = time.perf_counter() - t0
    for _ in range(2):
        t0 = time.perf_counter(); ctx.profile_bam(bam, 51); tp = time.perf_counter() - t0
        t0 = time.perf_counter()
        with ctx.pileup_bam(bam) as h:
            h.counters
        tq = time.perf_counter() - t0
    out["bam"].update({"profile_bam_reads_per_s": args.bam_reads / tp, "pileup_bam_reads_per_s": args.bam_reads / tq,
                       "host_threads": os.cpu_count()})
    # the whole tools: loop + output files
    for _ in range(2):
        t0 = time.perf_counter(); ctx.error_bam(bam, 51); te = time.perf_counter() - t0
        t0 = time.perf_counter(); ctr = ctx.clust_bam(bam, os.path.join(td, "clusters.tsv"), None, 1); tc = time.perf_counter() - t0
    out["bam"].update({"error_tool_reads_per_s": args.bam_reads / te, "clust_tool_reads_per_s": args.bam_reads / tc,
                       "clust_tool_clusters": int(ctr["n_clusters"])})
    if args.bam_repeat > 1:
        # a file large enough for the batcher's steady state: the same records repeated (the profile does not mind the
        # order; the header still says coordinate-sorted)
        import gzip, struct
        from parasuite_b200.bamio import _bgzf_block
        data = gzip.open(bam, "rb").read()
        l_text, = struct.unpack_from("<I", data, 4)
        o = 8 + l_text
        n_ref, = struct.unpack_from("<I", data, o)
        o += 4
        for _ in range(n_ref):
            ln, = struct.unpack_from("<I", data, o)
            o += 8 + ln
        big = os.path.join(td, "big.bam")
        with open(big, "wb") as f:
            stream_bytes = data[:o] + data[o:] * args.bam_repeat
            for k in range(0, len(stream_bytes), 0xFF00):
                f.write(_bgzf_block(stream_bytes[k:k + 0xFF00], 6))
            f.write(_bgzf_block(b""))
        n_big = args.bam_reads * args.bam_repeat
        for _ in range(2):
            t0 = time.perf_counter(); ctx.profile_bam(big, 51); tb = time.perf_counter() - t0
        out["bam"]["big_file"] = {"reads": n_big, "bam_bytes": os.path.getsize(big), "profile_bam_reads_per_s": n_big / tb,
                                  "note": "batcher-bound: BGZF inflate + record location + packing on the host threads"}
print(json.dumps(out, indent=1))
