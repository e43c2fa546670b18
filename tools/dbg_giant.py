import sys, os, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.join(R, "para-suite_b200")); sys.path.insert(0, os.path.join(R, "oracle"))
import numpy as np, torch
from parasuite_b200 import synth
from parasuite_b200.runtime import Context, DeviceBatch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
span = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000
ref = synth.synth_reference(77, [span], n_run=0)
b = synth.synth_reads(ref, n, 36, seed=5, n_ppm=0 if "--sorted" in sys.argv else 1000)
if "--sorted" in sys.argv:   # the generator sorts inside a cluster only; a BAM is sorted by start throughout
    from parasuite_b200.sharding import take_uniform
    b = take_uniform(b, np.argsort(b.ref_start, kind="stable"))
ctx = Context(0); ctx.upload_reference(ref)
d = DeviceBatch(b, "cuda:0")
for it in range(3):
    t0 = time.perf_counter()
    with ctx.pileup_run(d) as h:
        c = h.counters
    dt = time.perf_counter() - t0
    print("pileup wall ms", dt * 1e3, "stage ms", ctx.pileup_stage_ms(), c["n_clusters"], c["n_sites"])
