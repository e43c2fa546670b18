import sys, os, time, ctypes as C
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.join(R, "para-suite_b200")); sys.path.insert(0, os.path.join(R, "oracle"))
import numpy as np, torch
from parasuite_b200 import synth
from parasuite_b200.runtime import Context, DeviceBatch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 36
ref = synth.synth_reference(0x5EED0001, [100_000_000])
b = synth.synth_reads(ref, n, L, seed=0x5EED0002)
ctx = Context(0); ctx.upload_reference(ref)
d = DeviceBatch(b, "cuda:0")
best = None
for it in range(8):
    ctx.kernel_times_reset(True)
    t0 = time.perf_counter()
    res = ctx.pileup(d, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    st = ctx.pileup_stage_ms()
    best = st if best is None else np.minimum(best, st)
    print("device ms", ctx.kernel_times_ms(), "stages (flag, cluster, compact)", st, "wall ms", dt * 1e3, "clusters", len(res["clusters"]), "sites", len(res["sites"]))
print("BEST stages (flag, cluster, compact) ms", best, "sum", float(best.sum()))
