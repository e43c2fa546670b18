import sys, os, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.join(R, "para-suite_b200")); sys.path.insert(0, os.path.join(R, "oracle"))
import numpy as np, torch
from parasuite_b200 import synth
from parasuite_b200.runtime import Context, DeviceBatch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
ref = synth.synth_reference(0x5EED0001, [100_000_000])
b = synth.synth_reads(ref, n, 150, seed=0x5EED0002, mode=1)
ctx = Context(0); ctx.upload_reference(ref)
d = DeviceBatch(b, "cuda:0")
by = b.algorithmic_bytes(with_qual=True)
for it in range(3):
    ctx.profile_begin(176)
    ctx.kernel_times_reset(True)
    ctx.profile_batch_device(d, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ms = ctx.kernel_times_ms()
    print("kernel ms", ms, "GB/s", by / ms[0] / 1e6, "bytes/read", by / n)
    res = ctx.profile_end()
if "--check" in sys.argv:
    import oracle_lib
    oracle_lib.build()
    print("parity", np.array_equal(res["wide"], oracle_lib.profile_acc(ref, b, 176, threads=16)))
