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
ctx.lib.ps_debug_word.restype = C.c_uint64
ctx.lib.ps_debug_word.argtypes = [C.c_void_p, C.c_int]
for it in range(8):
    ctx.profile_begin(51)
    ctx.kernel_times_reset(True)
    ctx.profile_batch_device(d, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    print("kernel ms", ctx.kernel_times_ms(), "fast reads", ctx.lib.ps_debug_word(ctx.h, 1), "of", n)
    ctx.profile_end()
