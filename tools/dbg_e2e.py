import sys, os, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.join(R, "para-suite_b200"))
import numpy as np, torch
from parasuite_b200 import synth
from parasuite_b200.runtime import Context, PinnedBatch
ref = synth.synth_reference(0x5EED0001, [100_000_000])
b = synth.synth_reads(ref, 10_000_000, 36, seed=0x5EED0002)
ctx = Context(0); ctx.upload_reference(ref)
p = PinnedBatch(b)
def t(f, n=5):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
def up():
    v = ctx.upload(p); torch.cuda.synchronize(); return v
print("upload ms", t(up), "GB/s", p.h2d_bytes / t(up) / 1e6)
v = up()
def prof():
    ctx.profile_begin(51); ctx.profile_batch_device(v); return ctx.profile_end()
print("profile ms", t(prof))
def pile():
    with ctx.pileup_run(v) as h: return h.counters
print("pileup run ms", t(pile))
def pilef():
    with ctx.pileup_run(v) as h: return h.fetch(pinned=True, boundary=False)
print("pileup run+fetch ms", t(pilef))
def allf():
    v = ctx.upload(p)
    with ctx.pileup_run(v) as h: f = h.fetch(pinned=True, boundary=False)
    ctx.profile_begin(51); ctx.profile_batch_device(v); r = ctx.profile_end()
    return f
print("all ms", t(allf))
