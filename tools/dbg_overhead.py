import sys, os, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.join(R, "para-suite_b200"))
import numpy as np, torch
from parasuite_b200 import synth
from parasuite_b200.runtime import Context, DeviceBatch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
ref = synth.synth_reference(0x5EED0001, [100_000_000])
b = synth.synth_reads(ref, n, 36, seed=0x5EED0002)
ctx = Context(0); ctx.upload_reference(ref)
d = DeviceBatch(b, "cuda:0")
st = torch.cuda.current_stream()
def prof():
    ctx.profile_begin(51); ctx.profile_batch_device(d, st.cuda_stream); return ctx.profile_end()
def pile():
    with ctx.pileup_run(d, first_running_id=1, stream=st.cuda_stream) as h: return h.counters
def both():
    ctx.profile_begin(51); ctx.profile_batch_device(d, st.cuda_stream)
    with ctx.pileup_run(d, first_running_id=1, stream=st.cuda_stream) as h: c = h.counters
    return ctx.profile_end()
for name, fn in (("profile", prof), ("pileup", pile), ("both", both)):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    ctx.kernel_times_reset(True)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    K = 50
    t0 = time.perf_counter(); e0.record(st)
    for _ in range(K): fn()
    e1.record(st); torch.cuda.synchronize(); t1 = time.perf_counter()
    kt = ctx.kernel_times_ms()
    print(f"{name}: {e0.elapsed_time(e1)/K:.4f} ms/step (events), {(t1-t0)*1e3/K:.4f} ms/step (host), timed kernel sections sum {float(kt.sum())/K:.4f} ms/step")
# host-side cost of the calls alone (tiny batch)
small = synth.synth_reads(ref, 1024, 36, seed=3)
ds = DeviceBatch(small, "cuda:0")
def both_small():
    ctx.profile_begin(51); ctx.profile_batch_device(ds, st.cuda_stream)
    with ctx.pileup_run(ds, first_running_id=1, stream=st.cuda_stream) as h: c = h.counters
    return ctx.profile_end()
for _ in range(5): both_small()
t0 = time.perf_counter()
for _ in range(200): both_small()
print("tiny batch, both tools:", (time.perf_counter() - t0) * 1e3 / 200, "ms/step host")
# host time of every call of the step (big batch: includes waiting for the device inside the two syncs)
import collections
acc = collections.defaultdict(float)
def timed(name, fn, *a, **k):
    t = time.perf_counter(); r = fn(*a, **k); acc[name] += time.perf_counter() - t; return r
K = 100
for which, batch in (("tiny", ds), ("10M", d)):
    acc.clear()
    for _ in range(K):
        timed("profile_begin", ctx.profile_begin, 51)
        timed("profile_batch_device", ctx.profile_batch_device, batch, st.cuda_stream)
        h = timed("pileup_run", ctx.pileup_run, batch, first_running_id=1, stream=st.cuda_stream)
        timed("counters", lambda: h.counters)
        timed("close", h.close)
        timed("profile_end", ctx.profile_end)
    print(which, {k: round(v * 1e6 / K, 1) for k, v in acc.items()}, "us per call")
