import sys, os, time, tempfile
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.join(R, "para-suite_b200"))
import numpy as np
from parasuite_b200 import synth
from parasuite_b200.runtime import Context
from parasuite_b200.bamio import batch_to_records, write_bam, write_fasta
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
sref = synth.synth_reference(7, [5_000_000], n_run=1000)
sb = synth.synth_reads(sref, n, 36, seed=8)
codes = np.zeros(sref.n_bases, dtype=np.uint8)
for k in range(16):
    codes[k::16] = ((sref.seq2[: (sref.n_bases + 15) // 16] >> (2 * k)) & 3)[: len(codes[k::16])]
asc = np.frombuffer(b"ACGT", dtype=np.uint8)[codes].copy()
asc[np.unpackbits(sref.inv.view(np.uint8), bitorder="little")[: sref.n_bases].astype(bool)] = ord("N")
ctx = Context(0)
with tempfile.TemporaryDirectory() as td:
    fa, bam = os.path.join(td, "r.fa"), os.path.join(td, "r.bam")
    write_fasta(fa, [("chr1", asc.tobytes())])
    write_bam(bam, [("chr1", sref.n_bases)], batch_to_records(sb, sref), level=6)
    ctx.load_fasta(fa)
    for it in range(4):
        t0 = time.perf_counter(); ctx.profile_bam(bam, 51); tp = time.perf_counter() - t0
        t0 = time.perf_counter()
        with ctx.pileup_bam(bam) as h:
            h.counters
        tq = time.perf_counter() - t0
        print(f"iter {it}: profile_bam {tp*1e3:.1f} ms, pileup_bam {tq*1e3:.1f} ms", flush=True)
