#!/usr/bin/env python
"""BASELINE config 1 as a committed fixture: the reference's OWN example inputs -- examples/references/reference_chr1.fa,
reference_chr1_transcripts.fa and snp_db.vcf.gz, with the parameter files of examples/simulation -- taken through the two
tools the way bin/examples.sh:51,65 does (`error <bam> <fasta> 51`, `clust <bam> <fasta> <out> snp_db.vcf.gz 1`).

  python tests/golden/make_config1.py        (in the build container: reads /root/reference, writes tests/golden/config1/)

What cannot be reproduced here: the read simulator is a Perl script that needs Math::Random (absent) and is unseeded, the
aligners (BWA / Bowtie) and `comb` are absent too.  So reads are drawn by a SEEDED restatement of the simulator's model
(createSimulatedPARCLIPDataset.pl: clusters of reads on the transcripts, bound with p = 0.6, T>C conversion sites with the
rates of example.sitefrequency, substitution errors from example.errorprofile, qualities N(mu_j, sigma_j) from
example.qualities) and emitted DIRECTLY as the aligned records the pipeline would hand the two tools: genomic coordinates,
spliced reads as `aMbNcM` (what CombineGenomeTranscript writes), minus-strand transcripts with flag 16.  The shipped
snp_db.vcf.gz goes through the SNP filter as it is (bgzip); its T>C rows (1:17700, 1:29000) sit on C bases of the shipped
FASTA, so no T>C site can ever be at those positions and snpHit stays 0 on this fixture by construction of the reference's
own files (the filter's hits are covered by tests/test_flush_cpu.py and tests/test_gpu_stream.py).
Expected outputs come from the literal Python restatement of the Java loops (oracle/py_oracle.py):
  config1/reference_chr1.fa.gz   the FASTA as shipped (gzip; the test writes it back with its .fai)
  config1/snp_db.vcf.gz(.tbi)    the SNP file as shipped (bgzip + tabix index)
  config1/reads.json.gz          the records (flag, contig, pos, cigar, seq, qual)
  config1/expected.json.gz       profile arrays + the six clust output files + counters
"""
import gzip
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(REPO, "para-suite_b200"), os.path.join(REPO, "oracle"), os.path.join(REPO, "tests")):
    sys.path.insert(0, p)

import py_oracle as po  # noqa: E402
from helpers import to_py  # noqa: E402
from parasuite_b200 import Record  # noqa: E402

EX = "/root/reference/examples"
OUT = os.path.join(HERE, "config1")
READ_LEN = 36
MAX_LEN = 51
COMP = bytes.maketrans(b"ACGTacgt", b"TGCAtgca")


def read_fasta(path):
    out, name, buf = [], None, []
    for line in open(path, "rb").read().splitlines():
        if line.startswith(b">"):
            if name is not None:
                out.append((name, b"".join(buf)))
            name, buf = line[1:].decode(), []
        else:
            buf.append(line.strip())
    out.append((name, b"".join(buf)))
    return out


def main():
    rng = random.Random(0x5EED0C1)
    (chrom, genome), = [(n.split()[0], s) for n, s in read_fasta(f"{EX}/references/reference_chr1.fa")]
    err = [[float(x) for x in l.split()] for l in open(f"{EX}/simulation/example.errorprofile").read().splitlines() if l.strip()]
    site_rate = [float(x) for x in open(f"{EX}/simulation/example.sitefrequency").read().split()]
    quals = [[float(x) for x in l.split()] for l in open(f"{EX}/simulation/example.qualities").read().splitlines() if l.strip()]
    recs = []
    for header, _ in read_fasta(f"{EX}/references/reference_chr1_transcripts.fa"):
        gene, tx, _chr, starts, ends, strand = header.split("|")
        exons = sorted(zip([int(x) for x in starts.split(";")], [int(x) for x in ends.split(";")]))
        minus = strand.strip() == "-1"
        # transcript coordinate -> genomic position (1-based), exons in genomic order
        tpos = [p for a, b in exons for p in range(a, b + 1)]
        if len(tpos) < READ_LEN + 10 or max(tpos) > len(genome):
            continue
        n_clusters = max(2, len(tpos) // 60)
        for _ in range(n_clusters):
            c0 = rng.randrange(0, len(tpos) - READ_LEN)
            bound = rng.random() < 0.6
            # T>C sites of a bound cluster: T in transcript sense = T on the genome for +, A for - (read shows G)
            want = ord("A") if minus else ord("T")
            cand = [k for k in range(c0, min(c0 + READ_LEN, len(tpos))) if genome[tpos[k] - 1] in (want, want + 32)]
            sites = cand[:rng.randint(1, 4)] if bound else []
            for _ in range(max(1, int(rng.gauss(16, 10)))):
                s = min(max(0, c0 + rng.randint(-3, 3)), len(tpos) - READ_LEN)
                cols = tpos[s:s + READ_LEN]
                bases = bytearray(genome[p - 1] for p in cols).upper()
                for j, k in enumerate(range(s, s + READ_LEN)):
                    if k in sites and rng.random() < site_rate[min(sites.index(k), len(site_rate) - 1)]:
                        bases[j] = ord("G") if minus else ord("C")
                for j in range(READ_LEN):               # substitution errors, transcript sense
                    b = bases[j:j + 1].translate(COMP)[0] if minus else bases[j]
                    if b not in b"ACGT":
                        continue
                    row = err["ACGT".index(chr(b))]
                    x, acc = rng.random(), 0.0
                    for t, pr in enumerate(row):
                        acc += pr
                        if x < acc:
                            nb = ord("ACGT"[t])
                            break
                    bases[j] = bytes([nb]).translate(COMP)[0] if minus else nb
                q = bytes(min(64, max(3, int(rng.gauss(*quals[min(j, len(quals) - 1)])))) for j in range(READ_LEN))
                if minus:
                    q = q[::-1]                          # BAM stores minus-strand reads reversed with their qualities
                # cigar: runs of consecutive genomic positions, gaps as N
                ops, run = [], 1
                for a, b in zip(cols, cols[1:]):
                    if b == a + 1:
                        run += 1
                    else:
                        ops += [f"{run}M", f"{b - a - 1}N"]
                        run = 1
                ops.append(f"{run}M")
                recs.append(Record(16 if minus else 0, chrom, cols[0], "".join(ops), bytes(bases), q))
    recs.sort(key=lambda r: r.pos)
    genome_d = po.Genome({chrom: genome})
    keep = []
    for r in recs:                                       # records the JVM survives in both tools
        try:
            po.profile(to_py([r]), genome_d, MAX_LEN)
            po.pileup(to_py([r]), genome_d, po.SnpDb([]), 1)
            keep.append(r)
        except po.ReferenceWouldThrow:
            pass
    snps = []
    for line in gzip.open(f"{EX}/references/snp_db.vcf.gz", "rt"):
        if not line.startswith("#"):
            c = line.rstrip("\n").split("\t")
            snps.append((c[0], int(c[1]), c[3], c[4].split(",")[0]))
    prof = po.profile(to_py(keep), genome_d, MAX_LEN).wrapped()
    files = po.clust_files(to_py(keep), genome_d, po.SnpDb(snps), 1)
    st = po.pileup(to_py(keep), genome_d, po.SnpDb(snps), 1)
    os.makedirs(OUT, exist_ok=True)
    with gzip.GzipFile(os.path.join(OUT, "reference_chr1.fa.gz"), "wb", mtime=0) as f:
        f.write(open(f"{EX}/references/reference_chr1.fa", "rb").read())
    for name in ("snp_db.vcf.gz", "snp_db.vcf.gz.tbi"):     # the tabix index is what the reference's TabixReader opens
        with open(os.path.join(OUT, name), "wb") as f:
            f.write(open(f"{EX}/references/{name}", "rb").read())
    with gzip.GzipFile(os.path.join(OUT, "reads.json.gz"), "wb", mtime=0) as f:
        f.write(json.dumps([[r.flag, r.rname, r.pos, r.cigar, r.seq.decode(), list(r.qual)] for r in keep]).encode())
    exp = {"max_len": MAX_LEN, "min_cov": 1, "n_records": len(keep), "snps": snps,
           "profile": {k: (v if not hasattr(v, "tolist") else v.tolist()) for k, v in prof.items()}, "clust_files": files,
           "snp_hit": st.snp_hit, "double_stranded": st.double_stranded, "skipped_due_indel": st.skipped_due_indel,
           "n_clusters": len(st.clusters)}
    with gzip.GzipFile(os.path.join(OUT, "expected.json.gz"), "wb", mtime=0) as f:
        f.write(json.dumps(exp).encode())
    spliced = sum("N" in r.cigar for r in keep)
    print(f"{len(keep)} records ({spliced} spliced, {sum(r.flag & 16 > 0 for r in keep)} minus), {len(st.clusters)} clusters, "
          f"snp_hit {st.snp_hit}, T>C sites {sum(len(c.sites) for c in st.clusters)}")


if __name__ == "__main__":
    main()
