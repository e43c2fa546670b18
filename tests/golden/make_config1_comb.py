#!/usr/bin/env python
"""BASELINE config 1, the part of bin/examples.sh between the aligner and the two tools: `comb -g <genomic.bam>
-t <transcript.bam> -o <combined.bam>` (examples.sh:58), then `error <combined> reference_chr1.fa 51` (:51) and
`clust <combined> reference_chr1.fa <out> snp_db.vcf.gz 1` (:65).

  python tests/golden/make_config1_comb.py      (in the build container: reads /root/reference, writes tests/golden/config1/)

The aligners are absent, so their two outputs are drawn here, seeded, from the reference's own example files:
  transcript hits  reads of 36 nt on the sequences of examples/references/reference_chr1_transcripts.fa (the 16 transcripts,
                   both strands; reference names are the FASTA headers gene|transcript|chr|exonStarts|exonEnds|strand, which
                   is what CombineGenomeTranscript parses), clusters with T>C conversions, substitution errors from
                   example.errorprofile, qualities from example.qualities; some reads with a second (secondary) hit on
                   another transcript of the same gene (same genomic place: kept) or on another gene (dropped), some hits
                   with an insertion or deletion, some unplaced;
  genomic hits     reads drawn from reference_chr1.fa outside the transcripts.
Expected results come from the literal Python restatement of the Java (oracle/py_oracle.py): the combined records
(combine), then the profile arrays and the six clust files on them.  Transcript hits whose lifted record would kill the
JVM in one of the two tools (partial cigars of "indel + junction" hits) are left out of the inputs.
  config1/comb.json.gz     genomic hits, transcript hits (with read names), the expected combined records, the expected
                           profile and clust files
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
from make_config1 import EX, OUT, READ_LEN, MAX_LEN, read_fasta  # noqa: E402
from parasuite_b200 import Record  # noqa: E402

COMP = bytes.maketrans(b"ACGTacgt", b"TGCAtgca")


def rec_dict(name, flag, rname, pos, cigar, seq, qual, mapq=30):
    return {"name": name, "flag": flag, "rname": rname, "pos": pos, "cigar": cigar, "seq": seq, "qual": bytes(qual), "mapq": mapq}


def to_record(d):
    return Record(d["flag"], d["rname"], d["pos"], "" if d["cigar"] == "*" else d["cigar"], d["seq"], d["qual"])


def main():
    rng = random.Random(0xC0B1)
    (chrom, genome), = [(n.split()[0], s) for n, s in read_fasta(f"{EX}/references/reference_chr1.fa")]
    err = [[float(x) for x in l.split()] for l in open(f"{EX}/simulation/example.errorprofile").read().splitlines() if l.strip()]
    site_rate = [float(x) for x in open(f"{EX}/simulation/example.sitefrequency").read().split()]
    quals = [[float(x) for x in l.split()] for l in open(f"{EX}/simulation/example.qualities").read().splitlines() if l.strip()]
    transcripts = [(h, s.upper()) for h, s in read_fasta(f"{EX}/references/reference_chr1_transcripts.fa")]
    by_gene = {}
    for h, s in transcripts:
        by_gene.setdefault(h.split("|")[0], []).append((h, s))
    genome_d = po.Genome({chrom: genome})

    def mutate(seq, sites, s):
        b = bytearray(seq)
        for j in range(len(b)):
            if (s + j) in sites and b[j] == ord("T") and rng.random() < site_rate[min(sites.index(s + j), len(site_rate) - 1)]:
                b[j] = ord("C")
            if b[j] in b"ACGT":
                row = err["ACGT".index(chr(b[j]))]
                x, acc = rng.random(), 0.0
                for t, pr in enumerate(row):
                    acc += pr
                    if x < acc:
                        b[j] = ord("ACGT"[t])
                        break
        return bytes(b)

    t_recs, k = [], 0
    for h, tseq in transcripts:
        if len(tseq) < READ_LEN + 10:
            continue
        for _ in range(max(2, len(tseq) // 70)):
            c0 = rng.randrange(0, len(tseq) - READ_LEN)
            bound = rng.random() < 0.6
            cand = [p for p in range(c0, c0 + READ_LEN) if tseq[p:p + 1] == b"T"]
            sites = cand[:rng.randint(1, 4)] if bound else []
            for _ in range(max(1, int(rng.gauss(12, 8)))):
                s = min(max(0, c0 + rng.randint(-3, 3)), len(tseq) - READ_LEN - 2)
                L = READ_LEN
                cigar, ref_len = f"{L}M", L
                x = rng.random()
                if x < 0.04:
                    a = rng.randint(5, L - 8)
                    cigar, ref_len = f"{a}M1I{L - a - 1}M", L - 1
                elif x < 0.08:
                    a = rng.randint(5, L - 8)
                    cigar, ref_len = f"{a}M1D{L - a}M", L + 1
                seq = mutate(tseq[s:s + L], sites, s)
                q = bytes(min(64, max(3, int(rng.gauss(*quals[min(j, len(quals) - 1)])))) for j in range(L))
                name = "read%06d" % k
                k += 1
                hits = [rec_dict(name, 0, h, s + 1, cigar, seq, q)]
                y = rng.random()
                gene = h.split("|")[0]
                if y < 0.25 and len(by_gene[gene]) > 1:      # second hit on another transcript of the gene, same offset
                    h2, s2 = rng.choice([t for t in by_gene[gene] if t[0] != h])
                    if s + ref_len + 1 < len(s2):
                        hits.append(rec_dict(name, 0x100, h2, s + 1, cigar, seq, q))
                elif y < 0.30:                               # second hit somewhere else entirely
                    h2, s2 = rng.choice([t for t in transcripts if t[0].split("|")[0] != gene])
                    hits.append(rec_dict(name, 0x100, h2, rng.randint(1, max(1, len(s2) - 50)), cigar, seq, q))
                elif y < 0.33:                               # an unplaced record of the same read
                    hits.append(rec_dict(name, 4, "*", 0, "*", seq, q))
                t_recs.append(hits)
    g_recs = []
    covered = set()
    for h, _ in transcripts:
        f = h.split("|")
        for a, b in zip(f[3].split(";"), f[4].split(";")):
            covered.update(range(int(a) - 60, int(b) + 60))
    gk = 0
    while len(g_recs) < 400:
        c0 = rng.randrange(1000, len(genome) - 1000)
        if c0 in covered or b"N" in genome[c0:c0 + 60].upper():
            continue
        minus = rng.random() < 0.5
        for _ in range(rng.randint(3, 12)):
            s = c0 + rng.randint(-3, 3)
            seq = bytes(genome[s - 1:s - 1 + READ_LEN]).upper()
            q = bytes(min(64, max(3, int(rng.gauss(*quals[min(j, len(quals) - 1)])))) for j in range(READ_LEN))
            g_recs.append(rec_dict("gen%05d" % gk, 16 if minus else 0, chrom, s, f"{READ_LEN}M", seq, q))
            gk += 1
    g_recs.sort(key=lambda r: r["pos"])
    # keep the transcript reads whose lifted record survives both tools
    kept = []
    for hits in t_recs:
        out, _ = po.combine([chrom], "coordinate", [], hits)
        ok = True
        for d in out:
            try:
                po.profile(to_py([to_record(d)]), genome_d, MAX_LEN)
                po.pileup(to_py([to_record(d)]), genome_d, po.SnpDb([]), 1)
            except po.ReferenceWouldThrow:
                ok = False
        if ok:
            kept += hits
    combined, stats = po.combine([chrom], "coordinate", g_recs, kept)
    recs = [to_record(d) for d in combined]
    snps = []
    for line in gzip.open(f"{EX}/references/snp_db.vcf.gz", "rt"):
        if not line.startswith("#"):
            c = line.rstrip("\n").split("\t")
            snps.append((c[0], int(c[1]), c[3], c[4].split(",")[0]))
    prof = po.profile(to_py(recs), genome_d, MAX_LEN).wrapped()
    files = po.clust_files(to_py(recs), genome_d, po.SnpDb(snps), 1)

    def enc(d):
        return [d["name"], d["flag"], d["rname"], d["pos"], d["cigar"], d["seq"].decode(), list(d["qual"]), d["mapq"]]
    doc = {"genomic": [enc(d) for d in g_recs], "transcript": [enc(d) for d in kept],
           "transcripts": [[h, len(s)] for h, s in transcripts], "combined": [enc(d) for d in combined], "stats": stats,
           "profile": {k: (v if not hasattr(v, "tolist") else v.tolist()) for k, v in prof.items()}, "clust_files": files}
    with gzip.GzipFile(os.path.join(OUT, "comb.json.gz"), "wb", mtime=0) as f:
        f.write(json.dumps(doc).encode())
    print(f"{len(g_recs)} genomic, {len(kept)} transcript records ({sum(len(h) for h in t_recs) - len(kept)} left out), "
          f"{stats}, spliced cigars {sum('N' in d['cigar'] for d in combined)}, clusters {files['pileup'].count(chr(10)) - 1}")


if __name__ == "__main__":
    main()
