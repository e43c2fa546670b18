"""Shared test helpers: KAT -> records, random record generators, oracle <-> oracle comparison."""
import random

import numpy as np

import py_oracle as po
from parasuite_b200 import PackedReference, ReadBatch, Record

BASES = "ACGT"


def kat_records(reads, with_qual=True):
    out = []
    for t in reads:
        if len(t) == 5:
            f, p, c, s, ql = t
        else:
            f, p, c, s = t
            ql = bytes([30] * len(s))
        out.append(Record(f, "chr1", p, c, s.encode(), ql))
    return out


def to_py(records):
    """parasuite_b200.Record -> py_oracle.Rec (what htsjdk would hand the Java loop)."""
    out = []
    for r in records:
        q = bytes(r.qual)
        if len(q) == 0 or q[0] == 0xFF:
            q = b""
        out.append(po.Rec(r.flag, r.rname, r.pos, po.parse_cigar(r.cigar), bytes(r.seq), q))
    return out


def py_profile_dict(st: po.ProfileState):
    w = st.wrapped()
    return {
        "position_conversions": np.asarray(w["pos_conv"], dtype=np.int32),
        "quality_per_mismatch": np.asarray(w["qual_mm"], dtype=np.int32),
        "quality_per_mismatch_counts": np.asarray(w["qual_mm_cnt"], dtype=np.int32),
        "insertions_per_pos": np.asarray(w["ins_per_pos"], dtype=np.float64),
        "deletions_per_pos": np.asarray(w["del_per_pos"], dtype=np.float64),
        "counters": np.asarray(w["counters"], dtype=np.int32),
    }


def assert_profile_equal(a: dict, b: dict, what=""):
    for k in ("position_conversions", "quality_per_mismatch", "quality_per_mismatch_counts", "insertions_per_pos",
              "deletions_per_pos", "counters"):
        x, y = np.asarray(a[k]), np.asarray(b[k])
        assert x.shape == y.shape, (what, k, x.shape, y.shape)
        if not np.array_equal(x, y):
            idx = np.argwhere(x != y)[:5]
            raise AssertionError(f"{what}: {k} differs at {idx.tolist()}: {x[tuple(idx[0])]} vs {y[tuple(idx[0])]}")


def random_genome(rng: random.Random, n_contigs=2, length=400, n_frac=0.03, lower_frac=0.2):
    contigs = []
    for c in range(n_contigs):
        s = bytearray(rng.choice(b"ACGT") for _ in range(length))
        # N run + lower-case run + a stray IUPAC code
        a = rng.randrange(0, length - int(length * n_frac) - 1)
        for k in range(a, a + int(length * n_frac)):
            s[k] = ord("N")
        b = rng.randrange(0, length - int(length * lower_frac) - 1)
        for k in range(b, b + int(length * lower_frac)):
            s[k] = ord(chr(s[k]).lower())
        s[rng.randrange(length)] = ord("R")
        contigs.append((f"chr{c + 1}", bytes(s)))
    return contigs


def random_cigar(rng: random.Random, L: int, kind: str):
    """Return cigar text consuming exactly L read bases. kind: 'M', 'clip', 'indel', 'splice', 'wild'."""
    if kind == "M":
        return f"{L}M"
    ops = []
    left = L
    if kind in ("clip", "wild") and rng.random() < 0.6 and left > 6:
        n = rng.randint(1, 3)
        ops.append((n, "S")); left -= n
    tail = None
    if kind in ("clip", "wild") and rng.random() < 0.6 and left > 6:
        n = rng.randint(1, 3)
        tail = (n, "S"); left -= n
    n_ev = 0 if kind == "clip" else rng.randint(1, 2)
    for _ in range(n_ev):
        if left < 6:
            break
        m = rng.randint(2, left - 3)
        ops.append((m, rng.choice("M=X") if kind == "wild" else "M")); left -= m
        if kind == "indel":
            ev = rng.choice("ID")
        elif kind == "splice":
            ev = "N"
        else:
            ev = rng.choice("IDNP")
        n = rng.randint(1, 3)
        if ev == "I":
            n = min(n, left - 1)
            if n <= 0:
                continue
            left -= n
        ops.append((n, ev))
    if left > 0:
        ops.append((left, "M"))
    if tail:
        ops.append(tail)
    if kind == "wild" and rng.random() < 0.3:
        ops.insert(0, (2, "H"))
    # merge nothing; htsjdk keeps elements as given
    return "".join(f"{n}{op}" for n, op in ops)


def ref_len(cigar_text):
    return sum(n for op, n in po.parse_cigar(cigar_text) if op in "MDN=X")


def random_records(rng: random.Random, contigs, n, kinds=("M",), Lrange=(20, 40), err=0.05, sorted_=True,
                   flags_special=0.0):
    """Records that do not crash the JVM are not guaranteed; callers filter with the Python oracle."""
    recs = []
    comp = {65: 84, 67: 71, 71: 67, 84: 65}
    for _ in range(n):
        ci = rng.randrange(len(contigs))
        name, seq = contigs[ci]
        L = rng.randint(*Lrange)
        cg = random_cigar(rng, L, rng.choice(kinds))
        R = ref_len(cg)
        if R + 2 >= len(seq):
            continue
        pos = rng.randint(1, len(seq) - R)
        flag = 16 if rng.random() < 0.5 else 0
        if rng.random() < flags_special:
            flag |= rng.choice([0x4, 0x400, 0x100, 0x200, 0x800])
        # read bases: follow the reference through the cigar, with errors / T>C / N
        out = bytearray()
        rp = pos - 1
        for op, ln in po.parse_cigar(cg):
            if op in "M=X":
                for k in range(ln):
                    b = seq[rp + k] if rp + k < len(seq) else ord("A")
                    b = ord(chr(b).upper())
                    if b not in b"ACGT":
                        b = rng.choice(b"ACGT")
                    out.append(b)
                rp += ln
            elif op in "IS":
                out += bytes(rng.choice(b"ACGT") for _ in range(ln))
            elif op in "DN":
                rp += ln
        for k in range(len(out)):
            x = rng.random()
            if x < err:
                out[k] = rng.choice(b"ACGT")
            elif x < err + 0.01:
                out[k] = ord("N")
            elif out[k] == ord("T") and x < err + 0.15 and not (flag & 16):
                out[k] = ord("C")
            elif out[k] == ord("A") and x < err + 0.15 and (flag & 16):
                out[k] = ord("G")
        qual = bytes(rng.randint(3, 41) for _ in range(len(out)))
        p = pos
        if rng.random() < flags_special * 0.3:
            p = 0
        recs.append((ci, Record(flag, name, p, cg, bytes(out), qual)))
    if sorted_:
        recs.sort(key=lambda t: (t[0], t[1].pos))
    return [r for _, r in recs]
