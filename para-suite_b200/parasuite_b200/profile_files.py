"""Host side of the `error` tool after the record loop: the six output files of ErrorProfiling.java:410-591 from the
arrays the profile kernels fill -- `.errorprofile` and `.indelprofile` are what the error-tolerant aligner consumes
(`bwa parasuite -p/-g`, PARAsuiteMapping.java:69-72).  Vectorised; doubles are printed the way Java prints them."""
from __future__ import annotations

from typing import Dict

import numpy as np

from .flush import java_double as jd

BASES = "ACGT"


def _div(a, b):
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.asarray(a, dtype=np.float64) / np.asarray(b, dtype=np.float64)


def profile_file_texts(res: Dict[str, np.ndarray], infer_qualities: bool = False) -> Dict[str, str]:
    """res = Context.profile*() result (Java int fields already wrapped).  Returns {suffix: text} plus the value of the
    'Averaged T2C = ... EPR' log line under 'averaged_t2c_epr'."""
    pc = np.asarray(res["position_conversions"], dtype=np.int64)               # [m][ref][read]
    m = pc.shape[0]
    n_proc = int(res["counters"][0])
    out = {}
    # per (ref, read) totals over positions; per-position totals wrap like the Java int field totalCountsPerPos
    tot = pc.sum(axis=0).astype(np.float64)
    tot_base = tot.sum(axis=1)
    tot_pos = pc.reshape(m, 16).sum(axis=1).astype(np.int64).astype(np.int32, casting="unsafe").astype(np.int64)
    frac = _div(tot, tot_base[:, None])
    qpct = _div(np.asarray(res["quality_per_mismatch"], dtype=np.int64), np.asarray(res["quality_per_mismatch_counts"], dtype=np.int64))
    out["errorprofile.vcf"] = "".join(
        "".join(f"{BASES[j]}\t{BASES[k]}\t{jd(float(tot[j, k]))}\n" for k in range(4)) + "\n" for j in range(4))
    out["errorprofile"] = "".join("".join(jd(float(frac[j, k])) + "\t" for k in range(4)) + "\n" for j in range(4))
    out["qualityPerMismatch"] = "".join("".join(jd(float(qpct[j, k])) + "\t" for k in range(4)) + "\n" for j in range(4))
    # averaged T>C errors per read over the leading positions with T>C > 0 (:532-542)
    t2c = _div(pc[:, 3, 1], n_proc) * 100.0
    stop = np.nonzero(~(t2c > 0))[0]
    if len(stop):
        j = int(stop[0])
        acc = 0.0
        for v in t2c[:j]:
            acc += float(v)
        avg = acc / (j + 1)
    else:
        avg = 0.0
        for v in t2c:
            avg += float(v)
    out["averaged_t2c_epr"] = jd(avg)
    # indel rates per position (:553-589)
    ins = np.asarray(res["insertions_per_pos"], dtype=np.float64)
    dele = np.asarray(res["deletions_per_pos"], dtype=np.float64)
    seen = tot_pos != 0
    ins_r = np.where(seen, _div(ins, np.where(seen, tot_pos, 1)), 0.0)
    del_r = np.where(seen, _div(dele, np.where(seen, tot_pos, 1)), 0.0)
    out["indels"] = "".join(f"{jd(float(a))}\t{jd(float(b))}\n" for a, b in zip(ins_r, del_r))

    def mean_nonzero(r):
        acc, zero = 0.0, 0
        for v in r:                   # sequential sum in position order, like the Java loop
            if v > 0:
                acc += float(v)
            else:
                zero += 1
        return acc, zero
    ia, iz = mean_nonzero(ins_r)
    da, dz = mean_nonzero(del_r)
    if iz == m and dz == m:
        ia = da = 0.0
    else:
        ia = float(_div(ia, m - iz))
        da = float(_div(da, m - dz))
    out["indelprofile"] = f"{jd(ia)}\t{jd(da)}"
    out["qualities"] = ""
    if infer_qualities:               # mean is exact; the SD is a different summation order than Java's linked list
        h = np.asarray(res["quality_hist"], dtype=np.float64)               # [m][256]; the byte is a signed Java byte
        vals = np.arange(256, dtype=np.float64)
        vals[128:] -= 256.0
        n = h.sum(axis=1)
        mean = _div((h * vals).sum(axis=1), n)
        var = _div((h * (vals[None, :] - mean[:, None]) ** 2).sum(axis=1), n)
        sd = np.sqrt(var)
        out["qualities"] = "".join(f"{jd(float(a))}\t{jd(float(b))}\n" for a, b in zip(mean, sd))
    return out


def write_profile_files(bam_path: str, res: Dict[str, np.ndarray], infer_qualities: bool = False) -> Dict[str, str]:
    """Write <bam>.errorprofile, .errorprofile.vcf, .qualityPerMismatch, .indels, .indelprofile, .qualities."""
    texts = profile_file_texts(res, infer_qualities)
    for suffix in ("errorprofile", "errorprofile.vcf", "qualityPerMismatch", "indels", "indelprofile", "qualities"):
        with open(f"{bam_path}.{suffix}", "w") as f:
            f.write(texts[suffix])
    return texts


def write_profile_files_native(bam_path: str, res: Dict[str, np.ndarray], infer_qualities: bool = False) -> float:
    """The same six files written by the library (csrc/profile_writer.cpp: ps_profile_write_files).  Returns the value of
    the 'Averaged T2C' log line."""
    import ctypes as C
    from . import abi
    lib = abi.load_library()
    keep = {
        "pc": np.ascontiguousarray(res["position_conversions"], dtype=np.int32).reshape(-1),
        "qpm": np.ascontiguousarray(res["quality_per_mismatch"], dtype=np.int32).reshape(-1),
        "qpmc": np.ascontiguousarray(res["quality_per_mismatch_counts"], dtype=np.int32).reshape(-1),
        "ins": np.ascontiguousarray(res["insertions_per_pos"], dtype=np.float64),
        "dele": np.ascontiguousarray(res["deletions_per_pos"], dtype=np.float64),
        "ctr": np.ascontiguousarray(res["counters"], dtype=np.int32),
    }
    r = abi.ps_profile_result()
    r.position_conversions = keep["pc"].ctypes.data
    r.quality_per_mismatch = keep["qpm"].ctypes.data
    r.quality_per_mismatch_counts = keep["qpmc"].ctypes.data
    r.insertions_per_pos = keep["ins"].ctypes.data
    r.deletions_per_pos = keep["dele"].ctypes.data
    r.counters = keep["ctr"].ctypes.data
    if infer_qualities:
        keep["qh"] = np.ascontiguousarray(res["quality_hist"], dtype=np.int64).reshape(-1)
        r.quality_hist = keep["qh"].ctypes.data
    m = keep["pc"].size // 16
    avg = C.c_double(0.0)
    err = C.create_string_buffer(512)
    st = lib.ps_profile_write_files(C.byref(r), m, int(infer_qualities), bam_path.encode(), C.byref(avg), err, len(err))
    if st != abi.PS_OK:
        raise abi.PsError(st, err.value.decode() or lib.ps_strerror(st).decode())
    return avg.value
