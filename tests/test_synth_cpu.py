"""CPU: the synthetic workload generator is deterministic, sorted, and consistent with the SoA contract."""
import numpy as np
import pytest

from parasuite_b200 import abi


@pytest.fixture(scope="module")
def synth():
    import os
    if not os.path.exists(os.path.join(abi.LIB_DIR, "libps_synth.so")):
        pytest.skip("libps_synth.so not built")
    from parasuite_b200 import synth as s
    return s


def test_synth_sorted_deterministic(synth, oracle):
    ref = synth.synth_reference(1, [300_000, 200_000], n_run=500)
    a = synth.synth_reads(ref, 50_000, 36, seed=7, special_ppm=2000, threads=3)
    b = synth.synth_reads(ref, 50_000, 36, seed=7, special_ppm=2000, threads=1)
    for f in ("meta", "ref_start", "bases2", "qual", "cigar", "exc"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    assert np.all(np.diff(a.ref_start.astype(np.int64)) >= 0)
    assert a.algorithmic_bytes() == 66 * a.n_reads
    r = oracle.profile(ref, a, 51, threads=2)
    c = r["counters"]
    assert c[1] > 0 and c[2] > 0 and c[3] > 0                      # unmapped / duplicate / startZero present
    assert c[0] + c[1] + c[2] + c[3] == a.n_reads
    tot = r["position_conversions"].sum(axis=0)
    assert tot[3, 1] > 3 * tot[0, 1]                               # T>C enriched over A>C
    assert r["position_conversions"].sum() == c[7]


def test_synth_dense_cigar(synth, oracle):
    ref = synth.synth_reference(2, [400_000], n_run=0)
    a = synth.synth_reads(ref, 20_000, 150, seed=9, mode=1)
    assert a.uniform_ncigar == 0 and a.uniform_len == 150
    r = oracle.profile(ref, a, 176, threads=2)                     # must not fault by construction
    c = r["counters"]
    assert c[4] > 0 and c[5] > 0 and c[6] > 0                      # indel reads, skipped (Q5), longer indels
    assert r["insertions_per_pos"].sum() > 0 and r["deletions_per_pos"].sum() > 0


def test_synth_dense_is_sorted_throughout(synth):
    """Overlapping neighbouring clusters interleave: the batch is sorted by start like a BAM, not only inside clusters."""
    ref = synth.synth_reference(3, [30_000], n_run=0)
    a = synth.synth_reads(ref, 40_000, 36, seed=11, threads=2)
    assert np.all(np.diff(a.ref_start.astype(np.int64)) >= 0)
    b = synth.synth_reads(ref, 40_000, 150, seed=11, mode=1, threads=2)
    assert np.all(np.diff(b.ref_start.astype(np.int64)) >= 0)
