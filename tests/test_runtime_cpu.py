"""Host-side binding logic that needs no GPU: the single-allocation result layout of Context.profile_end."""
import numpy as np

from parasuite_b200 import abi
from parasuite_b200.runtime import Context


class _FakeLib:
    @staticmethod
    def ps_profile_acc_len(m, q):
        return 16 * m + 32 + 2 * m + abi.PS_PC_COUNT + (256 * m if q else 0)


def _ctx(max_len, infer_q):
    c = Context.__new__(Context)            # no device: only the array bookkeeping is exercised
    c.lib, c._max_len, c._infer_q, c._result_plans = _FakeLib, max_len, infer_q, {}
    return c


def test_profile_result_arrays_layout():
    for max_len, infer_q in ((51, False), (101, True), (3000, False)):
        c = _ctx(max_len, infer_q)
        out, r = Context._profile_result_arrays(c)
        assert out["position_conversions"].shape == (max_len, 4, 4) and out["position_conversions"].dtype == np.int32
        assert out["quality_per_mismatch"].shape == (4, 4) and out["quality_per_mismatch_counts"].shape == (4, 4)
        assert out["insertions_per_pos"].dtype == np.float64 and out["deletions_per_pos"].shape == (max_len,)
        assert out["counters"].shape == (abi.PS_PC_COUNT,)
        assert out["wide"].dtype == np.int64 and out["wide"].size == _FakeLib.ps_profile_acc_len(max_len, infer_q)
        assert ("quality_hist" in out) == infer_q
        if infer_q:
            assert out["quality_hist"].shape == (max_len, 256)
        else:
            assert not r.quality_hist
        # every array is a writable, 8-byte aligned, non-overlapping view; the struct points at exactly these views
        spans = []
        for name, a in out.items():
            p = a.__array_interface__["data"][0]
            assert p % 8 == 0 and a.flags.writeable and not a.any()
            assert getattr(r, name) == p
            spans.append((p, p + a.nbytes))
        spans.sort()
        assert all(e0 <= s1 for (_, e0), (s1, _) in zip(spans, spans[1:]))
        out["wide"][3] = 7
        assert out["wide"][3] == 7 and not out["counters"].any()
        # a second run gets fresh memory (results of the previous run stay valid), the plan is cached
        out2, _ = Context._profile_result_arrays(c)
        assert out2["wide"][3] == 0 and len(c._result_plans) == 1
