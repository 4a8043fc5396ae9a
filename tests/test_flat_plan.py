"""Work split of the tensor-core flat search (csrc/flat.cu: flat_plan, flat_seg_begin, flat_cta_of_tile -- the same
inline functions the kernel evaluates on the device), checked through svsb200_flat_plan.  No GPU needed."""
import ctypes as C

import numpy as np
import pytest

KC, CMAX, BM, BN = 33, 1024, 128, 256


def plan(nq, n, sms):
    from scalablevectorsearch_b200 import _lib
    lib = _lib.lib()
    ctas, share, lists = C.c_uint32(), C.c_uint32(), C.c_uint32()
    _lib.check(lib.svsb200_flat_plan(nq, n, sms, C.byref(ctas), C.byref(share), C.byref(lists), None))
    seg = np.zeros(ctas.value // share.value + 1, dtype=np.uint64)
    _lib.check(lib.svsb200_flat_plan(nq, n, sms, C.byref(ctas), C.byref(share), C.byref(lists), seg.ctypes.data))
    return ctas.value, share.value, lists.value, seg


@pytest.mark.parametrize("nq,n,sms", [
    (10_000, 1_000_000, 148), (100, 1_000_000, 148), (700, 1_000_000, 148), (1, 300, 148), (129, 257, 148),
    (1100, 40_000, 148), (600, 30_000, 148), (100_000, 12_500_000, 148), (5000, 999, 148), (10_000, 1_000_000, 132),
    (513, 100_000, 7), (4096, 65_536, 1),
])
def test_segments_partition_the_tiles_and_the_lists_fit(nq, n, sms):
    ctas, share, lists, seg = plan(nq, n, sms)
    mtiles, ntiles = -(-nq // BM), -(-n // BN)
    assert share in (1, 2, 4) and ctas % share == 0 and 1 <= ctas <= max(sms, share)
    assert share == (4 if mtiles >= 8 else 2 if mtiles >= 4 else 1)
    nseg, ngroups = ctas // share, -(-mtiles // share)
    total = ngroups * ntiles
    # contiguous, non-empty, near-equal runs that cover every (row group, base tile) pair exactly once
    assert seg[0] == 0 and seg[-1] == total and len(seg) == nseg + 1
    runs = np.diff(seg.astype(np.int64))
    assert runs.min() >= 1 and runs.max() - runs.min() <= 1
    # a row group is cut into at most lists / 2 pieces (two lists per piece), and they fit the rescoring kernel
    assert lists % 2 == 0 and lists * KC <= CMAX
    first = np.searchsorted(seg, np.arange(ngroups, dtype=np.uint64) * ntiles, side="right") - 1
    last = np.searchsorted(seg, (np.arange(ngroups, dtype=np.uint64) + 1) * ntiles - 1, side="right") - 1
    assert int((last - first + 1).max()) * 2 == lists
    # one CTA per SM at most, and as many as the candidate budget allows
    if nseg < max(1, sms // share) and nseg < total:
        _, _, lists_more, _ = plan_with(nq, n, ctas // share + 1, share)
        assert lists_more * KC > CMAX


def plan_with(nq, n, nseg, share):
    """The split the library would make with `nseg` segments (its own search starts from sm_count / share and goes down):
    restated here only to check that it stopped at the largest admissible count."""
    mtiles, ntiles = -(-nq // BM), -(-n // BN)
    ngroups = -(-mtiles // share)
    total = ngroups * ntiles
    seg = np.array([total * b // nseg for b in range(nseg + 1)], dtype=np.uint64)
    first = np.searchsorted(seg, np.arange(ngroups, dtype=np.uint64) * ntiles, side="right") - 1
    last = np.searchsorted(seg, (np.arange(ngroups, dtype=np.uint64) + 1) * ntiles - 1, side="right") - 1
    return nseg * share, share, int((last - first + 1).max()) * 2, seg


def test_flat_plan_rejects_bad_arguments():
    from scalablevectorsearch_b200 import _lib
    lib = _lib.lib()
    a = C.c_uint32()
    assert lib.svsb200_flat_plan(0, 10, 148, C.byref(a), C.byref(a), C.byref(a), None) != 0
    assert lib.svsb200_flat_plan(10, 10, 0, C.byref(a), C.byref(a), C.byref(a), None) != 0
    assert b"flat_plan" in lib.svsb200_last_error()
