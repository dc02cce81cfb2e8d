"""CPU test: the C-ABI library builds, loads, and exports every symbol include/goetia_b200.h declares.
No compute call is made (there is no GPU here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "goetia_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gt_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_and_exports_header():
    from goetia_b200 import build, _capi
    build.build()
    L = _capi.load()
    names = _declared()
    assert len(names) >= 40
    for n in names:
        assert hasattr(L, n), "libgoetia_b200.so does not export " + n
    # and the binding table covers exactly the header
    assert sorted(_capi.SIGNATURES) == names
    assert L.gt_abi_version() == 1


def test_host_only_entry_points():
    from goetia_b200 import get_n_primes_near_x, _capi
    assert get_n_primes_near_x(4, 10**6) == [999983, 999979, 999961, 999959]
    assert get_n_primes_near_x(1, 1) == [1]
    assert _capi.load().gt_max_hash_from_scaled(1000) == 18446744073709552
    assert _capi.load().gt_max_hash_from_scaled(0) == 0
    assert _capi.load().gt_max_hash_from_scaled(1) == 2**64 - 1


def test_compute_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import goetia_b200 as gb
    with pytest.raises(gb.GoetiaB200Error):
        gb.BitStorage(1000, 4)


def test_product_never_imports_oracle():
    """The product package must not reference the checker."""
    pkg = os.path.join(ROOT, "goetia_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inc", ".h", ".hh")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), os.path.join(dp, f)


def test_host_cursor_shifter_matches_golden(golden):
    """hash_base / shift_right / shift_left cursor members (host latency path)."""
    from goetia_b200.hashing import CanLemireShifter, FwdLemireShifter
    k = golden["kat"]
    s = CanLemireShifter(k["K"])
    h = s.hash_base(k["seq"])
    assert (str(h.fw_hash), str(h.rc_hash)) == (k["fw"], k["rc"])
    seq = golden["hash_vectors"]["seq"]
    for K in (21, 31, 64, 101):
        exp = golden["hash_vectors"]["cases"]["K%d_can1" % K]
        sh = CanLemireShifter(K)
        h = sh.hash_base(seq)
        got = [h]
        for i in range(1, 3):
            got.append(sh.shift_right(seq[i - 1], seq[i + K - 1]))
        assert [str(g.fw_hash) for g in got] == exp["fw_head"]
        assert [str(g.rc_hash) for g in got] == exp["rc_head"]
        # shift_left undoes shift_right (reference tests/test_hashing.py:303-329)
        back = sh.shift_left(seq[1], seq[2 + K - 1])
        assert (str(back.fw_hash), str(back.rc_hash)) == (exp["fw_head"][1], exp["rc_head"][1])
        f = FwdLemireShifter(K)
        assert str(f.hash_base(seq).value()) == golden["hash_vectors"]["cases"]["K%d_can0" % K]["fw_head"][0]


def test_bin_of_arithmetic_is_exact():
    """The cheap reduction k_bucket uses for big tables equals h % d for every divisor class and edge value."""
    import random
    from tests.csrc_check import bin_of_host, fastmod_host
    rnd = random.Random(7)
    ds = [2**28, 2**28 + 1, 2**29 - 3, 999999937, 2**30 + 7, 2**31 - 1, 2**31, 3999999979, 2**32 - 1, 2**32, 2**32 + 1,
          7999999957, 2**33 - 9, 2**40 + 15, 2**58, 2**28 - 1, 999983, 2**58 + 1]
    ds += [rnd.randrange(2**28, 2**36) for _ in range(40)]
    for d in ds:
        hs = [0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, 2**32 - 1, 2**32, 2**63, 2**64 - 1, 2**64 - 2, (2**64 // d) * d - 1,
              (2**64 // d) * d % 2**64]
        hs += [rnd.randrange(2**64) for _ in range(400)]
        hs += [(k * d + rnd.randrange(-2, 3)) % 2**64 for k in (1, 2, 3, 2**20, 2**31) for _ in range(3)]
        for h in hs:
            assert fastmod_host(h, d) == h % d
            got = bin_of_host(h, d)
            if d < 2**28 or d > 2**58:
                assert got is None
            else:
                assert got == h % d, (h, d)
