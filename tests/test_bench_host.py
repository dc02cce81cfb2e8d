"""CPU tests of bench.py's host logic (no GPU): which reference-pinned golden entry a run is checked against."""
import importlib.util
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_golden_entry_selection():
    b = _bench()
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "fullsize_golden.json")))
    for name in ("c1", "c2", "c3"):
        total = b.WORKLOADS[name][4]
        e = b.golden_entry(name, total, total)
        assert e is not None and e["reads"] == total and e["checksums"] == gold[name]["checksums"]
    # a pinned prefix of the read set is found by its read count; any other count has no golden entry
    total5 = b.WORKLOADS["c5"][4]
    e = b.golden_entry("c5", total5, total5)  # the full C5 set: 4.7 hours of the compiled reference on 8 threads
    assert e is not None and e["reads"] == total5 and e["checksums"] == gold["c5"]["checksums"]
    e = b.golden_entry("c5", 125_000, total5)
    assert e is not None and e["checksums"] == gold["c5_first_125k"]["checksums"]
    assert b.golden_entry("c3", 1_000_000, b.WORKLOADS["c3"][4])["checksums"] == gold["c3_first_1m"]["checksums"]
    assert b.golden_entry("c3", 123_456, b.WORKLOADS["c3"][4]) is None


def test_checksum_check_block():
    b = _bench()
    gold = {"checksums": [1, 2, 3], "n_occupied": 7, "threads": 8}
    assert b.checksum_check(gold, [1, 2, 3], 7)["tables_checksum_equal_reference"] is True
    assert b.checksum_check(gold, [1, 2, 4], 7)["tables_checksum_equal_reference"] is False
    assert b.checksum_check(gold, [1, 2, 3], 8)["tables_checksum_equal_reference"] is False
    assert b.checksum_check(None, [1], 1)["tables_checksum_equal_reference"] is None
