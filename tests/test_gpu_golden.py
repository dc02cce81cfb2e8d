"""GPU parity against the committed golden vectors (made from the unmodified reference by
tests/golden/make_golden.py) -- no oracle in the loop: CUDA path vs the reference's own outputs."""
import numpy as np
import pytest

from tests.test_oracle import _inputs, fnv
from tests.util import make_graph, read_str

pytestmark = pytest.mark.gpu


def _check(g, exp):
    tabs = g.get_raw()
    assert [t.size for t in tabs] == exp["nbytes"]
    assert [str(fnv(t)) for t in tabs] == exp["fnv"]


def test_kat_and_hash_vectors(gb, golden):
    k = golden["kat"]
    h = gb.CanLemireShifter(k["K"]).hash(k["seq"])
    assert (str(h.fw_hash), str(h.rc_hash)) == (k["fw"], k["rc"])
    seq = golden["hash_vectors"]["seq"]
    for name, exp in golden["hash_vectors"]["cases"].items():
        K = int(name[1:name.index("_")])
        can = int(name[-1])
        sh = [gb.FwdLemireShifter, gb.CanLemireShifter][can](K)
        fw, rc, _ = sh.hash_sequences(*gb._capi.reads_from_strings([seq]))
        assert fw.size == exp["n"]
        assert str(fnv(fw.view(np.uint8))) == exp["fw_fnv"], name
        if can:
            assert str(fnv(rc.view(np.uint8))) == exp["rc_fnv"], name


def test_reference_fixture_random20a(gb, golden):
    fx = golden["random20a"]
    bases, offsets = gb._capi.reads_from_strings(fx["reads"])
    for c in fx["cases"]:
        g = make_graph(gb, c["kind"], c["can"], c["K"], c["sizes"])
        assert g.insert_sequences(bases, offsets) == c["n_kmers"]
        _check(g, c["tables"])
        assert g.n_occupied() == c["n_occupied"]
        assert g.n_unique() == c["n_unique"]  # no in-batch slot collisions in this fixture
        if "query_first_read_after_3_passes" in c:
            g.insert_sequences(bases, offsets)
            g.insert_sequences(bases, offsets)
            assert g.query_sequence(fx["reads"][0]) == c["query_first_read_after_3_passes"]


@pytest.mark.parametrize("mode", [0, 1])
def test_synthetic_golden(gb, golden, mode):
    for c in golden["synthetic"]:
        bases, offsets = _inputs(c["input"])
        g = make_graph(gb, c["kind"], c["can"], c["K"], c["sizes"])
        tot = sum(g.insert_sequences(bases, offsets, mode=mode) for _ in range(c["passes"]))
        assert tot == c["n_kmers"], c["input"]
        _check(g, c["tables"])
        assert g.n_occupied() == c["n_occupied"]
        first = read_str(bases, offsets, 0)
        if len(first) >= c["K"]:
            assert g.query_sequence(first) == c["query_first_read"]


def test_insert_and_query_sequence_golden(gb, golden):
    for c in golden["insert_and_query_sequence"]:
        g = make_graph(gb, c["kind"], c["can"], c["K"], c["sizes"])
        assert g.insert_and_query_sequence(c["seq"]) == c["first"]
        assert g.insert_and_query_sequence(c["seq"]) == c["second"]
        assert g.n_unique() == c["n_unique"]


def test_median_golden(gb, golden):
    m = golden["median_count_at_least"]
    b, o = _inputs(m["insert"])
    qb, qo = _inputs(m["query"])
    for c in m["cases"]:
        g = make_graph(gb, c["kind"], c["can"], c["K"], c["sizes"])
        g.insert_sequences(b, o)
        for cutoff, exp in c["pass"].items():
            assert g.median_count_at_least(qb, qo, int(cutoff)).tolist() == exp, (c["kind"], cutoff)


def test_oxli_files_match_reference(gb, golden, tmp_path):
    b, o = _inputs(golden["oxli"]["input"])
    for c in golden["oxli"]["files"]:
        g = make_graph(gb, c["kind"], c["can"], c["K"], c["sizes"])
        g.insert_sequences(b, o)
        fn = str(tmp_path / ("t%d.oxli" % c["kind"]))
        g.save(fn)
        data = np.fromfile(fn, dtype=np.uint8)
        assert data.size == c["file_bytes"] and str(fnv(data)) == c["file_fnv"]
        # round trip through load()
        g2 = make_graph(gb, c["kind"], c["can"], c["K"], [11, 7, 5])
        g2.load(fn)
        assert g2.S.get_tablesizes() == c["sizes"]
        assert all(np.array_equal(x, y) for x, y in zip(g.get_raw(), g2.get_raw()))
        assert g2.n_occupied() == g.n_occupied()
