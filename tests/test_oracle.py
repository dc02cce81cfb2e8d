"""CPU tests: pin the plain-C oracle (oracle/goetia_oracle.c) against golden vectors produced by
the unmodified reference (tests/golden/golden.json, made by tests/golden/make_golden.py), and --
where oracle/_ref was built -- against the reference library directly."""
import os
import tempfile

import numpy as np
import pytest

from oracle.binding import Port, PortSketch, Ref, have_ref, synth_reads
from tests.util import genome_reads, ragged_reads, read_str

fnv = Port.fnv1a


def _inputs(spec):
    gen = spec["gen"]
    if gen == "synth_reads":
        return synth_reads(spec["n_reads"], spec["length"], spec["seed"])
    if gen == "genome_reads":
        return genome_reads(spec["n_reads"], spec["read_len"], spec["genome_len"], spec["seed"],
                            spec.get("sub_rate", 0.01))
    if gen == "ragged_reads":
        return ragged_reads(spec["n_reads"], spec["min_len"], spec["max_len"], spec["seed"])
    if gen == "tiled":
        one, _ = synth_reads(1, spec["length"], spec["seed"])
        reps = spec["reps"]
        return np.tile(one, reps), np.arange(reps + 1, dtype=np.uint64) * np.uint64(spec["length"])
    raise ValueError(gen)


def _check_tables(g, exp):
    tabs = g.tables()
    assert [t.size for t in tabs] == exp["nbytes"]
    assert [str(fnv(t)) for t in tabs] == exp["fnv"]
    assert [int(t.astype(np.uint64).sum()) for t in tabs] == exp["bytesum"]


def test_known_answer_vector(golden):
    k = golden["kat"]
    fw, rc = Port.hash_sequence(1, k["K"], k["seq"])
    assert (str(int(fw[0])), str(int(rc[0]))) == (k["fw"], k["rc"])
    assert str(int(Port.hash_sequence(0, k["K"], k["seq"])[0][0])) == k["fwd_only"]
    # literal from the reference's tests/test_hashing.py:15-19
    assert int(fw[0]) == 13194817695400542713 and int(rc[0]) == 4324216031038051805


def test_char_table_entries(golden):
    # single-base "k-mers" expose the table entries: K=1 hash of c is T[c]
    for c in "ACGT":
        assert str(int(Port.hash_sequence(0, 1, c)[0][0])) == golden["char_table"][c]


def test_hash_vectors(golden):
    seq = golden["hash_vectors"]["seq"]
    for name, exp in golden["hash_vectors"]["cases"].items():
        K = int(name[1:name.index("_")])
        can = int(name[-1])
        fw, rc = Port.hash_sequence(can, K, seq)
        assert fw.size == exp["n"]
        assert str(fnv(fw.view(np.uint8))) == exp["fw_fnv"], name
        if can:
            assert str(fnv(rc.view(np.uint8))) == exp["rc_fnv"], name


def test_rolling_equals_from_scratch():
    """reference tests/test_hashing.py:161-172, 221-231: rolled hash == hash of each k-mer alone,
    canonical == min(fwd(kmer), fwd(revcomp(kmer)))."""
    comp = str.maketrans("ACGT", "TGCA")
    seq = read_str(*synth_reads(1, 300, seed=8), 0)
    for K in (21, 27, 31, 64, 101):
        fw, rc = Port.hash_sequence(1, K, seq)
        for i in range(0, len(seq) - K + 1, 7):
            kmer = seq[i:i + K]
            f1, r1 = Port.hash_sequence(1, K, kmer)
            assert (fw[i], rc[i]) == (f1[0], r1[0])
            assert rc[i] == Port.hash_sequence(0, K, kmer.translate(comp)[::-1])[0][0]


def test_primes(golden):
    for key, exp in golden["primes"].items():
        n, x = (int(v) for v in key.split(","))
        assert Port.primes_near(n, x) == exp, key


def test_murmur(golden):
    for m in golden["murmur"]:
        assert [str(v) for v in Port.murmur3_x64_128(m["key"], m["seed"])] == m["h"], m["key"]
    assert Port.murmur3_x64_128("ACG", 42)[0] == 1731421407650554201  # sourmash's documented hash_murmur("ACG")


def test_reference_fixture_random20a(golden):
    fx = golden["random20a"]
    bases = np.frombuffer("".join(fx["reads"]).encode(), dtype=np.uint8)
    offsets = np.zeros(len(fx["reads"]) + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([len(r) for r in fx["reads"]])
    assert len(fx["reads"]) == 99
    for c in fx["cases"]:
        g = Port(c["kind"], c["can"], c["K"], c["sizes"])
        tot, _ = g.insert_reads(bases, offsets)
        assert tot == c["n_kmers"]
        assert g.stats() == (c["n_unique"], c["n_occupied"])
        _check_tables(g, c["tables"])
        fw, rc = Port.hash_sequence(c["can"], c["K"], fx["reads"][0])
        for i, (efw, erc) in enumerate(c["first_read_hashes"]):
            assert str(int(fw[i])) == efw
            if c["can"]:
                assert str(int(rc[i])) == erc
        if "query_first_read_after_3_passes" in c:
            g.insert_reads(bases, offsets)
            g.insert_reads(bases, offsets)
            assert [int(v) for v in g.query_sequence(fx["reads"][0])] == c["query_first_read_after_3_passes"]
            assert g.stats()[0] == c["n_unique_after_3_passes"]
    # literals quoted in SURVEY.md section 8c (Bit/Can K=31)
    c0 = fx["cases"][0]
    assert (c0["n_kmers"], c0["n_unique"], c0["n_occupied"]) == (2871, 2871, 2864)
    assert c0["tables"]["fnv"][0] == "15765873264133613085"


def test_synthetic_golden(golden):
    for c in golden["synthetic"]:
        bases, offsets = _inputs(c["input"])
        g = Port(c["kind"], c["can"], c["K"], c["sizes"])
        tot = 0
        first_new = None
        for _ in range(c["passes"]):
            t, _, n_new = g.insert_reads(bases, offsets, want_n_new=True)
            tot += t
            if first_new is None:
                first_new = n_new.copy()
        assert tot == c["n_kmers"], c["input"]
        assert g.stats() == (c["n_unique"], c["n_occupied"])
        assert str(fnv(first_new.astype(np.uint64).view(np.uint8))) == c["n_new_first_pass_fnv"]
        _check_tables(g, c["tables"])
        first = read_str(bases, offsets, 0)
        if len(first) >= c["K"]:
            assert [int(v) for v in g.query_sequence(first)] == c["query_first_read"]


def test_insert_and_query_sequence_golden(golden):
    for c in golden["insert_and_query_sequence"]:
        g = Port(c["kind"], c["can"], c["K"], c["sizes"])
        assert [int(v) for v in g.insert_and_query_sequence(c["seq"])] == c["first"]
        assert [int(v) for v in g.insert_and_query_sequence(c["seq"])] == c["second"]
        assert g.stats()[0] == c["n_unique"]


def test_median_count_at_least_golden(golden):
    m = golden["median_count_at_least"]
    b, o = _inputs(m["insert"])
    qb, qo = _inputs(m["query"])
    for c in m["cases"]:
        g = Port(c["kind"], c["can"], c["K"], c["sizes"])
        g.insert_reads(b, o)
        for cutoff, exp in c["pass"].items():
            got = [int(g.median_count_at_least(read_str(qb, qo, r), int(cutoff))) for r in range(len(exp))]
            assert got == exp, (c["kind"], cutoff)


def test_oxli_save_golden(golden):
    b, o = _inputs(golden["oxli"]["input"])
    with tempfile.TemporaryDirectory() as td:
        for c in golden["oxli"]["files"]:
            g = Port(c["kind"], c["can"], c["K"], c["sizes"])
            g.insert_reads(b, o)
            fn = os.path.join(td, "t.oxli")
            g.save(fn)
            data = np.fromfile(fn, dtype=np.uint8)
            assert data.size == c["file_bytes"]
            assert data[:32].tobytes().hex() == c["head_hex"]
            assert str(fnv(data)) == c["file_fnv"]


def test_diginorm_batch_one_is_serial():
    """batch == 1 of the batch-synchronous rule is the reference's read-at-a-time filter."""
    b, o = genome_reads(600, 80, 1500, seed=3)
    sizes = Port.primes_near(4, 500000)
    g = Port(1, 1, 21, sizes)
    keep = g.diginorm_reads(b, o, cutoff=5, batch=1)
    h = Port(1, 1, 21, sizes)
    exp = []
    for r in range(600):
        s = read_str(b, o, r)
        if h.median_count_at_least(s, 5):
            exp.append(0)
        else:
            h.insert_sequence(s)
            exp.append(1)
    assert keep.tolist() == exp
    assert 0 < sum(exp) < 600


def test_sketch_restatement_properties():
    """PARITY UNPINNED (libsourmash absent): only internal consistency + the pinned murmur step."""
    seq = read_str(*synth_reads(1, 2000, seed=4), 0)
    comp = str.maketrans("ACGT", "TGCA")
    K = 31
    sk = PortSketch(0, K, 42, scaled=10)
    assert sk.insert_sequence(seq) == len(seq) - K + 1
    mins = sk.mins()
    assert np.all(mins[1:] > mins[:-1]) and mins.size > 0 and int(mins[-1]) <= sk.max_hash
    hs = set()
    for i in range(len(seq) - K + 1):
        k = seq[i:i + K]
        rc = k.translate(comp)[::-1]
        h = Port.murmur3_x64_128(min(k, rc), 42)[0]
        if h <= sk.max_hash:
            hs.add(h)
    assert sorted(hs) == mins.tolist()
    # strand symmetry and N handling
    sk2 = PortSketch(0, K, 42, scaled=10)
    sk2.insert_sequence(seq.translate(comp)[::-1])
    assert np.array_equal(sk2.mins(), mins)
    sk3 = PortSketch(0, K, 42, scaled=10)
    sk3.insert_sequence(seq[:700].lower() + "N" + seq[701:])
    assert set(sk3.mins().tolist()) <= hs
    # bottom-k
    skn = PortSketch(50, K, 42, scaled=0)
    skn.insert_sequence(seq)
    allh = sorted({Port.murmur3_x64_128(min(seq[i:i + K], seq[i:i + K].translate(comp)[::-1]), 42)[0]
                   for i in range(len(seq) - K + 1)})
    assert skn.mins().tolist() == allh[:50]
    assert Port.max_hash_from_scaled(1000) == 18446744073709552
    # upstream sourmash's own known answer (tests/test__minhash.py::test_basic_dna): pins canonical choice +
    # murmur seed 42 + "first 64 bits" + bottom-k for the restatement
    kat = PortSketch(1, 4, 42, scaled=0)
    kat.insert_sequence("ATGC")
    assert kat.mins().tolist() == [12415348535738636339]
    kat.insert_sequence("GCAT")
    assert kat.mins().tolist() == [12415348535738636339]


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_port_equals_reference_on_random_inputs():
    """Where the compiled reference is present, check the restatement against it directly."""
    sizes = Ref.primes_near(4, 700000)
    for seed, (kind, can, K) in enumerate([(0, 1, 31), (0, 0, 21), (1, 1, 21), (1, 0, 33), (2, 1, 25), (2, 0, 64)]):
        b, o = genome_reads(1500, 120, 3000, seed=100 + seed)
        r, p = Ref(kind, can, K, sizes), Port(kind, can, K, sizes)
        tr, _, nr = r.insert_reads(b, o, want_n_new=True)
        tp, _, np_ = p.insert_reads(b, o, want_n_new=True)
        assert tr == tp and np.array_equal(nr, np_) and r.stats() == p.stats()
        for a, c in zip(r.tables(), p.tables()):
            assert np.array_equal(a, c)
        s = read_str(b, o, 3)
        assert np.array_equal(r.query_sequence(s), p.query_sequence(s))
        assert np.array_equal(r.insert_and_query_sequence(s), p.insert_and_query_sequence(s))
        hs = np.random.default_rng(seed).integers(0, 2**64, 3000, dtype=np.uint64)
        assert np.array_equal(r.insert_hashes(hs), p.insert_hashes(hs))
        assert np.array_equal(r.query_hashes(hs), p.query_hashes(hs))
