"""GPU parity tests of the write-combining insert path (goetia_b200/csrc/bucket.cuh): blind
inserts are bucketed by table slice and applied slice by slice; the final tables must equal
the CPU oracle's byte for byte, whatever the slice size, store size or input skew.

The path normally engages only for tables larger than L2; the tests force it on small tables
through the GT_BUCKET_* environment knobs the library reads when a storage builds its store.
"""
import os

import numpy as np
import pytest

from tests.util import (Port, SHIFTERS, STORAGES, assert_tables_equal, genome_reads, make_graph, ragged_reads,
                        read_str, synth_reads)

pytestmark = pytest.mark.gpu


# How the buckets are applied (bucket.cuh): "atomics" = k_apply (global RED / CAS per update); "win1" = k_apply_win
# with one shared-memory window per slice; "win2" = k_rebucket2 into 128-byte windows + k_apply_win (two-level);
# "win2tiny": the same with k_rebucket2's staging rows cut to 8 entries, so that rows overflow all the time and the excess
# takes the direct route.
APPLY = {"atomics": ("0", "16", None), "win1": ("2", "16", None), "win2": ("2", "7", None), "win2tiny": ("2", "7", "8")}


@pytest.fixture(params=sorted(APPLY))
def forced(monkeypatch, request):
    """Force the bucket path with tiny slices and a tiny store (many buckets, many flushes)."""
    mode, wlog2, row = APPLY[request.param]
    monkeypatch.setenv("GT_APPLY_WINDOWS", mode)
    monkeypatch.setenv("GT_WINDOW_LOG2_BYTES", wlog2)
    if row:
        monkeypatch.setenv("GT_REBUCKET_ROW", row)

    def set_env(slice_log2=12, entries=1 << 20, min_kmers=0):
        monkeypatch.setenv("GT_BUCKET_FORCE", "1")
        monkeypatch.setenv("GT_SLICE_LOG2_BYTES", str(slice_log2))
        monkeypatch.setenv("GT_PENDING_ENTRIES", str(entries))
        monkeypatch.setenv("GT_BUCKET_MIN_KMERS", str(min_kmers))
    set_env()
    set_env.apply = request.param
    return set_env


@pytest.mark.parametrize("kind,_n", STORAGES)
@pytest.mark.parametrize("can,_s", SHIFTERS)
@pytest.mark.parametrize("pieces", [0, 2])
def test_bucketed_insert_bit_exact(gb, forced, monkeypatch, kind, _n, can, _s, pieces):
    # pieces=2: room in the global buckets is taken a piece at a time (holes padded) even though the
    # store is tiny, so most pieces run past the capacity -- the tables must still be exact
    monkeypatch.setenv("GT_BUCKET_PIECES", str(pieces))
    if pieces:
        forced(slice_log2=14, entries=48 << 20)
    K = [31, 21, 25][kind]
    sizes = gb.get_n_primes_near_x(4, 1_000_000)
    bases, offsets = genome_reads(12000, 150, 30000, seed=100 + kind * 10 + can)
    g = make_graph(gb, kind, can, K, sizes)
    n = g.insert_sequences(bases, offsets, mode=0)
    info = g.S.pending_info()
    assert info["built"] == 1 and info["n_buckets"] > 4, info
    ref = Port(kind, can, K, sizes)
    n_ref, _ = ref.insert_reads(bases, offsets)
    assert n == n_ref
    assert_tables_equal(g.get_raw(), ref.tables())
    assert g.n_occupied() == ref.stats()[1]
    assert g.S.pending_info()["pending_kmers"] == 0
    q = g.query_sequences(bases[:150 * 300], offsets[:301])
    assert np.array_equal(q, ref.query_reads(bases[:150 * 300], offsets[:301]))


@pytest.mark.parametrize("kind,_n", STORAGES)
def test_bucketed_ragged_invalid_reads(gb, forced, kind, _n):
    K = 21
    sizes = gb.get_n_primes_near_x(4, 500_000)
    bases, offsets = ragged_reads(4000, 0, 300, seed=15, alphabet=b"ACGTacgt")
    bases = bases.copy()
    rng = np.random.default_rng(19)
    lens = (offsets[1:] - offsets[:-1]).astype(np.int64)
    bad = np.nonzero((rng.random(4000) < 0.02) & (lens > 0))[0]
    for r in bad:
        bases[int(offsets[r]) + int(rng.integers(0, lens[r]))] = ord("N")
    g = make_graph(gb, kind, 1, K, sizes)
    tot, status = g.insert_sequences(bases, offsets, mode=0, want_status=True)
    assert g.S.pending_info()["built"] == 1
    ref = Port(kind, 1, K, sizes)
    upper = np.frombuffer(bases.tobytes().upper(), dtype=np.uint8)
    exp = 0
    for r in range(4000):
        s = read_str(upper, offsets, r)
        if all(c in "ACGT" for c in s) and len(s) >= K:
            exp += ref.insert_sequence(s)[0]
    assert tot == exp
    assert_tables_equal(g.get_raw(), ref.tables())


@pytest.mark.parametrize("kind,_n", STORAGES)
def test_bucket_overflow_goes_direct(gb, forced, kind, _n):
    """Skewed input (few distinct k-mers, repeated) overflows its buckets; the excess is applied
    directly and the tables are still exact -- and the counters saturate exactly."""
    K = 21
    sizes = gb.get_n_primes_near_x(4, 400_000)
    one = np.concatenate([np.full(150, ord("A"), dtype=np.uint8), synth_reads(2, 150, seed=77)[0]])
    bases = np.tile(one, 4000)  # every third read is poly-A: one k-mer, one slot per table
    offsets = np.arange(12001, dtype=np.uint64) * np.uint64(150)
    g = make_graph(gb, kind, 1, K, sizes)
    g.insert_sequences(bases, offsets, mode=0)
    g.flush()
    assert g.S.pending_info()["n_direct"] > 0
    ref = Port(kind, 1, K, sizes)
    ref.insert_reads(bases, offsets)
    assert_tables_equal(g.get_raw(), ref.tables())
    if kind:
        assert int(g.query_sequences(bases[:150], offsets[:2]).max()) == (255 if kind == 1 else 15)


@pytest.mark.parametrize("n_tables", [1, 3, 5])
def test_bucketed_other_table_counts(gb, forced, n_tables):
    sizes = gb.get_n_primes_near_x(n_tables, 700_000)
    bases, offsets = synth_reads(8000, 100, seed=31)
    g = make_graph(gb, 0, 1, 25, sizes)
    g.insert_sequences(bases, offsets, mode=0)
    ref = Port(0, 1, 25, sizes)
    ref.insert_reads(bases, offsets)
    assert_tables_equal(g.get_raw(), ref.tables())


@pytest.mark.parametrize("kind,_n", STORAGES)
def test_bucketed_resident_and_device_ascii(gb, forced, kind, _n):
    import torch
    from goetia_b200.batch import PackedBatch
    forced(slice_log2=13, entries=8 << 20)
    K = 31
    sizes = gb.get_n_primes_near_x(4, 2_000_000)
    bases, offsets = genome_reads(10000, 150, 50000, seed=41)
    ref = Port(kind, 1, K, sizes)
    ref.insert_reads(bases, offsets)
    ref.insert_reads(bases, offsets)
    g = make_graph(gb, kind, 1, K, sizes)
    pb = PackedBatch.from_host(bases, offsets)
    assert pb.insert_into(g, mode=0) == 10000 * 120
    d_b = torch.from_numpy(bases).cuda()
    d_o = torch.from_numpy(offsets.astype(np.int64)).cuda()
    torch.cuda.synchronize()
    assert g.insert_sequences_dev(d_b.data_ptr(), d_o.data_ptr(), 10000, bases.size, mode=0) == 10000 * 120
    assert g.S.pending_info()["pending_kmers"] > 0
    assert_tables_equal(g.get_raw(), ref.tables())


def test_pending_is_visible_to_every_reader(gb, forced):
    """Blind inserts may sit in the buckets; query / stats / tracked inserts must see them."""
    forced(slice_log2=12, entries=64 << 20)
    K = 31
    sizes = gb.get_n_primes_near_x(4, 1_000_000)
    bases, offsets = synth_reads(9000, 150, seed=51)
    g = make_graph(gb, 0, 1, K, sizes)
    g.insert_sequences(bases, offsets, mode=0)
    assert g.S.pending_info()["pending_kmers"] > 0
    q = g.query_sequences(bases[:150 * 100], offsets[:101])
    assert bool((q == 1).all())
    g.insert_sequences(bases, offsets, mode=0)
    assert g.S.pending_info()["pending_kmers"] > 0
    # a tracked insert of the same reads finds nothing new
    before = g.n_unique()
    g.insert_sequences(bases[:150 * 2000], offsets[:2001], mode=1)
    assert g.n_unique() == before
    # reset drops pending updates
    g.insert_sequences(bases, offsets, mode=0)
    g.reset()
    assert g.n_occupied() == 0
    assert not any(t.any() for t in g.get_raw())


def test_default_policy_small_tables_stay_direct(gb, monkeypatch):
    for k in ("GT_BUCKET_FORCE", "GT_SLICE_LOG2_BYTES", "GT_PENDING_ENTRIES", "GT_BUCKET_MIN_KMERS"):
        monkeypatch.delenv(k, raising=False)
    sizes = gb.get_n_primes_near_x(4, 1_000_000)
    bases, offsets = synth_reads(20000, 150, seed=61)
    g = make_graph(gb, 0, 1, 31, sizes)
    g.insert_sequences(bases, offsets, mode=0)
    assert g.S.pending_info()["built"] == 0  # tables fit in L2: atomics already run at the L2 rate
    ref = Port(0, 1, 31, sizes)
    ref.insert_reads(bases, offsets)
    assert_tables_equal(g.get_raw(), ref.tables())


def test_large_tables_default_policy(gb, monkeypatch):
    """Default knobs, tables larger than L2 (4 x 2^31 bits = 1 GB): the bucket path engages by itself."""
    for k in ("GT_BUCKET_FORCE", "GT_SLICE_LOG2_BYTES", "GT_PENDING_ENTRIES", "GT_BUCKET_MIN_KMERS"):
        monkeypatch.delenv(k, raising=False)
    K = 31
    sizes = gb.get_n_primes_near_x(4, 2**31)
    bases, offsets = synth_reads(40000, 150, seed=71)
    g = make_graph(gb, 0, 1, K, sizes)
    g.insert_sequences(bases, offsets, mode=0)
    info = g.S.pending_info()
    assert info["built"] == 1 and info["slice_shift"] == 28, info
    ref = Port(0, 1, K, sizes)
    ref.insert_reads(bases, offsets)
    # compare through a checksum of each table (256 MB each)
    for a, b in zip(g.get_raw(), ref.tables()):
        assert a.size == b.size and Port.fnv1a(a) == Port.fnv1a(b)
    assert g.n_occupied() == ref.stats()[1]
    g.S.close()


@pytest.mark.parametrize("x", [2**28 + 1000, 2**29 + 12345, 2**30 + 99, 3_000_000_000, 2**32 + 5000])
def test_big_table_reducers(gb, monkeypatch, x):
    """Tables of >= 2^28 slots take the 32-bit-reciprocal reduction (bin_of, rs = 4..0): default knobs,
    tables compared with the oracle through checksums."""
    for k in ("GT_BUCKET_FORCE", "GT_SLICE_LOG2_BYTES", "GT_PENDING_ENTRIES", "GT_BUCKET_MIN_KMERS"):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setenv("GT_PENDING_ENTRIES", str(64 << 20))
    monkeypatch.setenv("GT_BUCKET_FORCE", "1")
    K = 31
    sizes = gb.get_n_primes_near_x(2, x)
    bases, offsets = synth_reads(30000, 150, seed=81)
    g = make_graph(gb, 0, 1, K, sizes)
    g.insert_sequences(bases, offsets, mode=0)
    assert g.S.pending_info()["built"] == 1
    ref = Port(0, 1, K, sizes)
    ref.insert_reads(bases, offsets)
    for a, b in zip(g.get_raw(), ref.tables()):
        assert a.size == b.size and Port.fnv1a(a) == Port.fnv1a(b)
    g.S.close()


@pytest.mark.parametrize("kind,_n", [(1, "ByteStorage"), (2, "NibbleStorage")])
def test_bucketed_counting_saturation(gb, forced, monkeypatch, kind, _n):
    """Counting storages through the window apply: a window's image in shared memory holds the hits of one apply per
    counter, saturating at the counter's maximum, and is merged with a per-counter saturating add (bucket.cuh, K2w).
    Deep coverage drives many counters through their maximum inside the buckets AND inside one apply; tables must
    still equal the oracle's min(max, hits) byte for byte, over several applies, under every apply variant."""
    forced(slice_log2=13, entries=1 << 24)
    K = 21
    sizes = gb.get_n_primes_near_x(4, 700_001)
    bases, offsets = genome_reads(30000, 100, 2500, seed=31)  # ~1000x coverage of 2.5 kb: counts far past 255
    g = make_graph(gb, kind, 1, K, sizes)
    ref = Port(kind, 1, K, sizes)
    for _ in range(2):
        for r0 in range(0, 30000, 10000):  # three calls per pass: several stores / applies
            b = bases[int(offsets[r0]):int(offsets[r0 + 10000])]
            o = offsets[r0:r0 + 10001] - offsets[r0]
            g.insert_sequences(b, o, mode=0)
            ref.insert_reads(b, o)
    g.flush()
    assert g.S.pending_info()["built"]
    assert_tables_equal(g.get_raw(), ref.tables())
    q = g.query_sequences(bases[:300], offsets[:4])
    assert int(q.max()) == (255 if kind == 1 else 15)
    # mixed: unique reads on top (no saturation in most slices), then the tables once more
    ub, uo = synth_reads(5000, 100, seed=32)
    g.insert_sequences(ub, uo, mode=0)
    ref.insert_reads(ub, uo)
    assert_tables_equal(g.get_raw(), ref.tables())
    ref.close()
