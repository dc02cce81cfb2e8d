"""GT_MODE_EXACT: order-dependent outputs under the serial first-toucher rule.

The reference inserts k-mers one at a time (dbg.hh:307-318 counts `n_new` from Storage::insert's
return value: bitstorage.hh:195-219, bytestorage.cc:60-113, nibblestorage.cc:60-100), so whether a
k-mer is "new" depends on everything inserted before it.  The GPU's exact mode must reproduce the
per-read n_new, the running n_unique and the per-hash is_new of that serial loop bit for bit --
including on tiny, collision-heavy tables where the atomic-winner rule (GT_MODE_FAST) differs.
"""
import os

import numpy as np
import pytest

from tests.util import Port, STORAGES, assert_tables_equal, genome_reads, make_graph, ragged_reads

pytestmark = pytest.mark.gpu

EXACT = 2


@pytest.mark.parametrize("kind,_n", STORAGES)
@pytest.mark.parametrize("can", [0, 1])
@pytest.mark.parametrize("x", [5_000, 400_000])
def test_n_new_per_read_serial_rule(gb, kind, _n, can, x):
    """Small tables: most slots collide, many k-mers repeat inside the batch."""
    K = [31, 21, 25][kind]
    sizes = gb.get_n_primes_near_x(4, x)
    bases, offsets = genome_reads(4000, 150, 30000, seed=31 + kind * 2 + can)
    g = make_graph(gb, kind, can, K, sizes)
    ref = Port(kind, can, K, sizes)
    for rnd in range(2):  # the second pass sees a populated table: almost nothing is new
        tot, n_new = g.insert_sequences(bases, offsets, mode=EXACT, want_n_new=True)
        tot_ref, _, n_new_ref = ref.insert_reads(bases, offsets, want_n_new=True)
        assert tot == tot_ref
        assert np.array_equal(n_new, n_new_ref), (rnd, int(np.nonzero(n_new != n_new_ref)[0][0]))
        assert g.n_unique() == ref.stats()[0]
        assert g.n_occupied() == ref.stats()[1]
        assert_tables_equal(g.get_raw(), ref.tables())


def test_exact_differs_from_atomic_winner_only_where_stated(gb):
    """On a crowded table the atomic-winner count may differ from the serial count; exact may not."""
    K, sizes = 21, [1009, 1013, 1019, 1021]
    bases, offsets = ragged_reads(500, 10, 200, seed=77)
    ref = Port(0, 1, K, sizes)
    _, _, n_new_ref = ref.insert_reads(bases, offsets, want_n_new=True)
    g = make_graph(gb, 0, 1, K, sizes)
    _, n_new = g.insert_sequences(bases, offsets, mode=EXACT, want_n_new=True)
    assert np.array_equal(n_new, n_new_ref)
    assert g.n_unique() == ref.stats()[0] == int(n_new_ref.sum())


@pytest.mark.parametrize("kind,_n", STORAGES)
def test_exact_multiple_claim_ranges_and_chunks(gb, kind, _n, monkeypatch):
    """Tiny claim map (many tile ranges per chunk) and several pipeline chunks: order must hold across both."""
    monkeypatch.setenv("GT_EXACT_LOG2_CAP", "16")  # 2^16 slots -> one 8192-position tile per range
    K = 25
    sizes = gb.get_n_primes_near_x(4, 150_000)
    bases, offsets = genome_reads(5000, 120, 20000, seed=5 + kind)  # 600 kb -> 3 chunks of 200 kb (conftest)
    g = make_graph(gb, kind, 1, K, sizes)
    ref = Port(kind, 1, K, sizes)
    _, n_new = g.insert_sequences(bases, offsets, mode=EXACT, want_n_new=True)
    _, _, n_new_ref = ref.insert_reads(bases, offsets, want_n_new=True)
    assert np.array_equal(n_new, n_new_ref)
    assert g.n_unique() == ref.stats()[0]
    assert_tables_equal(g.get_raw(), ref.tables())


@pytest.mark.parametrize("kind,_n", STORAGES)
def test_is_new_per_hash_serial_rule(gb, kind, _n, monkeypatch):
    """Storage::insert over a hash vector with duplicates and colliding values."""
    sizes = [997, 991, 983, 977]
    rng = np.random.default_rng(kind)
    hs = rng.integers(0, 2**64, 30000, dtype=np.uint64)
    hs = np.concatenate([hs, hs[:5000], hs[::-1][:5000]])  # repeats later in the stream
    st = [gb.BitStorage, gb.ByteStorage, gb.NibbleStorage][kind](sizes)
    ref = Port(kind, 0, 21, sizes)
    for cap in ("27", "12"):
        monkeypatch.setenv("GT_EXACT_LOG2_CAP", cap)
        got = st.insert_many(hs, mode=EXACT)
        exp = ref.insert_hashes(hs)
        assert np.array_equal(np.asarray(got, dtype=np.uint8), exp)
        assert st.n_unique_kmers() == ref.stats()[0]
        assert_tables_equal(st.get_raw_tables(), ref.tables())


def test_exact_after_blind_sees_pending_updates(gb):
    """A tracked insert must see every earlier write-combined (pending) insert."""
    K = 31
    sizes = gb.get_n_primes_near_x(4, 2_000_000)
    bases, offsets = genome_reads(3000, 150, 40000, seed=3)
    os.environ["GT_BUCKET_FORCE"] = "1"
    os.environ["GT_BUCKET_MIN_KMERS"] = "1"
    try:
        g = make_graph(gb, 0, 1, K, sizes)
        ref = Port(0, 1, K, sizes)
        g.insert_sequences(bases[:150 * 1500], offsets[:1501], mode=0)  # blind: stays pending
        ref.insert_reads(bases[:150 * 1500], offsets[:1501])
        before = ref.stats()[0]
        _, n_new = g.insert_sequences(bases, offsets, mode=EXACT, want_n_new=True)
        _, _, n_new_ref = ref.insert_reads(bases, offsets, want_n_new=True)
        assert np.array_equal(n_new, n_new_ref)
        assert int(n_new.sum()) == ref.stats()[0] - before
        assert_tables_equal(g.get_raw(), ref.tables())
    finally:
        os.environ.pop("GT_BUCKET_FORCE", None)
        os.environ.pop("GT_BUCKET_MIN_KMERS", None)
