"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle, bit for bit.

Modelled on the reference's tests/test_dbg.py and tests/test_hashing.py (storage x shifter
cartesian product, known-answer vector, rolling == from-scratch, storage semantics).
"""
import numpy as np
import pytest

from tests.util import (Port, SHIFTERS, STORAGES, assert_tables_equal, genome_reads, make_graph, ragged_reads,
                        read_str, synth_reads)

pytestmark = pytest.mark.gpu

KAT_SEQ = "TCACCTGTGTTGTGCTACTTGCGGCGC"  # reference tests/test_hashing.py:15-19
KAT_FW, KAT_RC = 13194817695400542713, 4324216031038051805


def test_hash_known_answer(gb):
    fw = gb.FwdLemireShifter(27).hash(KAT_SEQ)
    can = gb.CanLemireShifter(27).hash(KAT_SEQ)
    assert fw.value() == KAT_FW
    assert (can.fw_hash, can.rc_hash) == (KAT_FW, KAT_RC)
    assert can.value() == KAT_RC
    # host cursor path agrees with the kernel
    assert gb.CanLemireShifter(27).hash_base(KAT_SEQ) == can


@pytest.mark.parametrize("K", [1, 2, 21, 27, 31, 32, 33, 63, 64, 65, 101, 200])
@pytest.mark.parametrize("can", [0, 1])
def test_hash_sequences_vs_oracle(gb, K, can):
    bases, offsets = ragged_reads(300, 0, 400, seed=K * 2 + can)
    sh = [gb.FwdLemireShifter, gb.CanLemireShifter][can](K)
    fw, rc, status = sh.hash_sequences(bases, offsets)
    pos = 0
    for r in range(offsets.size - 1):
        s = read_str(bases, offsets, r)
        if len(s) < K:
            assert status[r] & 1
            continue
        assert status[r] == 0
        efw, erc = Port.hash_sequence(can, K, s)
        n = efw.size
        assert np.array_equal(fw[pos:pos + n], efw), (r, K)
        if can:
            assert np.array_equal(rc[pos:pos + n], erc), (r, K)
        pos += n
    assert pos == fw.size


@pytest.mark.parametrize("kind,_n", STORAGES)
@pytest.mark.parametrize("can,_s", SHIFTERS)
@pytest.mark.parametrize("mode", [0, 1])
def test_insert_tables_bit_exact(gb, kind, _n, can, _s, mode):
    K = [31, 21, 25][kind]
    sizes = gb.get_n_primes_near_x(4, 3_000_000)
    bases, offsets = genome_reads(6000, 150, 20000, seed=kind * 10 + can)
    g = make_graph(gb, kind, can, K, sizes)
    n = g.insert_sequences(bases, offsets, mode=mode)
    ref = Port(kind, can, K, sizes)
    n_ref, _ = ref.insert_reads(bases, offsets)
    assert n == n_ref == 6000 * (150 - K + 1)
    assert_tables_equal(g.get_raw(), ref.tables())
    n_unique, n_occ = ref.stats()
    assert g.n_occupied() == n_occ
    if mode == 1:
        # atomic-winner rule: exact unless two distinct k-mers of the batch collide on a fresh slot
        assert abs(g.n_unique() - n_unique) <= max(2, n_unique // 200)
    # query every k-mer back
    q = g.query_sequences(bases[:150 * 500], offsets[:501])
    assert np.array_equal(q, ref.query_reads(bases[:150 * 500], offsets[:501]))


@pytest.mark.parametrize("kind,_n", STORAGES)
def test_ragged_short_and_invalid_reads(gb, kind, _n):
    K = 21
    sizes = gb.get_n_primes_near_x(4, 1_000_000)
    bases, offsets = ragged_reads(3000, 0, 300, seed=5, alphabet=b"ACGTacgt")
    bases = bases.copy()
    # poison ~2 % of the reads with one non-ACGT byte
    rng = np.random.default_rng(9)
    lens = (offsets[1:] - offsets[:-1]).astype(np.int64)
    bad = np.nonzero((rng.random(3000) < 0.02) & (lens > 0))[0]
    for r in bad:
        bases[int(offsets[r]) + int(rng.integers(0, lens[r]))] = ord(rng.choice(list("NnXR-")))
    g = make_graph(gb, kind, 1, K, sizes)
    tot, n_new, status = g.insert_sequences(bases, offsets, want_n_new=True, want_status=True)
    # oracle sees what FastxParser + InserterProcessor would pass on: upper-cased, valid, len >= K
    ref = Port(kind, 1, K, sizes)
    exp_tot = 0
    upper = np.frombuffer(bases.tobytes().upper(), dtype=np.uint8)
    for r in range(3000):
        s = read_str(upper, offsets, r)
        invalid = any(c not in "ACGT" for c in s)
        assert bool(status[r] & 2) == invalid, r
        assert bool(status[r] & 1) == (len(s) < K), r
        if not invalid and len(s) >= K:
            exp_tot += ref.insert_sequence(s)[0]
    assert tot == exp_tot
    assert_tables_equal(g.get_raw(), ref.tables())
    assert int(n_new.sum()) == g.n_unique()
    assert np.all(n_new[status != 0] == 0)


@pytest.mark.parametrize("kind,_n", STORAGES)
def test_hash_vector_members_and_fastmod_edges(gb, kind, _n):
    sizes = [999983, 65537, 4, 1]  # includes tiny and degenerate divisors
    rng = np.random.default_rng(3)
    hs = rng.integers(0, 2**64, 20000, dtype=np.uint64)
    edge = np.array([0, 1, 2, 3, 4, 65536, 65537, 65538, 999982, 999983, 999984, 2**32 - 1, 2**32, 2**63 - 1, 2**63,
                     2**64 - 2, 2**64 - 1] + [999983 * k for k in range(1, 50)], dtype=np.uint64)
    hs = np.concatenate([hs, edge, edge])
    st = [gb.BitStorage, gb.ByteStorage, gb.NibbleStorage][kind](sizes)
    ref = Port(kind, 0, 21, sizes)
    st.insert_many(hs, mode=0)
    ref.insert_hashes(hs)
    assert_tables_equal(st.get_raw_tables(), ref.tables())
    assert np.array_equal(st.query_many(hs), ref.query_hashes(hs))
    big = [2**40 + 15, 2**62 + 1, 2**63]  # table sizes beyond 2^32: only the arithmetic is exercised
    from tests.csrc_check import fastmod_host
    for d in big + sizes:
        for h in edge.tolist() + hs[:200].tolist():
            assert fastmod_host(int(h), d) == int(h) % d


@pytest.mark.parametrize("kind,maxc", [(1, 255), (2, 15)])
def test_counter_saturation(gb, kind, maxc):
    K = 21
    sizes = gb.get_n_primes_near_x(4, 100_000)
    one, _ = synth_reads(1, 60, seed=1)
    reps = maxc + 45
    bases = np.tile(one, reps)
    offsets = np.arange(reps + 1, dtype=np.uint64) * np.uint64(60)
    g = make_graph(gb, kind, 1, K, sizes)
    g.insert_sequences(bases, offsets)
    ref = Port(kind, 1, K, sizes)
    ref.insert_reads(bases, offsets)
    assert_tables_equal(g.get_raw(), ref.tables())
    assert set(g.query_sequence(one.tobytes().decode())) == {maxc}


@pytest.mark.parametrize("kind,_n", STORAGES)
@pytest.mark.parametrize("K", [21, 101])
def test_single_kmer_semantics(gb, kind, _n, K):
    """reference tests/test_dbg.py:17-101 (presence / counting insert, insert_and_query, 10 passes)."""
    sizes = gb.get_n_primes_near_x(4, 1_000_000)
    g = make_graph(gb, kind, 1, K, sizes)
    seq = read_str(*synth_reads(1, K + 12, seed=K), 0)
    kmers = [seq[i:i + K] for i in range(len(seq) - K + 1)]
    counting = kind != 0
    for kmer in kmers:
        h = g.hash(kmer)
        assert g.query(kmer) == 0 and g.query(h) == 0
        assert g.insert(kmer) is True
        assert g.query(kmer) == 1 and g.get(h) == 1
        assert g.insert(kmer) is False
        assert g.query(kmer) == (2 if counting else 1)
    g.reset()
    assert g.n_unique() == 0 and g.n_occupied() == 0
    for kmer in kmers:
        assert g.insert_and_query(kmer) == 1
        assert g.insert_and_query(g.hash(kmer)) == (2 if counting else 1)
    assert g.hash("A" * K) == g.hash("A" * K + "TTTT")  # tests/test_dbg.py:125-128
    # the graph is a shifter too: cursor members and pythonize_dbg.py's get_hash
    assert not g.is_initialized()
    assert g.set_cursor(kmers[0]) == g.hash(kmers[0]) == g.get_hash()
    assert g.shift_right(seq[0], seq[K]) == g.hash(kmers[1]) == g.get_hash()
    assert g.shift_left(seq[0], seq[K]) == g.hash(kmers[0])
    with pytest.raises(Exception):
        g.insert_sequence("A" * (K - 1))  # tests/test_dbg.py:222-227
    c = g.clone()
    assert c.n_unique() == 0 and c.S.get_tablesizes() == g.S.get_tablesizes()  # :297-305


@pytest.mark.parametrize("kind,_n", [(1, "ByteStorage"), (2, "NibbleStorage")])
def test_counting_passes(gb, kind, _n):
    """tests/test_dbg.py:79-101: after i insert_sequence passes every k-mer counts i."""
    K = 21
    g = make_graph(gb, kind, 0, K, gb.get_n_primes_near_x(4, 1_000_000))
    seq = read_str(*synth_reads(1, 500, seed=11), 0)
    for it in range(10):
        assert set(g.query_sequence(seq)) == {it}
        assert g.insert_sequence(seq) == len(seq) - K + 1
    assert g.n_unique() == len(seq) - K + 1


@pytest.mark.parametrize("kind,_n", [(1, "ByteStorage"), (2, "NibbleStorage")])
def test_median_count_at_least(gb, kind, _n):
    K = 21
    sizes = gb.get_n_primes_near_x(4, 2_000_000)
    bases, offsets = genome_reads(4000, 100, 5000, seed=21)
    g = make_graph(gb, kind, 1, K, sizes)
    ref = Port(kind, 1, K, sizes)
    g.insert_sequences(bases, offsets)
    ref.insert_reads(bases, offsets)
    qb, qo = genome_reads(1500, 100, 5000, seed=22, sub_rate=0.08)
    qb2, qo2 = ragged_reads(300, 0, 60, seed=23)
    qbases = np.concatenate([qb, qb2])
    qoffs = np.concatenate([qo, qo2[1:] + qo[-1]])
    for cutoff in (1, 5, 14, 40):
        got = g.median_count_at_least(qbases, qoffs, cutoff)
        for r in range(qoffs.size - 1):
            s = read_str(qbases, qoffs, r)
            exp = ref.median_count_at_least(s, cutoff) if len(s) >= K else False
            assert bool(got[r]) == exp, (r, cutoff)


def test_long_reads_nibble(gb):
    """config 5 shape in miniature: 10 kb reads, NibbleStorage K=25."""
    K = 25
    sizes = gb.get_n_primes_near_x(4, 5_000_000)
    bases, offsets = genome_reads(60, 10000, 40000, seed=31)
    g = make_graph(gb, 2, 1, K, sizes)
    ref = Port(2, 1, K, sizes)
    assert g.insert_sequences(bases, offsets) == ref.insert_reads(bases, offsets)[0]
    assert_tables_equal(g.get_raw(), ref.tables())


def test_update_from_and_resident_batch(gb):
    K = 31
    sizes = gb.get_n_primes_near_x(4, 2_000_000)
    b1, o1 = synth_reads(2000, 150, seed=41)
    b2, o2 = synth_reads(2000, 150, seed=42)
    g1 = make_graph(gb, 0, 1, K, sizes)
    g2 = make_graph(gb, 0, 1, K, sizes)
    g1.insert_sequences(b1, o1)
    g2.insert_sequences(b2, o2)
    g1.S.update_from(g2.S)
    ref = Port(0, 1, K, sizes)
    ref.insert_reads(b1, o1)
    ref.insert_reads(b2, o2)
    assert_tables_equal(g1.get_raw(), ref.tables())
    assert g1.n_occupied() == ref.stats()[1]
    # device-resident packed batch, inserted twice (idempotent for Bit)
    from goetia_b200.batch import PackedBatch
    pb = PackedBatch.from_host(b1, o1)
    g3 = make_graph(gb, 0, 1, K, sizes)
    assert pb.insert_into(g3) == 2000 * 120
    assert pb.insert_into(g3, mode=0) == 2000 * 120
    gb._capi.check(gb._capi.lib().gt_synchronize())
    ref3 = Port(0, 1, K, sizes)
    ref3.insert_reads(b1, o1)
    assert_tables_equal(g3.get_raw(), ref3.tables())


@pytest.mark.parametrize("kind,_n", STORAGES)
def test_sequence_overloads(gb, kind, _n):
    """The other insert_sequence / query_sequence overloads of dbg.hh:249-294, 364-394, replayed k-mer by k-mer
    on the oracle (insert / insert_and_query / query on hash values)."""
    K = 21
    sizes = gb.get_n_primes_near_x(4, 50_021)  # small tables: collisions and repeated k-mers do occur
    g = make_graph(gb, kind, 1, K, sizes)
    ref = Port(kind, 1, K, sizes)
    bases, offsets = genome_reads(12, 90, 300, seed=21)  # overlapping reads: many k-mers repeat
    for r in range(12):
        s = read_str(bases, offsets, r)
        fw, rc = Port.hash_sequence(1, K, s)
        vals = np.minimum(fw, rc)
        if r % 3 == 0:
            n, hs, counts = g.insert_sequence_counts(s)
            assert counts == [int(ref.insert_and_query(int(v))) for v in vals]
        elif r % 3 == 1:
            n, new = g.insert_sequence_new_kmers(s)
            want_new = {int(v) for v in vals if ref.insert(int(v))}
            assert {h.value() for h in new} == want_new
            hs = None
        else:
            n, hs = g.insert_sequence_hashes(s)
            for v in vals:
                ref.insert(int(v))
        assert n == len(s) - K + 1
        if hs is not None:
            assert [h.value() for h in hs] == [int(v) for v in vals]
            assert [(h.fw_hash, h.rc_hash) for h in hs] == [(int(a), int(b)) for a, b in zip(fw, rc)]
    for a, b in zip(g.get_raw(), ref.tables()):
        assert np.array_equal(a, b)
    assert g.n_unique() == ref.stats()[0] and g.n_occupied() == ref.stats()[1]
    q_b, q_o = genome_reads(3, 90, 300, seed=22)
    for r in range(3):
        s = read_str(q_b, q_o, r)
        counts, hs, new = g.query_sequence_hashes(s, want_new=True)
        want = ref.query_sequence(s).tolist()
        assert counts == want
        assert {h.value() for h in new} == {h.value() for h, c in zip(hs, want) if c == 0}
    ref.close()


def test_dev_entry_points_reject_misaligned_pointers(gb):
    """The _dev entry points read the bases with 16 B loads: a misaligned base pointer is an error, not a fault."""
    import torch
    g = make_graph(gb, 0, 1, 21, gb.get_n_primes_near_x(4, 100_003))
    bases = torch.full((4096,), 65, dtype=torch.uint8, device="cuda")
    offs = torch.tensor([0, 100], dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    with pytest.raises(gb.GoetiaB200Error):
        g.insert_sequences_dev(bases.data_ptr() + 4, offs.data_ptr(), 1, 100, mode=0)
    assert g.insert_sequences_dev(bases.data_ptr() + 16, offs.data_ptr(), 1, 100, mode=0) == 80
