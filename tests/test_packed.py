"""The host 2-bit packer (CPU) and the packed insert path (GPU) -- the parsing-to-device pipeline's 0.25 B/base route."""
import numpy as np
import pytest

from tests.util import Port, STORAGES, assert_tables_equal, make_graph, ragged_reads, read_str


def numpy_pack(bases, offsets):
    """plain restatement: code = A0 C1 G2 T3 (either case), base p at bits 2*(p%32) of word p/32; flags = 2 for invalid reads"""
    b = bases[int(offsets[0]):int(offsets[-1])]
    up = b & 0xDF
    code = np.zeros(b.size, dtype=np.uint64)
    for ch, v in ((ord("C"), 1), (ord("G"), 2), (ord("T"), 3)):
        code[up == ch] = v
    valid = (up == ord("A")) | (up == ord("C")) | (up == ord("G")) | (up == ord("T"))
    n_words = (b.size + 31) // 32
    pad = np.zeros(n_words * 32, dtype=np.uint64)
    pad[:b.size] = code
    words = (pad.reshape(n_words, 32) << (np.arange(32, dtype=np.uint64) * np.uint64(2))[None, :]).sum(axis=1, dtype=np.uint64)
    flags = np.zeros(offsets.size - 1, dtype=np.uint8)
    for r in range(offsets.size - 1):
        lo, hi = int(offsets[r] - offsets[0]), int(offsets[r + 1] - offsets[0])
        if not valid[lo:hi].all():
            flags[r] = 2
    return words, flags, valid


def dirty_reads(n, seed):
    bases, offsets = ragged_reads(n, 0, 300, seed=seed, alphabet=b"ACGTacgt")
    bases = bases.copy()
    rng = np.random.default_rng(seed + 1)
    lens = (offsets[1:] - offsets[:-1]).astype(np.int64)
    for r in np.nonzero((rng.random(n) < 0.03) & (lens > 0))[0]:
        bases[int(offsets[r]) + int(rng.integers(0, lens[r]))] = rng.choice(np.frombuffer(b"NnXRY.-*", dtype=np.uint8))
    return bases, offsets


@pytest.mark.parametrize("threads", [1, 4])
def test_host_packer_matches_restatement(threads):
    from goetia_b200.batch import pack_reads_host
    bases, offsets = dirty_reads(3000, seed=5)
    words, flags = pack_reads_host(bases, offsets, n_threads=threads)
    ew, ef, valid = numpy_pack(bases, offsets)
    assert np.array_equal(flags[:ef.size], ef)
    # codes of invalid bytes are unspecified: compare the words with the invalid positions masked out
    n_words = ew.size
    mask = np.zeros(n_words * 32, dtype=np.uint64)
    mask[:valid.size] = np.where(valid, 3, 0)
    m = (mask.reshape(n_words, 32) << (np.arange(32, dtype=np.uint64) * np.uint64(2))[None, :]).sum(axis=1, dtype=np.uint64)
    assert np.array_equal(words[:n_words] & m, ew & m)


def test_host_packer_large_multithreaded_and_offset_start():
    from goetia_b200.batch import pack_reads_host
    rng = np.random.default_rng(9)
    n, L = 40000, 151
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n * L + 77)]
    offsets = np.uint64(77) + np.arange(n + 1, dtype=np.uint64) * np.uint64(L)   # the batch starts mid-buffer
    w1, f1 = pack_reads_host(bases, offsets, n_threads=1)
    w8, f8 = pack_reads_host(bases, offsets, n_threads=8)
    assert np.array_equal(w1, w8) and np.array_equal(f1, f8) and not f1.any()
    ew, _, _ = numpy_pack(bases, offsets)
    assert np.array_equal(w1[:ew.size], ew)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,_n", STORAGES)
@pytest.mark.parametrize("bucket", [False, True])
def test_packed_insert_matches_oracle(gb, monkeypatch, kind, _n, bucket):
    """gt_insert_sequences_packed == the oracle's tables: ragged / short / invalid / lower-case reads, chunks that start in
    the middle of a packed word (conftest sets a small GT_CHUNK_BASES), direct and write-combined path."""
    from goetia_b200.batch import pack_reads_host
    if bucket:
        monkeypatch.setenv("GT_BUCKET_FORCE", "1")
        monkeypatch.setenv("GT_SLICE_LOG2_BYTES", "13")
        monkeypatch.setenv("GT_BUCKET_MIN_KMERS", "0")
        monkeypatch.setenv("GT_PENDING_ENTRIES", str(1 << 22))
    K = 23
    sizes = gb.get_n_primes_near_x(4, 900_001)
    bases, offsets = dirty_reads(9000, seed=21)
    words, flags = pack_reads_host(bases, offsets)
    g = make_graph(gb, kind, 1, K, sizes)
    tot = g.insert_sequences_packed(words, offsets, flags)
    ref = Port(kind, 1, K, sizes)
    upper = np.frombuffer(bases.tobytes().upper(), dtype=np.uint8)
    exp = 0
    for r in range(offsets.size - 1):
        s = read_str(upper, offsets, r)
        if all(c in "ACGT" for c in s) and len(s) >= K:
            exp += ref.insert_sequence(s)[0]
    assert tot == exp
    assert_tables_equal(g.get_raw(), ref.tables())
    assert g.S.pending_info()["built"] == (1 if bucket else 0)
    ref.close()


@pytest.mark.gpu
def test_packed_dev_async_matches_ascii_path(gb):
    import torch
    from goetia_b200.batch import pack_reads_host
    K = 31
    sizes = gb.get_n_primes_near_x(4, 2_000_003)
    bases, offsets = ragged_reads(20000, 40, 200, seed=33)
    words, flags = pack_reads_host(bases, offsets)
    a = make_graph(gb, 0, 1, K, sizes)
    n_ascii = a.insert_sequences(bases, offsets, mode=0)
    b = make_graph(gb, 0, 1, K, sizes)
    dw = torch.from_numpy(np.concatenate([words, np.zeros(4, dtype=np.uint64)]).view(np.int64)).cuda()
    do = torch.from_numpy(offsets.view(np.int64)).cuda()
    df = torch.from_numpy(flags).cuda()
    tot = torch.zeros(1, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    b.insert_packed_dev_async(dw.data_ptr(), dw.numel(), do.data_ptr(), df.data_ptr(), offsets.size - 1, int(offsets[-1]),
                              d_kmer_total_ptr=tot.data_ptr())
    b.flush()
    gb._capi.lib().gt_synchronize()
    assert int(tot.item()) == n_ascii
    assert_tables_equal(b.get_raw(), a.get_raw())
