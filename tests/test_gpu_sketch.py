"""SourmashSketch on the GPU vs the CPU restatement (oracle PortSketch) -- hash sets bit-exact.

Reference behaviour: sketches/sourmash_sketch.hh:24-82 (+ libsourmash 3.4.0 add_sequence, restated in
oracle/goetia_oracle.c).  Known answers: MurmurHash3 of the vendored smhasher (hash_murmur("ACG")) and
upstream sourmash's own tests/test__minhash.py::test_basic_dna value for MinHash(1, 4).add_sequence("ATGC").
"""
import numpy as np
import pytest

from oracle.binding import Port, PortSketch
from tests.util import ragged_reads, read_str, synth_reads

pytestmark = pytest.mark.gpu


def _sk(gb, num, K, scaled, seed=42):
    return gb.SourmashSketch.Sketch(num, K, False, False, False, seed, scaled)


def test_known_answers(gb):
    sk = _sk(gb, 1, 4, 0)
    sk.add_sequence("ATGC")
    assert sk.mins().tolist() == [12415348535738636339]  # sourmash tests/test__minhash.py::test_basic_dna
    sk.add_sequence("GCAT")  # same canonical k-mer: nothing new
    assert sk.mins().tolist() == [12415348535738636339] and sk.size() == 1
    sk3 = _sk(gb, 0, 3, 1)
    sk3.add_sequence("ACG")  # min("ACG", "CGT") = "ACG"
    assert sk3.mins().tolist() == [1731421407650554201]  # hash_murmur("ACG"), SURVEY.md section 8c
    assert gb.SourmashSketch.Sketch.max_hash_from_scaled(1000) == 18446744073709552
    assert gb.SourmashSketch.Sketch.max_hash_from_scaled(0) == 0
    assert gb.SourmashSketch.Sketch.max_hash_from_scaled(1) == 2**64 - 1
    assert gb.SourmashSketch.Sketch.scaled_from_max_hash(18446744073709552) in (999, 1000)


@pytest.mark.parametrize("K", [1, 4, 15, 16, 17, 21, 31, 32, 33, 47, 48, 49, 63, 64])
def test_scaled_sketch_all_K(gb, K):
    bases, offsets = ragged_reads(300, 1, 400, seed=100 + K)
    scaled = 1 if K < 4 else 7
    g = _sk(gb, 0, K, scaled)
    o = PortSketch(0, K, 42, scaled=scaled)
    n = g.insert_sequences(bases, offsets)
    n_ref, _ = o.add_reads(bases, offsets)
    assert n == n_ref
    assert np.array_equal(g.mins(), o.mins())
    assert g.size() == o.mins().size


def test_scaled_1000_streaming_batches(gb):
    """C4's parameters (K=31, scaled=1000) over several calls; the set only grows and ends bit-exact."""
    g = _sk(gb, 0, 31, 1000)
    o = PortSketch(0, 31, 42, scaled=1000)
    prev = 0
    for i in range(4):
        bases, offsets = synth_reads(20000, 150, seed=42 + i)
        assert g.insert_sequences(bases, offsets) == o.add_reads(bases, offsets)[0] == 20000 * 120
        assert g.size() >= prev
        prev = g.size()
    m = g.mins()
    assert np.array_equal(m, o.mins())
    assert m.size > 5000 and int(m[-1]) <= 18446744073709552 and np.all(m[1:] > m[:-1])


def test_set_growth_beyond_initial_capacity(gb):
    """scaled=2 keeps half of all k-mers: the device hash set has to grow and replay (no loss)."""
    bases, offsets = synth_reads(12000, 150, seed=9)
    g = _sk(gb, 0, 21, 2)
    o = PortSketch(0, 21, 42, scaled=2)
    g.insert_sequences(bases, offsets)
    o.add_reads(bases, offsets)
    assert np.array_equal(g.mins(), o.mins())
    assert g.size() > (1 << 19)


def test_non_acgt_windows_are_skipped_and_case_folded(gb):
    bases, offsets = ragged_reads(400, 20, 300, seed=5, alphabet=b"ACGTacgtNnRY")
    g = _sk(gb, 0, 21, 3)
    o = PortSketch(0, 21, 42, scaled=3)
    assert g.insert_sequences(bases, offsets) == o.add_reads(bases, offsets)[0]
    assert np.array_equal(g.mins(), o.mins())
    assert g.size() > 0
    with pytest.raises(ValueError):
        g.add_sequence("ACGTNACGTACGTACGTACGTACGTACGT", force=False)


@pytest.mark.parametrize("num", [1, 50, 500, 10000])
def test_bottom_k(gb, num):
    bases, offsets = synth_reads(3000, 150, seed=11)
    g = _sk(gb, num, 31, 0)
    o = PortSketch(num, 31, 42, scaled=0)
    for lo in range(0, 3000, 1000):  # several calls: the running num-th smallest tightens the admission bound
        b = bases[int(offsets[lo]):int(offsets[lo + 1000])]
        off = offsets[lo:lo + 1001] - offsets[lo]
        g.insert_sequences(b, off)
        o.add_reads(b, off)
    assert np.array_equal(g.mins(), o.mins())
    assert g.size() == num


def test_strand_symmetry_merge_common_and_add_hash(gb):
    bases, offsets = synth_reads(40, 500, seed=3)
    comp = str.maketrans("ACGT", "TGCA")
    a, b = _sk(gb, 0, 31, 5), _sk(gb, 0, 31, 5)
    for r in range(40):
        s = read_str(bases, offsets, r)
        assert a.insert_sequence(s) == len(s) - 31 + 1
        b.insert_sequence(s.translate(comp)[::-1])
    assert np.array_equal(a.mins(), b.mins())
    # merge / count_common / jaccard against set arithmetic
    x, y = _sk(gb, 0, 31, 5), _sk(gb, 0, 31, 5)
    x.insert_sequences(bases[:int(offsets[25])], offsets[:26])
    y.insert_sequences(bases[int(offsets[15]):], offsets[15:] - offsets[15])
    sx, sy = set(x.mins().tolist()), set(y.mins().tolist())
    assert x.count_common(y) == len(sx & sy)
    assert abs(x.jaccard(y) - len(sx & sy) / len(sx | sy)) < 1e-12
    x.merge(y)
    assert np.array_equal(x.mins(), a.mins())
    # add_hash honours max_hash; duplicates collapse
    z = _sk(gb, 0, 31, 5)
    z.add_hashes([5, 5, 7, 2**64 - 1, z.max_hash(), z.max_hash() + 1])
    assert z.mins().tolist() == [5, 7, z.max_hash()]
    z.reset()
    assert z.size() == 0
    c = a.copy()
    assert np.array_equal(c.mins(), a.mins())


def test_device_resident_reads(gb):
    import torch
    bases, offsets = ragged_reads(5000, 10, 300, seed=21, alphabet=b"ACGTN")
    d_b = torch.from_numpy(bases).cuda()
    d_o = torch.from_numpy(offsets.view(np.int64)).cuda()
    g = _sk(gb, 0, 31, 10)
    o = PortSketch(0, 31, 42, scaled=10)
    n = g.insert_sequences_dev(d_b.data_ptr(), d_o.data_ptr(), offsets.size - 1, int(offsets[-1]))
    assert n == o.add_reads(bases, offsets)[0]
    assert np.array_equal(g.mins(), o.mins())


def test_full_size_properties(gb):
    """Size-independent checks at a size the oracle does not run in test time: idempotence (a second
    pass adds nothing), strand symmetry of a whole batch, and the expected kept fraction 1/scaled."""
    import torch
    n_reads, L, K = 400_000, 150, 31
    gen = torch.Generator(device="cuda")
    gen.manual_seed(42)
    codes = torch.randint(0, 4, (n_reads * L,), device="cuda", generator=gen)
    lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device="cuda")
    d_b = lut[codes]
    d_o = (torch.arange(n_reads + 1, dtype=torch.int64, device="cuda") * L)
    torch.cuda.synchronize()  # the library runs on its own streams: inputs must be complete before the call
    g = _sk(gb, 0, K, 1000)
    nk = g.insert_sequences_dev(d_b.data_ptr(), d_o.data_ptr(), n_reads, n_reads * L)
    assert nk == n_reads * (L - K + 1)
    m1 = g.mins()
    g.insert_sequences_dev(d_b.data_ptr(), d_o.data_ptr(), n_reads, n_reads * L)
    assert np.array_equal(g.mins(), m1)
    expect = nk / 1000.0
    assert abs(m1.size - expect) < 6 * expect ** 0.5
    # reverse complement of the whole stream = the same set of canonical k-mers except the windows that
    # straddle read boundaries; use one long read to make the sets identical
    one = torch.tensor([0, n_reads * L], dtype=torch.int64, device="cuda")
    a, b = _sk(gb, 0, K, 1000), _sk(gb, 0, K, 1000)
    a.insert_sequences_dev(d_b.data_ptr(), one.data_ptr(), 1, n_reads * L)
    rc = lut[(3 - codes).flip(0)]
    torch.cuda.synchronize()
    b.insert_sequences_dev(rc.data_ptr(), one.data_ptr(), 1, n_reads * L)
    assert np.array_equal(a.mins(), b.mins())
    # spot-check membership of a prefix against the oracle
    o = PortSketch(0, K, 42, scaled=1000)
    hb = d_b[:20000 * L].cpu().numpy()
    o.add_reads(hb, np.arange(20001, dtype=np.uint64) * np.uint64(L))
    assert set(o.mins().tolist()) <= set(m1.tolist())


def test_stream_fastx_intervals(gb, tmp_path):
    """The streaming loop of `goetia sketch sourmash` (signature_runner.py:131-157): snapshots every `interval`
    k-mers; at every snapshot the hash set equals the oracle's after the same prefix of reads, and the reported
    similarity is the Jaccard of consecutive oracle snapshots."""
    import os
    from tests.util import genome_reads, read_str
    K, n_reads, L, interval = 21, 3000, 100, 40_000
    bases, offsets = genome_reads(n_reads, L, 60_000, seed=8)
    fn = os.path.join(str(tmp_path), "s.fa")
    with open(fn, "w") as f:
        for r in range(n_reads):
            f.write(">r%d\n%s\n" % (r, read_str(bases, offsets, r)))
    g = _sk(gb, 0, K, 50)
    snaps = list(g.stream_fastx(fn, interval=interval))
    per = L - K + 1
    assert snaps[-1]["t"] == n_reads * per and snaps[-1]["sequences"] == n_reads
    assert len(snaps) == -(-n_reads // -(-interval // per))  # every interval closes at the read that fills it
    o = PortSketch(0, K, 42, scaled=50)
    done, prev = 0, None
    for s in snaps:
        n = s["sequences"]
        assert s["t"] == n * per
        o.add_reads(bases[int(offsets[done]):int(offsets[n])], offsets[done:n + 1] - offsets[done])
        done = n
        cur = set(o.mins().tolist())
        assert s["size"] == len(cur)
        if prev is None:
            assert s["similarity"] is None
        else:
            assert abs(s["similarity"] - len(cur & prev) / len(cur | prev)) < 1e-12
        prev = cur
    assert np.array_equal(g.mins(), o.mins())
