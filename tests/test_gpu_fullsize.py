"""Parity at BASELINE.json's full sizes, pinned to the compiled reference.

tests/golden/fullsize_golden.json holds, for every full-size workload, the position-weighted checksum of each table
(gt_storage_checksum's definition) and n_occupied as produced by the UNMODIFIED reference (oracle/_ref) on the
counter-based synthetic reads of goetia_b200/synth.py; tests/golden/make_fullsize_golden.py made it.  Here the same
reads are generated on the device (gt_synth_bases_dev, same arithmetic), inserted through the write-combined path
(k_bucket -> k_rebucket -> k_apply_win) AND through the direct path (k_walk), and both must reproduce the reference's
checksums exactly.  The checksum itself is pinned against table bytes below.  Plus size-independent properties:
idempotence of a second Bloom pass, occupancy, exact doubling of count-min answers, the per-read median query.
"""
import json
import math
import os

import numpy as np
import pytest

from tests.util import Port, make_graph, synth_reads, table_checksum

pytestmark = pytest.mark.gpu

GOLD_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fullsize_golden.json")


def golden(name):
    with open(GOLD_PATH) as f:
        g = json.load(f)
    if name not in g:
        pytest.skip("no golden entry %r yet (tests/golden/make_fullsize_golden.py %s)" % (name, name))
    return g[name]


def device_reads(torch, n_reads, length, seed, first_read=0):
    """The reads the golden file was made from: counter-based stream `seed`, reads [first_read, +n_reads)."""
    from goetia_b200 import _capi
    L = _capi.lib()
    out = torch.empty(n_reads * length + 16, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    _capi.check(L.gt_synth_bases_dev(out.data_ptr(), n_reads * length, seed, first_read * length), "gt_synth_bases_dev")
    L.gt_synchronize()
    return out


def test_device_generator_equals_numpy_twin(gb):
    import torch
    from goetia_b200 import _capi
    from goetia_b200.synth import synth_bases
    L = _capi.lib()
    for seed, first, n in [(44, 0, 100_000), (43, 123_456_789, 70_001), (46, 2**33 + 17, 4099)]:
        d = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        _capi.check(L.gt_synth_bases_dev(d.data_ptr(), n, seed, first), "gt_synth_bases_dev")
        L.gt_synchronize()
        assert np.array_equal(d[:n].cpu().numpy(), synth_bases(seed, first, n))


def insert_all(torch, graph, reads, n_reads, length, per_call):
    per_call -= per_call % 16  # the _dev entry points want 16-byte aligned base pointers
    offs = torch.arange(per_call + 1, dtype=torch.int64, device="cuda") * length
    torch.cuda.synchronize()
    total = 0
    for r0 in range(0, n_reads, per_call):
        n = min(per_call, n_reads - r0)
        total += graph.insert_sequences_dev(reads[r0 * length:].data_ptr(), offs.data_ptr(), n, n * length, mode=0)
    graph.flush()
    return total


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_checksum_is_the_checksum_of_the_bytes(gb, kind):
    K = 21
    sizes = gb.get_n_primes_near_x(4, 3_000_017)
    g = make_graph(gb, kind, 1, K, sizes)
    bases, offsets = synth_reads(3000, 100, seed=3)
    g.insert_sequences(bases, offsets, mode=0)
    ref = Port(kind, 1, K, sizes)
    ref.insert_reads(bases, offsets)
    for i, (a, b) in enumerate(zip(g.get_raw(), ref.tables())):
        assert np.array_equal(a, b)
        assert g.S.checksum(i) == table_checksum(b)
    ref.close()


def both_paths_equal_reference(gb, torch, name, per_call):
    """-> (graph built by the write-combined path, reads, golden entry) after checking BOTH insert paths against the
    reference's checksums."""
    gold = golden(name)
    kind, K, L, n_reads, seed = gold["kind"], gold["K"], gold["read_len"], gold["reads"], gold["seed"]
    sizes = gb.get_n_primes_near_x(4, gold["x"])
    assert [int(x) for x in sizes] == gold["tablesizes"]
    reads = device_reads(torch, n_reads, L, seed, gold["first_read"])
    a = make_graph(gb, kind, 1, K, sizes)
    nk = insert_all(torch, a, reads, n_reads, L, per_call)
    assert nk == gold["kmers"]
    info = a.S.pending_info()
    assert info["built"] == 1 and info["n_direct"] == 0  # really the bucket path, no overflow
    assert [a.S.checksum(i) for i in range(4)] == gold["checksums"], "write-combined path differs from the reference"
    assert a.n_occupied() == gold["n_occupied"]
    os.environ["GT_BUCKET"] = "0"  # direct path: k_walk, one atomic per (k-mer, table) on the tables
    try:
        b = make_graph(gb, kind, 1, K, sizes)
        assert insert_all(torch, b, reads, n_reads, L, per_call) == nk
        assert not b.S.pending_info()["built"]
        assert [b.S.checksum(i) for i in range(4)] == gold["checksums"], "direct path differs from the reference"
    finally:
        os.environ.pop("GT_BUCKET", None)
    del b
    return a, reads, gold


@pytest.mark.parametrize("name", ["c1", "c3_first_1m", "c2_first_1m"])
def test_prefix_workloads_equal_reference(gb, name):
    import torch
    both_paths_equal_reference(gb, torch, name, 400_000)


def test_c3_bitstorage_full_size(gb):
    """C3: BitStorage K=31, 4 x 8e9 bits, 50 M x 150 bp reads (6.0e9 k-mers)."""
    import torch
    n_reads, L, K = 50_000_000, 150, 31
    g, reads, gold = both_paths_equal_reference(gb, torch, "c3", 6_000_000)
    sizes, sums = gold["tablesizes"], gold["checksums"]
    # idempotence: a second pass over the same reads leaves a Bloom table unchanged
    insert_all(torch, g, reads, n_reads, L, 6_000_000)
    assert [g.S.checksum(i) for i in range(4)] == sums
    # occupancy of table 0 after n distinct uniform hashes: size * (1 - exp(-n/size)) (n ~ 6e9 canonical 31-mers)
    n = n_reads * (L - K + 1)
    expect = sizes[0] * (1.0 - math.exp(-n / sizes[0]))
    assert abs(g.n_occupied() - expect) / expect < 2e-3
    # every k-mer of a sample of the reads is present
    sample = reads[:2000 * L].cpu().numpy()
    q = g.query_sequences(sample, np.arange(2001, dtype=np.uint64) * np.uint64(L))
    assert q.size == 2000 * (L - K + 1) and bool((q == 1).all())


def test_c2_bytestorage_full_size(gb):
    """C2: ByteStorage K=21 count-min insert over 20 M x 150 bp reads, then the per-read median-count query."""
    import torch
    n_reads, L, K = 20_000_000, 150, 21
    g, reads, gold = both_paths_equal_reference(gb, torch, "c2", 6_000_000)
    # counters are counts: the byte sum of every table equals the k-mers inserted (nothing near saturation here)
    sample_n = 200_000
    sample = reads[:sample_n * L].cpu().numpy()
    so = np.arange(sample_n + 1, dtype=np.uint64) * np.uint64(L)
    assert bool(g.median_count_at_least(sample, so, 1).all())       # every k-mer was counted at least once
    assert int(g.median_count_at_least(sample, so, 3).sum()) == 0   # uniform random reads: medians stay at 1-2
    q = g.query_sequences(sample[:1000 * L], so[:1001])
    assert int(q.min()) >= 1 and int(q.max()) <= 6
    # a second pass doubles every count: the query of the sample moves up by exactly its first-pass value
    insert_all(torch, g, reads, n_reads, L, 6_000_000)
    q2 = g.query_sequences(sample[:1000 * L], so[:1001])
    assert np.array_equal(q2, 2 * q)


def test_c5_nibblestorage_long_reads(gb):
    """C5 shape at one GPU's share: NibbleStorage K=25, 10 kb reads (the long-sequence rolling-hash path)."""
    import torch
    n_reads, L, K = 125_000, 10_000, 25  # 1/8 of C5's 1 M reads = what one of 8 GPUs hashes
    g, reads, gold = both_paths_equal_reference(gb, torch, "c5_first_125k", 60_000)
    sample = reads[:20 * L].cpu().numpy()
    q = g.query_sequences(sample, np.arange(21, dtype=np.uint64) * np.uint64(L))
    assert q.size == 20 * (L - K + 1) and int(q.min()) >= 1 and int(q.max()) <= 15
