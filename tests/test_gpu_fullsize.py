"""Parity properties at BASELINE.json's full sizes, where the CPU oracle cannot run in test time.

The oracle pins the DIRECT path (k_walk: atomics straight on the tables) bit for bit at small sizes
(tests/test_gpu_dbg.py); here the write-combined path (k_bucket + k_apply) must produce the same tables as the
direct path on the full workloads, compared through gt_storage_checksum (a position-weighted checksum computed in
HBM and itself pinned against the table bytes below), plus idempotence / occupancy / count properties.
"""
import math
import os

import numpy as np
import pytest

from tests.util import Port, make_graph, synth_reads, table_checksum

pytestmark = pytest.mark.gpu


def device_reads(torch, n_reads, length, seed):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device="cuda")
    out = torch.empty(n_reads * length, dtype=torch.uint8, device="cuda")
    step = 1 << 27
    for i in range(0, out.numel(), step):
        m = min(step, out.numel() - i)
        out[i:i + m] = lut[torch.randint(0, 4, (m,), device="cuda", generator=g)]
    torch.cuda.synchronize()
    return out


def insert_all(torch, graph, reads, n_reads, length, per_call):
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


def paths_agree(gb, torch, kind, K, x, n_reads, length, seed, per_call):
    """-> (graph built by the write-combined path, reads) after checking it against the direct path."""
    sizes = gb.get_n_primes_near_x(4, x)
    reads = device_reads(torch, n_reads, length, seed)
    a = make_graph(gb, kind, 1, K, sizes)
    nk = insert_all(torch, a, reads, n_reads, length, per_call)
    assert nk == n_reads * (length - K + 1)
    assert a.S.pending_info()["n_direct"] == 0  # really the bucket path, no overflow
    sums = [a.S.checksum(i) for i in range(4)]
    os.environ["GT_BUCKET"] = "0"  # direct path: k_walk, one atomic per (k-mer, table) on the tables
    try:
        b = make_graph(gb, kind, 1, K, sizes)
        assert insert_all(torch, b, reads, n_reads, length, per_call) == nk
        assert not b.S.pending_info()["built"]
        assert [b.S.checksum(i) for i in range(4)] == sums
        occ_b = b.n_occupied()
    finally:
        os.environ.pop("GT_BUCKET", None)
    assert a.n_occupied() == occ_b
    del b
    return a, reads, sizes, sums


def test_c3_bitstorage_full_size(gb):
    """C3: BitStorage K=31, 4 x 8e9 bits, 50 M x 150 bp reads (6.0e9 k-mers)."""
    import torch
    n_reads, L, K = 50_000_000, 150, 31
    g, reads, sizes, sums = paths_agree(gb, torch, 0, K, int(8e9), n_reads, L, 44, 6_000_000)
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
    g, reads, sizes, sums = paths_agree(gb, torch, 1, K, int(4e9), n_reads, L, 43, 6_000_000)
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
    g, reads, sizes, sums = paths_agree(gb, torch, 2, K, int(8e9), n_reads, L, 46, 60_000)
    sample = reads[:20 * L].cpu().numpy()
    q = g.query_sequences(sample, np.arange(21, dtype=np.uint64) * np.uint64(L))
    assert q.size == 20 * (L - K + 1) and int(q.min()) >= 1 and int(q.max()) <= 15
