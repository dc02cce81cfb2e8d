"""DiginormFilter / StreamingSolidFilter / FilterProcessor on the GPU dBG vs the oracle.

Default = the reference's SERIAL semantics (diginorm.hh:111-119 judges and inserts read by read; dbg.hh:327-340 returns
every k-mer's count after its own insert): whatever the batch size the GPU path is given, the result must equal the
oracle's one-read-at-a-time loop.  ``batch_synchronous=True`` is the documented one-round approximation (SURVEY.md
section 8a) and is compared with the oracle run under the same batch size."""
import os

import numpy as np
import pytest

from oracle.binding import Port
from tests.util import genome_reads, make_graph, ragged_reads, read_str

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,K", [(1, 21), (2, 25)])
@pytest.mark.parametrize("batch", [1, 64, 0])
@pytest.mark.parametrize("sync", [False, True])
def test_diginorm_matches_oracle(gb, kind, K, batch, sync):
    sizes = gb.get_n_primes_near_x(4, 2_000_003)
    n_reads = 300 if batch == 1 else 3000
    bases, offsets = genome_reads(n_reads, 100, 3000, seed=77)  # ~100x coverage: most late reads are filtered
    cutoff = 5
    g = make_graph(gb, kind, 1, K, sizes)
    f = gb.DiginormFilter.build(g, cutoff, batch_synchronous=sync)
    ref = Port(kind, 1, K, sizes)
    B = batch or n_reads
    # serial (default): the oracle's one-at-a-time loop, whatever batches the GPU path gets
    want_keep = ref.diginorm_reads(bases, offsets, cutoff, batch=B if sync else 1)
    got_keep = np.zeros(n_reads, dtype=np.uint8)
    judged = 0
    for r0 in range(0, n_reads, B):
        r1 = min(n_reads, r0 + B)
        b = bases[int(offsets[r0]):int(offsets[r1])]
        o = offsets[r0:r1 + 1] - offsets[r0]
        k, nk = f.filter_sequences(b, o)
        got_keep[r0:r1] = k
        judged += nk
    assert judged == n_reads * (100 - K + 1)
    assert np.array_equal(got_keep, want_keep[:n_reads])
    if batch or not sync:  # (one synchronous batch over everything judges every read against the empty table: all kept)
        assert 0 < got_keep.sum() < n_reads
    for a, b in zip(g.get_raw(), ref.tables()):
        assert np.array_equal(a, b)
    ref.close()


@pytest.mark.parametrize("kind,K", [(0, 31), (1, 21), (2, 25)])
def test_insert_and_query_sequences_serial_counts(gb, kind, K):
    """gt_insert_and_query_sequences == the oracle's insert_and_query_sequence read after read: overlapping reads, exact
    duplicates, poly-A reads (one slot hit > 100 times in a row: counters saturate inside the batch), short and
    invalid reads."""
    sizes = gb.get_n_primes_near_x(4, 300_007)
    b1, o1 = genome_reads(300, 90, 1500, seed=5)
    reads = [read_str(b1, o1, r) for r in range(300)]
    reads[40:40] = ["A" * 200, "ACGT" * 30, "A" * 180, "ACG", "ACGTNACGT" * 12, reads[3], reads[3]]
    reads += ["T" * 400] * 2  # the reverse complement of poly-A: same canonical k-mer
    bases, offsets = gb._capi.reads_from_strings(reads)
    g = make_graph(gb, kind, 1, K, sizes)
    got, status = g.insert_and_query_sequences(bases, offsets, want_status=True)
    ref = Port(kind, 1, K, sizes)
    want = []
    for s in reads:
        if len(s) >= K and all(c in "ACGT" for c in s):
            want.append(ref.insert_and_query_sequence(s))
    want = np.concatenate(want)
    assert got.size == want.size
    assert np.array_equal(got, want)
    if kind:
        assert int(got.max()) == (255 if kind == 1 else 15)
    for a, b in zip(g.get_raw(), ref.tables()):
        assert np.array_equal(a, b)
    # the single-sequence member
    s = reads[10]
    assert g.insert_and_query_sequence(s) == [int(c) for c in ref.insert_and_query_sequence(s)]
    ref.close()


def test_diginorm_ragged_and_invalid(gb):
    """Short and invalid reads are dropped (the processor swallows their exceptions, processors.hh:389-417)."""
    kind, K, cutoff = 1, 21, 2
    sizes = gb.get_n_primes_near_x(4, 500_009)
    bases, offsets = ragged_reads(400, 5, 120, seed=3, alphabet=b"ACGTACGTACGTACGTN")
    g = make_graph(gb, kind, 1, K, sizes)
    f = gb.DiginormFilter.build(g, cutoff)
    ref = Port(kind, 1, K, sizes)
    for _ in range(3):  # the same reads again: by the third pass everything long enough is filtered
        want = ref.diginorm_reads(bases, offsets, cutoff, batch=1)
        got, _ = f.filter_sequences(bases, offsets)
        assert np.array_equal(got, want[:400])
    for a, b in zip(g.get_raw(), ref.tables()):
        assert np.array_equal(a, b)
    # single-sequence member
    s = read_str(bases, offsets, int(np.argmax(offsets[1:] - offsets[:-1])))
    if "N" not in s:
        passed, nk = f.filter_sequence(s)
        assert nk == len(s) - K + 1 and passed is False
    ref.close()


def test_filter_processor_end_to_end(gb, tmp_path):
    """FASTQ in -> FilterProcessor<DiginormFilter> -> FASTQ out holding exactly the reads the oracle keeps."""
    kind, K, cutoff, n_reads = 1, 21, 4, 2000
    sizes = gb.get_n_primes_near_x(4, 1_000_003)
    bases, offsets = genome_reads(n_reads, 80, 2000, seed=11)
    fn = os.path.join(str(tmp_path), "in.fq")
    with open(fn, "w") as fh:
        for r in range(n_reads):
            fh.write("@r%d\n%s\n+\n%s\n" % (r, read_str(bases, offsets, r), "I" * 80))
    g = make_graph(gb, kind, 1, K, sizes)
    out = os.path.join(str(tmp_path), "out.fq")
    # default batch size (100000 reads > the whole input): the output must still be the reference's serial output
    proc = gb.FilterProcessor.build(gb.DiginormFilter.build(g, cutoff), out)
    n_seqs, time = proc.process(fn)
    assert n_seqs == n_reads and time == n_reads * (80 - K + 1)
    ref = Port(kind, 1, K, sizes)
    want = ref.diginorm_reads(bases, offsets, cutoff, batch=1)[:n_reads]
    assert 0 < int(want.sum()) < n_reads
    names = [ln[1:].strip() for ln in open(out) if ln.startswith("@r")]
    assert names == ["r%d" % r for r in range(n_reads) if want[r]]
    assert proc.n_passed == int(want.sum())
    ref.close()


def test_streaming_solid_filter(gb):
    """StreamingSolidFilter::Filter::filter_sequence (solidifier.hh:58-76) replayed on the oracle's
    insert_and_query_sequence."""
    kind, K = 1, 21
    sizes = gb.get_n_primes_near_x(4, 200_003)
    bases, offsets = genome_reads(60, 70, 400, seed=4)  # overlapping reads: later ones are mostly solid
    g = make_graph(gb, kind, 1, K, sizes)
    f = gb.StreamingSolidFilter.build(g, 0.6, 2)
    ref = Port(kind, 1, K, sizes)
    want = []
    for r in range(60):
        s = read_str(bases, offsets, r)
        counts = ref.insert_and_query_sequence(s)
        n_not = int((counts < 2).sum())
        want.append(0 if float(np.float32(n_not) / np.float32(len(s) - K + 1)) >= 1.0 - float(np.float32(0.6)) else 1)
    keep, judged = f.filter_sequences(bases, offsets)
    assert keep.tolist() == want and 0 < sum(want) < 60
    assert judged == 60 * (70 - K + 1)
    for a, b in zip(g.get_raw(), ref.tables()):
        assert np.array_equal(a, b)
    ref.close()
