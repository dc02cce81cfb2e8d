"""InserterProcessor / FileProcessor::advance (goetia_b200/processors.py, gt_insert_fastx_advance) against the UNMODIFIED
reference's advance loop (processors.hh:208-229; IntervalCounter::poll, metrics.hh:129-138).

tests/golden/advance_golden.json holds every <n_sequences, time_total, remaining> the compiled reference returned over
the files of tests/fastx_cases.py at several k-mer intervals, and the tables' FNV afterwards
(tests/golden/make_advance_golden.py)."""
import json
import os

import numpy as np
import pytest

from oracle.binding import Port
from tests.fastx_cases import cases, write_case

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "advance_golden.json")))
CASES = {c[0]: c for c in cases()}


@pytest.mark.parametrize("idx", range(len(GOLD["cases"])), ids=lambda i: "%s@%d" % (GOLD["cases"][i]["file"], GOLD["cases"][i]["interval"]))
def test_advance_matches_reference(tmp_path, gb, idx):
    from goetia_b200.parsing import FastxParser
    c = GOLD["cases"][idx]
    gz = c["file"].endswith(".gz")
    name = c["file"][:-3] if gz else c["file"]
    fn = write_case(str(tmp_path), name, CASES[name][1], gz=gz)
    g = gb.dBG[gb.BitStorage, gb.CanLemireShifter].build(gb.BitStorage(GOLD["sizes"]), GOLD["K"])
    proc = gb.InserterProcessor.build(g, c["interval"])
    parser = FastxParser.build(fn, c["strict"], c["min_length"])
    trace, remaining = [], True
    while remaining:
        n, t, remaining = proc.advance(parser)
        trace.append([n, t, int(remaining)])
        assert len(trace) <= len(c["trace"]) + 1
    assert trace == c["trace"]
    assert (proc.n_sequences(), proc.time_elapsed()) == tuple(c["trace"][-1][:2])
    assert [str(Port.fnv1a(t)) for t in g.get_raw()] == c["table_fnv"]
    parser.close()


@pytest.mark.parametrize("by_interval", [False, True])
def test_process_whole_file(tmp_path, gb, by_interval, monkeypatch):
    """process(): same totals and tables whichever way the file is consumed; the uncompressed file goes through the
    parser workers' 2-bit packed batches (several of them), the gz one through the ASCII batches."""
    want = [c for c in GOLD["cases"] if c["file"] == "big.fq" and c["interval"] == 100000][0]
    monkeypatch.setenv("GT_FASTX_BATCH_BASES", str(1 << 20))
    monkeypatch.setenv("GT_FASTX_CHUNK_BYTES", str(256 << 10))
    monkeypatch.setenv("GT_FASTX_THREADS", "4")
    for gz in (False, True):
        fn = write_case(str(tmp_path), "big.fq", CASES["big.fq"][1], gz=gz)
        g = gb.dBG[gb.BitStorage, gb.CanLemireShifter].build(gb.BitStorage(GOLD["sizes"]), GOLD["K"])
        proc = gb.InserterProcessor.build(g, 100000)
        assert proc.process(fn, by_interval=by_interval) == tuple(want["trace"][-1][:2])
        assert [str(Port.fnv1a(t)) for t in g.get_raw()] == want["table_fnv"]
        # a second file on the same processor: the totals are cumulative (FileProcessor keeps _n_sequences and the timer)
        n2, t2 = proc.process(fn, by_interval=by_interval)
        assert (n2, t2) == (2 * want["trace"][-1][0], 2 * want["trace"][-1][1])


def test_process_skips_bad_records(tmp_path, gb):
    """handle_next (processors.hh:148-172) swallows InvalidRead and, under a strict parser, InvalidCharacterException: the
    run goes on with the next record."""
    fn = os.path.join(str(tmp_path), "bad.fq")
    good = "ACGTTGCATGCCGATAGCTAGCTAGGATCGA"
    with open(fn, "w") as f:
        f.write("@a\n%s\n+\n%s\n@b\n%s\n+\nIII\n@c\n%sNN\n+\n%s\n@d\n%s\n+\n%s\n" % (good, "I" * len(good), good, good, "I" * (len(good) + 2),
                                                                                 good[::-1], "I" * len(good)))
    K = 21
    for strict in (False, True):
        g = gb.dBG[gb.BitStorage, gb.CanLemireShifter].build(gb.BitStorage(GOLD["sizes"]), K)
        proc = gb.InserterProcessor.build(g, 1000)
        n, t = proc.process(fn, strict=strict)
        assert (n, t) == (2, 2 * (len(good) - K + 1))
        ref = Port(0, 1, K, GOLD["sizes"])
        for s in (good, good[::-1]):
            ref.insert_reads(np.frombuffer(s.encode(), dtype=np.uint8), np.array([0, len(s)], dtype=np.uint64))
        for a, b in zip(g.get_raw(), ref.tables()):
            assert np.array_equal(a, b)
        ref.close()
