"""Generate tests/golden/advance_golden.json: every return value <n_sequences, time_total, remaining> of the UNMODIFIED
reference's FileProcessor<InserterProcessor<dBG>>::advance (processors.hh:208-229, compiled into oracle/_ref) called
until nothing remains, over tests/fastx_cases.py files at several intervals; and the tables' FNV afterwards.

Run only where /root/reference exists:   python tests/golden/make_advance_golden.py
"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.binding import Port, Ref, build_ref  # noqa: E402
from tests.fastx_cases import cases, write_case  # noqa: E402

K = 21
SIZES = [999983, 999979, 999961, 999959]
# (case, interval): small files at small intervals (every record boundary matters), the large one at realistic ones
PLAN = [("reads.fq", 1), ("reads.fq", 97), ("reads.fq", 1000), ("reads_min50.fq", 250), ("multiline_crlf.fa", 300),
        ("wrapped.fq", 777), ("truncated.fq", 5), ("strict.fa", 1), ("simple.fa", 3), ("junk.fa", 2), ("empty.fa", 10),
        ("big.fq", 100000), ("big.fq", 500000), ("big.fq", 1300013)]


def main():
    build_ref()
    G = {"_generator": "tests/golden/make_advance_golden.py: Ref.advance_trace (FileProcessor::advance loop), dBG<BitStorage, "
                       "CanLemireShifter> K=%d" % K, "K": K, "sizes": SIZES, "cases": []}
    by_name = {c[0]: c for c in cases()}
    with tempfile.TemporaryDirectory() as td:
        for name, interval in PLAN:
            _, data, min_length, strict = by_name[name]
            for gz in (False, True):
                if gz and name == "big.fq" and interval != 100000:
                    continue
                fn = write_case(td, name, data, gz)
                r = Ref(0, 1, K, SIZES)
                trace = r.advance_trace(fn, interval, strict=strict, min_length=min_length)
                G["cases"].append({"file": os.path.basename(fn), "interval": interval, "strict": strict, "min_length": min_length,
                                   "trace": [[a, b, int(c)] for a, b, c in trace],
                                   "table_fnv": [str(Port.fnv1a(t)) for t in r.tables()]})
                r.close()
    out = os.path.join(ROOT, "tests", "golden", "advance_golden.json")
    with open(out, "w") as f:
        json.dump(G, f, separators=(",", ":"))
    print("wrote", out, len(G["cases"]), "cases,", os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
