"""Generate tests/golden/fastx_golden.json: what the UNMODIFIED reference parser
(FastxParser<DNA_SIMPLE>, compiled into oracle/_ref) returns for tests/fastx_cases.py.

Run only where /root/reference exists:   python tests/golden/make_fastx_golden.py
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.binding import Port, Ref, build_ref  # noqa: E402
from tests.fastx_cases import cases, write_case  # noqa: E402


def main():
    build_ref()
    G = {"_generator": "tests/golden/make_fastx_golden.py: Ref.parse_file (FastxParser<DNA_SIMPLE>::next loop) on tests/fastx_cases.py"}
    with tempfile.TemporaryDirectory() as td:
        for name, data, min_length, strict in cases():
            for gz in (False, True):
                fn = write_case(td, name, data, gz)
                key = os.path.basename(fn)
                try:
                    bb, oo, sk = Ref.parse_file(fn, strict=strict, min_length=min_length)
                except ValueError as e:
                    G[key] = {"error": str(e), "min_length": min_length, "strict": strict}
                    continue
                lens = (oo[1:] - oo[:-1]).astype(np.uint64)
                G[key] = {"min_length": min_length, "strict": strict, "n_reads": int(oo.size - 1), "n_skipped": sk,
                          "n_bases": int(bb.size), "seq_fnv": str(Port.fnv1a(bb)),
                          "len_fnv": str(Port.fnv1a(lens.view(np.uint8))),
                          "first": bb[:int(oo[1])].tobytes().decode()[:80] if oo.size > 1 else ""}
    out = os.path.join(ROOT, "tests", "golden", "fastx_golden.json")
    with open(out, "w") as f:
        json.dump(G, f, indent=1)
    print("wrote", out, len(G) - 1, "cases")


if __name__ == "__main__":
    main()
