"""Throughput probe of the write-combining insert path at C3 scale (wall clock around
gt_synchronize; NOT the bench).  PROBE_READS reads per sub-batch, PROBE_BATCHES sub-batches."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import goetia_b200 as gb
from goetia_b200 import _capi

gb.init(0)
L = _capi.lib()
n_reads = int(os.environ.get("PROBE_READS", 6_250_000))
n_batches = int(os.environ.get("PROBE_BATCHES", 4))
kind = int(os.environ.get("PROBE_KIND", 0))
K = int(os.environ.get("PROBE_K", 31))
x = int(float(os.environ.get("PROBE_X", 8e9)))
lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device="cuda")
batches = []
g = torch.Generator(device="cuda"); g.manual_seed(42)
for b in range(n_batches):
    codes = torch.randint(0, 4, (n_reads * 150,), dtype=torch.uint8, device="cuda", generator=g)
    batches.append(lut[codes.long()])
    del codes
offs = (torch.arange(n_reads + 1, dtype=torch.int64, device="cuda") * 150)
torch.cuda.synchronize()
st = [gb.BitStorage, gb.ByteStorage, gb.NibbleStorage][kind](x, 4)
G = gb.dBG[type(st), gb.CanLemireShifter].build(st, K)
for rep in range(4):
    L.gt_synchronize(); t0 = time.perf_counter()
    nk = 0
    for b in batches:
        nk += G.insert_sequences_dev(b.data_ptr(), offs.data_ptr(), n_reads, n_reads * 150, mode=0)
    t1 = time.perf_counter()
    st.flush(); L.gt_synchronize(); t2 = time.perf_counter()
    print("rep %d: %d k-mers  insert calls %.1f ms  final flush %.1f ms  total %.1f ms -> %.2f G k-mers/s (%.0f GB/s algorithmic)"
          % (rep, nk, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t2 - t0) * 1e3, nk / (t2 - t0) / 1e9, nk * 256 / (t2 - t0) / 1e9), flush=True)
print(st.pending_info(), "launches", L.gt_launch_count())
print("n_occupied", st.n_occupied(), "expected ~", "n/a")
