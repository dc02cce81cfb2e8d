"""Quick throughput probe (wall-clock around gt_synchronize; NOT the bench -- see bench.py)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import goetia_b200 as gb
from goetia_b200.batch import PackedBatch
from goetia_b200 import _capi
from oracle.binding import synth_reads

gb.init(0)
L = _capi.lib()
n_reads = int(os.environ.get("PROBE_READS", 4_000_000))
t = time.time(); bases, offsets = synth_reads(n_reads, 150, 42); print("synth %.1fs" % (time.time() - t), flush=True)
t = time.time(); pb = PackedBatch.from_host(bases, offsets); print("pack+h2d (pageable) %.3fs" % (time.time() - t), flush=True)

def run(name, kind, K, x, mode, can=1, reps=3):
    st = [gb.BitStorage, gb.ByteStorage, gb.NibbleStorage][kind](x, 4)
    g = gb.dBG[type(st), [gb.FwdLemireShifter, gb.CanLemireShifter][can]].build(st, K)
    nk = pb.n_kmers(K)
    best = 1e9
    for r in range(reps):
        st.reset()
        L.gt_synchronize(); t0 = time.perf_counter()
        pb.insert_into(g, mode=mode)
        L.gt_synchronize(); dt = time.perf_counter() - t0
        best = min(best, dt)
    print("%-34s %8.2f ms  %7.2f G k-mers/s  (%.0f GB/s algorithmic)" % (name, best * 1e3, nk / best / 1e9, nk * 256 / best / 1e9), flush=True)
    # second pass over a full table (all bits already set)
    L.gt_synchronize(); t0 = time.perf_counter(); pb.insert_into(g, mode=mode); L.gt_synchronize(); dt = time.perf_counter() - t0
    print("%-34s %8.2f ms  %7.2f G k-mers/s  (2nd pass)" % (name, dt * 1e3, nk / dt / 1e9), flush=True)
    return g

for x in (int(1e9), int(8e9)):
    for mode, mn in ((0, "blind"), (1, "fast")):
        run("Bit K=31 x=%.0e %s" % (x, mn), 0, 31, x, mode)
run("Bit K=31 x=8e9 blind FWD", 0, 31, int(8e9), 0, can=0)
for mode, mn in ((0, "blind"), (1, "fast")):
    run("Byte K=21 x=4e9 %s" % mn, 1, 21, int(4e9), mode)
    run("Nibble K=25 x=8e9 %s" % mn, 2, 25, int(8e9), mode)

# host-API end to end (pinned)
import torch
pin = torch.empty(bases.size, dtype=torch.uint8, pin_memory=True); pin.numpy()[:] = bases
pino = torch.empty(offsets.size, dtype=torch.int64, pin_memory=True); pino.numpy()[:] = offsets.astype(np.int64)
os.environ["GT_CHUNK_BASES"] = str(64 << 20)
st = gb.BitStorage(int(8e9), 4); g = gb.dBG[gb.BitStorage, gb.CanLemireShifter].build(st, 31)
for r in range(3):
    st.reset(); t0 = time.perf_counter()
    nk = g.insert_sequences(pin.numpy(), pino.numpy().view(np.uint64), mode=0)
    dt = time.perf_counter() - t0
    print("e2e host API Bit K=31 blind: %.2f ms %.2f G k-mers/s  h2d=%.1f GB/s-equivalent" % (dt * 1e3, nk / dt / 1e9, bases.size / dt / 1e9), flush=True)
# query
t0 = time.perf_counter(); q = g.query_sequences(pin.numpy()[:150 * 1000000], pino.numpy().view(np.uint64)[:1000001]); dt = time.perf_counter() - t0
print("query_sequences 1M reads e2e: %.2f ms %.2f G k-mers/s all-ones=%s" % (dt * 1e3, q.size / dt / 1e9, bool((q == 1).all())), flush=True)
