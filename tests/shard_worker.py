"""Worker of the multi-rank tests (one process per rank; RANK / WORLD_SIZE / MASTER_* from the env).

    python tests/shard_worker.py cpu  KIND   -- gloo; the two kernels are emulated with numpy on top of the oracle's
                                                hashes, so what is tested is the plan, the buffer layout and the
                                                all-to-all bookkeeping of goetia_b200/shard.py (no GPU needed)
    python tests/shard_worker.py cpu_p2p KIND -- gloo; the same for the peer transport's inbox layout, fill exchange and
                                                overflow lists (gt_shard_peer_layout, ShardPlan.fill_perm)
    python tests/shard_worker.py cuda KIND   -- nccl; the real library on one GPU per rank

Rank 0 gathers every rank's table parts, concatenates them in rank order and compares them byte
for byte with the oracle's tables after the same reads.  Exit code 0 = equal.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from goetia_b200.shard import ShardExchange, ShardPlan  # noqa: E402
from oracle.binding import Port  # noqa: E402
from tests.util import genome_reads  # noqa: E402


def part_bytes(kind, lo, hi, last):
    if hi <= lo:
        return 0
    if kind == 0:
        return hi // 8 + 1 - lo // 8 if last else (hi - lo) // 8
    if kind == 1:
        return hi - lo
    return hi // 2 + 1 - lo // 2 if last else (hi - lo) // 2


def np_apply(kind, slab, local_slots):
    """Emulates k_apply on a local part (slot indices relative to the part's first slot)."""
    if kind == 0:
        np.bitwise_or.at(slab, local_slots >> 3, (1 << (local_slots & 7)).astype(np.uint8))
    elif kind == 1:
        cnt = np.bincount(local_slots, minlength=slab.size)[:slab.size]
        slab[:] = np.minimum(255, slab.astype(np.int64) + cnt).astype(np.uint8)
    else:
        cnt = np.bincount(local_slots, minlength=2 * slab.size)
        for par, sh in ((1, 0), (0, 4)):  # odd slot = low nibble, even slot = high nibble
            c = cnt[par::2][:slab.size]
            cur = (slab >> sh) & 15
            new = np.minimum(15, cur.astype(np.int64) + c[:slab.size]).astype(np.uint8)
            slab[:] = (slab & (0xF0 if sh == 0 else 0x0F)) | (new << sh)


def route_mode(kind, rank, world, K, sizes, bases, offsets, my_b, my_o, budget, slice_log2):
    """gloo: ShardRouter's plumbing (route / ship / back) with the owner-side compute emulated in numpy on the oracle's
    tables: routed queries and the first-toucher flags of a tracked insert must equal the oracle's serial answers."""
    from goetia_b200.shard import ShardRouter
    dist.init_process_group("gloo", rank=rank, world_size=world)
    plan = ShardPlan(kind, sizes, world, budget, slice_log2)
    T = len(sizes)
    spr = [max(1, int(plan.own_hi[0, t] - plan.own_lo[0, t])) for t in range(T)]
    rt = ShardRouter(torch, dist, world)

    def requests(values):
        req = np.zeros(values.size * T, dtype=np.uint64)
        own = np.zeros(values.size * T, dtype=np.int64)
        for t in range(T):
            b = values % np.uint64(sizes[t])
            req[t::T] = (np.uint64(t) << np.uint64(59)) | b
            own[t::T] = (b // np.uint64(spr[t])).astype(np.int64)
        return torch.from_numpy(req.view(np.int64)), torch.from_numpy(own)

    def hashes_of(b, o):
        out = []
        for r in range(o.size - 1):
            fw, rc = Port.hash_sequence(1, K, b[int(o[r]):int(o[r + 1])].tobytes().decode())
            out.append(np.minimum(fw, rc))
        return np.concatenate(out) if out else np.zeros(0, dtype=np.uint64)

    # the state the owners answer from: the oracle's tables after the first half of the reads (every rank holds a copy and
    # answers only for the slots it owns -- a request for a foreign slot is an error of the routing)
    half = (offsets.size - 1) // 2
    ref = Port(kind, 1, K, sizes)
    ref.insert_reads(bases[:int(offsets[half])], offsets[:half + 1])

    def slot_values(got):
        g = got.numpy().view(np.uint64)
        t = (g >> np.uint64(59)).astype(np.int64)
        b = g & np.uint64((1 << 59) - 1)
        out = np.zeros(g.size, dtype=np.uint8)
        for i in range(g.size):
            assert plan.own_lo[rank, t[i]] <= b[i] < plan.own_hi[rank, t[i]], "request routed to the wrong owner"
            tab = tabs[int(t[i])]
            bb = int(b[i])
            out[i] = ((tab[bb >> 3] >> (bb & 7)) & 1) if kind == 0 else tab[bb] if kind == 1 else (
                (tab[bb >> 1] & 15) if (bb & 1) else (tab[bb >> 1] >> 4))
        return out

    tabs = ref.tables()
    vals = hashes_of(my_b, my_o)
    req, own = requests(vals)
    handle, got = rt.route(req, own)
    ans = torch.from_numpy(slot_values(got))
    back = rt.back(handle, ans).numpy().reshape(-1, T)
    got_q = back.min(axis=1).astype(np.int16) if vals.size else np.zeros(0, dtype=np.int16)
    want_q = ref.query_hashes(vals)
    ok = np.array_equal(got_q, want_q)
    # tracked insert of ALL reads' hashes on top: first-toucher flags by serial ordinal (ranks in order)
    ns = [None] * world
    dist.all_gather_object(ns, int(vals.size))
    base = sum(ns[:rank])
    ords = torch.from_numpy(np.repeat(base + np.arange(vals.size, dtype=np.int64), T))
    got_ord = rt.ship(handle, ords).numpy()
    g = got.numpy().view(np.uint64)
    zero_before = slot_values(got) == 0
    first = np.zeros(g.size, dtype=np.uint8)
    winners = {}
    for i in range(g.size):
        if zero_before[i]:
            k = int(g[i])
            if k not in winners or got_ord[i] < winners[k]:
                winners[k] = int(got_ord[i])
    for i in range(g.size):
        first[i] = 1 if zero_before[i] and winners[int(g[i])] == int(got_ord[i]) else 0
    flags = rt.back(handle, torch.from_numpy(first)).numpy().reshape(-1, T)
    is_new = flags.max(axis=1) if vals.size else np.zeros(0, dtype=np.uint8)
    # the oracle: every rank's hashes in rank order, one at a time
    all_vals = [None] * world
    dist.all_gather_object(all_vals, vals.tobytes())
    want_new = []
    for q in range(world):
        v = np.frombuffer(all_vals[q], dtype=np.uint64)
        flags_q = np.array([ref.insert(int(h)) for h in v], dtype=np.uint8)
        if q == rank:
            want_new = flags_q
    ok = ok and np.array_equal(is_new, want_new)
    ref.close()
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("shard_worker cpu_route kind=%d world=%d: %s" % (kind, world, "routes exact" if int(flag.item()) else "MISMATCH"))
    dist.destroy_process_group()
    return 0 if int(flag.item()) else 1


def main():
    mode, kind = sys.argv[1], int(sys.argv[2])
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    K = [31, 21, 25][kind]
    x_size = int(os.environ.get("SHARD_TABLE_X", "300000"))
    n_reads = int(os.environ.get("SHARD_READS", "600"))
    sizes = Port.primes_near(4, x_size)
    bases, offsets = genome_reads(n_reads, 100, 4000, seed=5 + kind)
    per = (n_reads + world - 1) // world
    r0, r1 = min(n_reads, rank * per), min(n_reads, (rank + 1) * per)
    my_b = bases[int(offsets[r0]):int(offsets[r1])]
    my_o = offsets[r0:r1 + 1] - offsets[r0]
    # gt_insert_sequences_dev bounds a batch by its bases; x3 because these reads are heavily duplicated
    # (a 4 kb genome), so bucket loads are far from the uniform share the capacities are sized for
    # (SHARD_BUDGET_X=1 makes buckets of foreign slices overflow: the overflow lists of the peer transport)
    budget = max(1024, per * 100 * int(os.environ.get("SHARD_BUDGET_X", "3")))
    slice_log2 = int(os.environ.get("SHARD_SLICE_LOG2", "10"))

    if mode == "cpu_route":
        return route_mode(kind, rank, world, K, sizes, bases, offsets, my_b, my_o, budget, slice_log2)
    if mode == "cpu":
        dist.init_process_group("gloo", rank=rank, world_size=world)
        plan = ShardPlan(kind, sizes, world, budget, slice_log2)
        x = ShardExchange(plan, rank, torch, "cpu")
        first = [int(np.nonzero(plan.table == t)[0][0]) for t in range(len(sizes))]
        fill = np.zeros(plan.nb, dtype=np.int64)
        outbox = x.outbox.numpy()
        for r in range(my_o.size - 1):
            seq = my_b[int(my_o[r]):int(my_o[r + 1])].tobytes()
            fw, rc = Port.hash_sequence(1, K, seq)
            h = np.minimum(fw, rc)
            for t, size in enumerate(sizes):
                bins = h % np.uint64(size)
                bs = first[t] + (bins >> np.uint64(plan.shift)).astype(np.int64)
                offs = (bins & np.uint64((1 << plan.shift) - 1)).astype(np.int64)
                for b, o in zip(bs, offs):
                    assert fill[b] < plan.cap[b]
                    outbox[x.bucket_off[b] + fill[b]] = o
                    fill[b] += 1
        x.fill_send[:plan.nb] = torch.from_numpy(fill.astype(np.int32))
        x.exchange()
        parts = []
        for t, size in enumerate(sizes):
            lo, hi = int(plan.own_lo[rank, t]), int(plan.own_hi[rank, t])
            slab = np.zeros(part_bytes(kind, lo, hi, hi == size), dtype=np.uint8)
            for j, b in enumerate(plan.owned[rank]):
                if plan.table[b] != t:
                    continue
                for q in range(world):
                    ent = x.source_view(q, j).numpy().astype(np.int64)
                    if ent.size:
                        np_apply(kind, slab, int(plan.slot0[b]) - lo + ent)
            parts.append(slab)
    elif mode == "cpu_p2p":
        # The peer transport's layout (gt_shard_peer_layout, the arithmetic gt_storage_attach_peers uses) and fill
        # bookkeeping (ShardPlan.fill_perm), with the NVLink stores emulated by an all-to-all of the regions: rank p
        # writes region p of every owner's inbox; every 17th update of a foreign bucket takes the overflow-list route.
        dist.init_process_group("gloo", rank=rank, world_size=world)
        plan = ShardPlan(kind, sizes, world, budget, slice_log2)
        lay, perm = plan.peer_layout(), plan.fill_perm()
        R, in_region = lay["region"], lay["in_region"]
        assert all(int(lay["ovf_offset_bytes"][q]) >= world * int(R[q]) * 4 and int(lay["ovf_offset_bytes"][q]) % 16 == 0
                   for q in range(world))
        first = [int(np.nonzero(plan.table == t)[0][0]) for t in range(len(sizes))]
        regions = [np.zeros(int(R[q]), dtype=np.int32) for q in range(world)]  # what I store into q's inbox
        ovf = [[] for _ in range(world)]
        fill = np.zeros(plan.nb + world, dtype=np.int64)
        seen = 0
        for r in range(my_o.size - 1):
            seq = my_b[int(my_o[r]):int(my_o[r + 1])].tobytes()
            fw, rc = Port.hash_sequence(1, K, seq)
            h = np.minimum(fw, rc)
            for t, size in enumerate(sizes):
                bins = h % np.uint64(size)
                for bn in bins:
                    b = first[t] + (int(bn) >> plan.shift)
                    q = int(plan.owner[b])
                    seen += 1
                    if q != rank and seen % 17 == 0:
                        ovf[q].append((t << 59) | int(bn))
                        fill[plan.nb + q] += 1
                        continue
                    assert fill[b] < plan.cap[b]
                    regions[q][int(in_region[b]) + int(fill[b])] = int(bn) & ((1 << plan.shift) - 1)
                    fill[b] += 1
        # "stores": region `rank` of every inbox
        inbox = torch.zeros(world * int(R[rank]), dtype=torch.int32)
        dist.all_to_all_single(inbox, torch.from_numpy(np.concatenate(regions)), [int(R[rank])] * world, [int(R[q]) for q in range(world)])
        n_own = plan.n_owned[rank]
        fill_recv = torch.zeros(world * (n_own + 1), dtype=torch.int64)
        dist.all_to_all_single(fill_recv, torch.from_numpy(fill[perm]), [n_own + 1] * world, [plan.n_owned[q] + 1 for q in range(world)])
        fill_recv = fill_recv.numpy().reshape(world, n_own + 1)
        cap_rec = max(1, max(len(o) for o in ovf))
        cap_t = torch.tensor([cap_rec])
        dist.all_reduce(cap_t, op=dist.ReduceOp.MAX)
        cap_rec = int(cap_t.item())
        lists_out = np.zeros((world, cap_rec), dtype=np.int64)
        for q in range(world):
            lists_out[q, :len(ovf[q])] = np.array(ovf[q], dtype=np.uint64).view(np.int64) if ovf[q] else []
        lists_in = torch.zeros(world * cap_rec, dtype=torch.int64)
        dist.all_to_all_single(lists_in, torch.from_numpy(lists_out.reshape(-1)))
        lists_in = lists_in.numpy().view(np.uint64).reshape(world, cap_rec)
        inbox = inbox.numpy()
        parts = []
        for t, size in enumerate(sizes):
            lo, hi = int(plan.own_lo[rank, t]), int(plan.own_hi[rank, t])
            slab = np.zeros(part_bytes(kind, lo, hi, hi == size), dtype=np.uint8)
            for j, b in enumerate(plan.owned[rank]):
                if plan.table[b] != t:
                    continue
                for p in range(world):
                    n = int(fill_recv[p, j])
                    o = p * int(R[rank]) + int(in_region[b])
                    ent = inbox[o:o + n].astype(np.int64)
                    if ent.size:
                        np_apply(kind, slab, int(plan.slot0[b]) - lo + ent)
            for p in range(world):
                recs = lists_in[p, :int(fill_recv[p, n_own])]
                mine = recs[(recs >> np.uint64(59)) == np.uint64(t)] & np.uint64((1 << 59) - 1)
                if mine.size:
                    assert int(mine.min()) >= lo and int(mine.max()) < hi
                    for v in mine.astype(np.int64):  # one at a time: repeated slots must count every time
                        np_apply(kind, slab, np.array([v - lo], dtype=np.int64))
            parts.append(slab)
    else:
        import goetia_b200 as gb
        from goetia_b200 import _capi
        from goetia_b200.shard import ShardedStorage
        local = int(os.environ.get("LOCAL_RANK", rank))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
        gb.init(local)
        st = ShardedStorage(kind, sizes, budget, slice_log2_bytes=slice_log2,
                            transport=os.environ.get("SHARD_TRANSPORT", "p2p"))
        plan = st.plan
        d_b = torch.from_numpy(my_b.copy()).cuda()
        d_o = torch.from_numpy(my_o.astype(np.int64)).cuda()
        torch.cuda.synchronize()
        rounds = int(os.environ.get("SHARD_ROUNDS", "2"))
        for _ in range(rounds):  # the same reads twice: counters must count both passes
            nk = st.bucket_sequences_dev(_capi.SHIFTER_CAN, K, d_b.data_ptr(), d_o.data_ptr(), my_o.size - 1, my_b.size)
            assert nk == (my_o.size - 1) * (100 - K + 1), nk
            st.exchange_and_apply()
        st.synchronize()
        info = st.pending_info()
        assert info["pending_kmers"] == 0
        parts = st.local_tables()
        # routed queries: every rank asks about its own reads (plus one absent k-mer) and gets the oracle's counts
        q_o = my_o[:min(my_o.size, 41)]
        q_b = my_b[:int(q_o[-1])]
        x_b, x_o = genome_reads(5 + rank, 100, 4000, seed=900 + rank)  # mostly absent k-mers; a different count per rank
        q_b = np.concatenate([q_b, x_b])
        q_o = np.concatenate([q_o, x_o[1:] + q_o[-1]])
        got_q = st.query_sequences(_capi.SHIFTER_CAN, K, q_b, q_o)
        ref_q = Port(kind, 1, K, sizes)
        for _ in range(rounds):
            ref_q.insert_reads(bases, offsets)
        want_q = ref_q.query_reads(q_b, q_o)
        if not np.array_equal(got_q, want_q):
            print("rank %d: routed query differs (%d of %d)" % (rank, int((got_q != want_q).sum()), want_q.size))
            parts = [p[:0] for p in parts]  # forces a mismatch on rank 0
        ref_q.close()
        # tracked inserts on the sharded tables (the n_unique reverse route): per-read n_new under the serial rule, ranks in
        # order == the oracle's one-at-a-time loop over the whole read set; then OXLI save / load of the whole storage
        saved_parts = parts
        st.reset()
        nk_t, n_new = st.insert_sequences_tracked(_capi.SHIFTER_CAN, K, my_b, my_o)
        ref_t = Port(kind, 1, K, sizes)
        _, _, want_new = ref_t.insert_reads(bases, offsets, want_n_new=True)
        good = nk_t == (my_o.size - 1) * (100 - K + 1) and np.array_equal(n_new, want_new[r0:r1])
        good = good and st.n_unique_kmers() == ref_t.stats()[0]
        fn = os.path.join(os.environ.get("SHARD_TMP", "/tmp"), "shard_%d_%d.oxli" % (kind, world))
        st.save(fn, K)
        if rank == 0:
            ref_t.save(fn + ".ref")
            good = good and open(fn, "rb").read() == open(fn + ".ref", "rb").read()
        one_pass = st.local_tables()
        st.reset()
        good = good and st.load(fn) == K
        good = good and all(np.array_equal(a, b) for a, b in zip(st.local_tables(), one_pass))
        q1 = st.query_sequences(_capi.SHIFTER_CAN, K, q_b, q_o)
        good = good and np.array_equal(q1, ref_t.query_reads(q_b, q_o))
        if not good:
            print("rank %d: tracked insert / save / load check failed" % rank)
            saved_parts = [p[:0] for p in saved_parts]
        parts = saved_parts
        ref_t.close()
        st.close()

    # gather the parts on rank 0 and compare with the oracle
    ok = True
    gathered = [None] * world
    dist.gather_object([p.tobytes() for p in parts], gathered if rank == 0 else None, dst=0)
    if rank == 0:
        ref = Port(kind, 1, K, sizes)
        passes = 1 if mode.startswith("cpu") else int(os.environ.get("SHARD_ROUNDS", "2"))
        for _ in range(passes):
            ref.insert_reads(bases, offsets)
        for t, want in enumerate(ref.tables()):
            got = np.frombuffer(b"".join(gathered[r][t] for r in range(world)), dtype=np.uint8)
            if got.size != want.size or not np.array_equal(got, want):
                ok = False
                print("table %d differs (%d vs %d bytes, %d mismatches)" % (
                    t, got.size, want.size, int((got[:min(got.size, want.size)] != want[:min(got.size, want.size)]).sum())))
        print("shard_worker %s kind=%d world=%d shift=%d buckets=%d: %s" % (mode, kind, world, plan.shift, plan.nb,
                                                                          "tables bit-exact" if ok else "MISMATCH"))
    flag = torch.tensor([1 if ok else 0])
    if mode == "cuda":
        flag = flag.cuda()
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    return 0 if int(flag.item()) else 1


if __name__ == "__main__":
    sys.exit(main())
