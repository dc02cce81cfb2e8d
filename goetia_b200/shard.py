"""Multi-GPU layer: reads sharded over ranks, table ownership partitioned by slot range.

goetia has nothing like this; it is the new shard/exchange layer the
north star asks for.  One process per GPU (torchrun).  Every table is cut into slices of
2**shift slots and rank r holds a contiguous run of slices of every table, so concatenating the
ranks' parts in rank order reproduces the single-GPU / reference table byte for byte (``h % size``
scatters any hash range over a whole table, so ownership has to be by slot range, not by hash
range -- SURVEY.md section 8e).  Per round every rank

  1. hashes its own reads and appends each (k-mer, table) update -- a 4-byte slice-local slot
     offset -- to the bucket of the slice it falls in (k_bucket, the same kernel as on one GPU;
     buckets of foreign slices live in the outbox region of their owner),
  2. exchanges bucket fill counts and bucket regions with NCCL all-to-all over NVLink,
  3. applies its own buckets and the received ones slice by slice (k_apply).

``ShardPlan`` (host arithmetic from the C library, no GPU) and ``ShardExchange`` (buffer layout +
collectives, any torch device / backend) are what the CPU gloo tests exercise; ``ShardedStorage``
binds them to the CUDA library.
"""
import ctypes as C

import numpy as np

from . import _capi

MAX_BUCKETS = 1024


class ShardPlan:
    """gt_shard_plan: slices, owners, capacities.  Identical on every rank."""

    def __init__(self, kind, sizes, world, budget_kmers, slice_log2_bytes=0):
        L = _capi.load()
        self.kind, self.world, self.budget_kmers = int(kind), int(world), int(budget_kmers)
        self._slice_log2_bytes = int(slice_log2_bytes)
        self.sizes = np.ascontiguousarray(sizes, dtype=np.uint64)
        n = self.sizes.size
        shift_nb = np.zeros(2, dtype=np.int32)
        table = np.zeros(MAX_BUCKETS, dtype=np.int32)
        owner = np.zeros(MAX_BUCKETS, dtype=np.int32)
        slot0 = np.zeros(MAX_BUCKETS, dtype=np.uint64)
        slots = np.zeros(MAX_BUCKETS, dtype=np.uint64)
        cap = np.zeros(MAX_BUCKETS, dtype=np.uint32)
        own_lo = np.zeros(world * n, dtype=np.uint64)
        own_hi = np.zeros(world * n, dtype=np.uint64)
        _capi.check(L.gt_shard_plan(self.kind, self.sizes.ctypes.data_as(_capi.u64p), n, self.world, self.budget_kmers,
                                    int(slice_log2_bytes), shift_nb.ctypes.data, table.ctypes.data, owner.ctypes.data,
                                    slot0.ctypes.data, slots.ctypes.data, cap.ctypes.data, own_lo.ctypes.data,
                                    own_hi.ctypes.data), "gt_shard_plan")
        self.shift, self.nb = int(shift_nb[0]), int(shift_nb[1])
        nb = self.nb
        self.table, self.owner = table[:nb].copy(), owner[:nb].copy()
        self.slot0, self.slots, self.cap = slot0[:nb].copy(), slots[:nb].copy(), cap[:nb].astype(np.int64)
        self.own_lo = own_lo.reshape(world, n)
        self.own_hi = own_hi.reshape(world, n)
        # derived layout (include/goetia_b200.h, gt_storage_attach_exchange)
        self.owned = [np.nonzero(self.owner == r)[0] for r in range(world)]
        self.n_owned = [int(o.size) for o in self.owned]
        self.region = [int(self.cap[o].sum()) for o in self.owned]  # entries of rank r's region
        self.perm = np.concatenate(self.owned) if nb else np.zeros(0, dtype=np.int64)  # bucket ids, owner-major
        self.in_region = np.zeros(nb, dtype=np.int64)  # offset of bucket b inside its owner's region
        for o in self.owned:
            self.in_region[o] = np.cumsum(self.cap[o]) - self.cap[o]

    def peer_layout(self):
        """gt_shard_peer_layout: the inbox layout of the peer transport (what gt_storage_attach_peers lays out).
        -> dict(region[world], in_region[nb], ovf_offset_bytes[world], inbox_bytes[world])."""
        L = _capi.load()
        W = self.world
        region = np.zeros(W, dtype=np.uint64)
        in_region = np.zeros(max(self.nb, 1), dtype=np.uint64)
        ovf = np.zeros(W, dtype=np.uint64)
        inbox = np.zeros(W, dtype=np.uint64)
        _capi.check(L.gt_shard_peer_layout(self.kind, self.sizes.ctypes.data_as(_capi.u64p), self.sizes.size, W,
                                           self.budget_kmers, self._slice_log2_bytes, region.ctypes.data,
                                           in_region.ctypes.data, ovf.ctypes.data, inbox.ctypes.data),
                    "gt_shard_peer_layout")
        return {"region": region.astype(np.int64), "in_region": in_region[:self.nb].astype(np.int64),
                "ovf_offset_bytes": ovf.astype(np.int64), "inbox_bytes": inbox.astype(np.int64)}

    def fill_perm(self):
        """Gather order of the peer transport's fill exchange: fill_send = [bucket cursors (nb), overflow
        cursors per owner (world)]; owner q receives the cursors of its buckets, then the overflow count."""
        return np.concatenate([np.concatenate([self.owned[q], [self.nb + q]]) for q in range(self.world)]).astype(np.int64)

    def outbox_offsets(self, me):
        """(offset of every bucket in rank `me`'s outbox, total entries, entries destined to peers)."""
        start = np.zeros(self.world, dtype=np.int64)
        off = 0
        for q in range(self.world):
            if q != me:
                start[q] = off
                off += self.region[q]
        others = off
        start[me] = off
        off += self.region[me]
        return start[self.owner] + self.in_region, off, others

    def inbox_entries(self, me):
        return (self.world - 1) * self.region[me]


class ShardExchange:
    """The two all-to-alls of one round, on whatever device/backend the tensors live on."""

    def __init__(self, plan, rank, torch, device, group=None):
        import torch.distributed as dist
        self.plan, self.rank, self.torch, self.dist, self.group = plan, rank, torch, dist, group
        W = plan.world
        self.bucket_off, total, self.others = plan.outbox_offsets(rank)
        self.outbox = torch.zeros(max(total, 1), dtype=torch.int32, device=device)
        self.inbox = torch.zeros(max(plan.inbox_entries(rank), 1), dtype=torch.int32, device=device)
        self.fill_send = torch.zeros(max(plan.nb, 1), dtype=torch.int32, device=device)
        self.fill_recv = torch.zeros(max(W * plan.n_owned[rank], 1), dtype=torch.int32, device=device)
        self._perm = torch.as_tensor(plan.perm, dtype=torch.int64, device=device)
        self._fill_in = [plan.n_owned[p] for p in range(W)]
        self._fill_out = [plan.n_owned[rank]] * W
        self._data_in = [plan.region[p] if p != rank else 0 for p in range(W)]
        self._data_out = [plan.region[rank] if q != rank else 0 for q in range(W)]

    def exchange(self):
        """fill counts (owner-major gather of fill_send) and bucket regions -> fill_recv / inbox."""
        torch, dist, plan = self.torch, self.dist, self.plan
        fx = self.fill_send[:plan.nb].index_select(0, self._perm)
        dist.all_to_all_single(self.fill_recv[:plan.world * plan.n_owned[self.rank]], fx, self._fill_out, self._fill_in,
                               group=self.group)
        if plan.world > 1:
            dist.all_to_all_single(self.inbox[:plan.inbox_entries(self.rank)], self.outbox[:self.others],
                                   self._data_out, self._data_in, group=self.group)

    def source_view(self, q, j):
        """Entries rank q produced for this rank's j-th owned bucket (valid after exchange())."""
        plan, me = self.plan, self.rank
        b = int(plan.owned[me][j])
        n = int(self.fill_recv[q * plan.n_owned[me] + j])
        n = min(n, int(plan.cap[b]))
        if q == me:
            o = int(self.bucket_off[b])
            return self.outbox[o:o + n]
        o = (q if q < me else q - 1) * plan.region[me] + int(plan.in_region[b])
        return self.inbox[o:o + n]


class ShardRouter:
    """Owner-routed requests: the collective plumbing, on any torch device / backend.

    A source rank turns its hash values into requests (table << 59 | slot) with an owner rank each; `route` ships every
    request to its owner with one all-to-all of counts and one of payloads, `ship` sends a second payload (e.g. serial
    ordinals) along the same route, and `back` returns one answer per request along the reverse route and puts the answers
    in request order.  The compute (making requests, answering them) is done by the caller -- CUDA kernels through the C
    ABI on a GPU, numpy in the gloo tests."""

    def __init__(self, torch, dist, world, group=None):
        self.torch, self.dist, self.world, self.group = torch, dist, world, group

    def route(self, req, owner):
        torch, dist = self.torch, self.dist
        order = torch.argsort(owner, stable=True)
        send = torch.bincount(owner, minlength=self.world).to(torch.int64)
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=self.group)
        send_l, recv_l = [int(x) for x in send.tolist()], [int(x) for x in recv.tolist()]
        got = torch.empty(sum(recv_l), dtype=req.dtype, device=req.device)
        dist.all_to_all_single(got, req[order].contiguous(), recv_l, send_l, group=self.group)
        return (order, send_l, recv_l), got

    def ship(self, handle, payload):
        order, send_l, recv_l = handle
        got = self.torch.empty(sum(recv_l), dtype=payload.dtype, device=payload.device)
        self.dist.all_to_all_single(got, payload[order].contiguous(), recv_l, send_l, group=self.group)
        return got

    def back(self, handle, answers):
        order, send_l, recv_l = handle
        got = self.torch.empty(sum(send_l), dtype=answers.dtype, device=answers.device)
        self.dist.all_to_all_single(got, answers.contiguous(), send_l, recv_l, group=self.group)
        out = self.torch.empty_like(got)
        out[order] = got
        return out


class ShardedStorage:
    """This rank's part of a BitStorage / ByteStorage / NibbleStorage sharded over the process group.

    Two transports move the buckets of foreign slices to their owners:

    ``p2p``            k_bucket stores every run of entries straight into the owner's inbox over
                       NVLink (peer memory mapped with CUDA IPC, ``gt_storage_attach_peers``): the
                       transfer is fused with the hashing, run by run.  Only the per-bucket fill
                       counts travel by collective (one small all-to-all), which doubles as the
                       "every producer has finished" signal.
    ``nccl``           k_bucket fills a local outbox; one NCCL all-to-all ships the regions.
    ``ce`` (default)   the peer layout, but k_bucket writes foreign buckets into LOCAL staging areas
                       (``gt_storage_attach_staged``) and the copy engines ship them into the owners' inboxes
                       (``gt_peer_copy_async`` on copy streams) while the SMs hash the next round: NVLink costs no SM
                       time and no CTA slots, so k_bucket and the window apply run at their single-GPU rates.  The
                       round is finished -- fill counts exchanged, buckets applied -- one call later
                       (``exchange_and_apply`` of the next round, or ``join`` / ``synchronize``, which are therefore
                       COLLECTIVE with this transport: every rank must call them at the same points).

    Either way there are two buffer sets and two streams: k_apply of round i (apply stream)
    overlaps k_bucket of round i+1 into the other set (compute stream).
    """

    def __init__(self, kind, sizes, budget_kmers, group=None, slice_log2_bytes=0, transport=None):
        import os
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.transport = transport or os.environ.get("GT_SHARD_TRANSPORT", "ce")
        if self.transport not in ("p2p", "nccl", "ce"):
            raise ValueError("transport must be 'p2p', 'ce' or 'nccl'")
        L = _capi.lib()
        self.kind = int(kind)
        self.plan = plan = ShardPlan(kind, sizes, self.world, budget_kmers, slice_log2_bytes)
        self._sizes = plan.sizes
        self._h = L.gt_storage_create_sharded(self.kind, self._sizes.ctypes.data_as(_capi.u64p), self._sizes.size,
                                              self.rank, self.world, int(budget_kmers), int(slice_log2_bytes))
        if not self._h:
            raise _capi.GoetiaB200Error("gt_storage_create_sharded: " + _capi.last_error())
        dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
        self.stream = torch.cuda.Stream(priority=hi)   # pack / hash / bucket (+ the exchange)
        self.apply_stream = torch.cuda.Stream(priority=lo)
        _capi.check(L.gt_set_compute_stream(self.stream.cuda_stream), "gt_set_compute_stream")
        _capi.check(L.gt_set_apply_stream(self.apply_stream.cuda_stream), "gt_set_apply_stream")
        self.cur = 0
        self._applied = [None, None]   # event: k_apply of the set has finished (its buffers are free again)
        self._peer_ptrs, self._own_inbox = [], []
        self._pending = None           # ce transport: (set, copies-landed events) of the round still to be finished
        self._inbox_ptrs, self._stage = [], []
        self.shipped_bytes = 0         # ce transport: bytes handed to the copy engines (whole regions + overflow lists)
        self._copy_prof = [] if os.environ.get("GT_SHARD_PROFILE_COPIES", "0") == "1" else None
        W, me = self.world, self.rank
        if self.transport == "nccl":
            self.sets = [ShardExchange(plan, me, torch, dev, group) for _ in range(2)]
            for w, x in enumerate(self.sets):
                _capi.check(L.gt_storage_attach_exchange(self._h, w, x.outbox.data_ptr(), x.inbox.data_ptr(),
                                                         x.fill_send.data_ptr(), x.fill_recv.data_ptr()),
                            "gt_storage_attach_exchange")
            self.x = self.sets[0]
        else:
            # fill_send = [bucket cursors (nb), overflow-list cursors per owner (W)]; what goes to owner q is
            # the cursors of q's buckets followed by the overflow count for q
            self._perm = torch.as_tensor(plan.fill_perm(), dtype=torch.int64, device=dev)
            self._fill_in = [plan.n_owned[q] + 1 for q in range(W)]
            self._fill_out = [plan.n_owned[me] + 1] * W
            self._n_fill = plan.nb + W
            self.fill_send, self.fill_recv, self._fx = [], [], []
            if self.transport == "ce":
                lay = plan.peer_layout()
                self._R = [int(x) for x in lay["region"]]
                self._ovf_off = [int(x) for x in lay["ovf_offset_bytes"]]
                self._ovf_bytes = [(int(lay["inbox_bytes"][q]) - self._ovf_off[q]) // W for q in range(W)]
            for w in range(2):
                own = L.gt_peer_alloc(int(L.gt_storage_inbox_bytes(self._h, me)))
                if not own:
                    raise _capi.GoetiaB200Error("gt_peer_alloc: " + _capi.last_error())
                self._own_inbox.append(own)
                handle = np.zeros(64, dtype=np.uint8)
                _capi.check(L.gt_peer_export(own, handle.ctypes.data), "gt_peer_export")
                mine = torch.from_numpy(handle).to(dev)
                allh = torch.empty(W * 64, dtype=torch.uint8, device=dev)
                dist.all_gather_into_tensor(allh, mine, group=group)
                allh = allh.cpu().numpy().reshape(W, 64)
                ptrs = (C.c_void_p * W)()
                for q in range(W):
                    if q == me:
                        ptrs[q] = own
                    else:
                        hq = np.ascontiguousarray(allh[q])
                        pq = L.gt_peer_open(hq.ctypes.data)
                        if not pq:
                            raise _capi.GoetiaB200Error("gt_peer_open(rank %d): %s" % (q, _capi.last_error()))
                        self._peer_ptrs.append(pq)
                        ptrs[q] = pq
                fs = torch.zeros(self._n_fill, dtype=torch.int32, device=dev)
                fr = torch.zeros(W * (plan.n_owned[me] + 1), dtype=torch.int32, device=dev)
                self.fill_send.append(fs)
                self.fill_recv.append(fr)
                self._fx.append(torch.zeros(self._n_fill, dtype=torch.int32, device=dev))
                if self.transport == "ce":
                    # local staging area per foreign owner; the copy engines ship it (exchange_and_apply).  (Mixing in direct
                    # NVLink stores from k_bucket for some of the peers -- gt_storage_attach_areas allows it -- was measured at
                    # 8 ranks and lost: three direct peers slow k_bucket from 2.3 to 5.8 ms per round, 158 vs 226 G k-mers/s.)
                    areas = [None if q == me else torch.zeros(int(L.gt_storage_stage_bytes(self._h, q)), dtype=torch.uint8, device=dev)
                             for q in range(W)]
                    sp = (C.c_void_p * W)()
                    for q in range(W):
                        sp[q] = None if q == me else areas[q].data_ptr()
                    _capi.check(L.gt_storage_attach_staged(self._h, w, own, sp, fs.data_ptr(), fr.data_ptr()),
                                "gt_storage_attach_staged")
                    self._stage.append(areas)
                    self._inbox_ptrs.append([int(ptrs[q]) for q in range(W)])
                else:
                    _capi.check(L.gt_storage_attach_peers(self._h, w, ptrs, fs.data_ptr(), fr.data_ptr()),
                                "gt_storage_attach_peers")
            if self.transport == "ce":
                n_cs = max(1, min(int(os.environ.get("GT_SHARD_COPY_STREAMS", "4")), W - 1))
                self.copy_streams = [torch.cuda.Stream() for _ in range(n_cs)]
            torch.cuda.synchronize()
            dist.barrier(group=group)  # every rank has mapped every inbox before anyone stores into one

    @property
    def handle(self):
        return self._h

    def bucket_sequences_dev(self, shifter_kind, K, d_bases_ptr, d_offsets_ptr, n_reads, n_bases, n_kmers_upper=None):
        """Step 1 of a round: pack + hash + bucket this rank's (device-resident) reads into the
        current buffer set (p2p: the entries of foreign slices go straight to their owners)."""
        L = _capi.lib()
        with self.torch.cuda.stream(self.stream):
            if self._applied[self.cur] is not None:
                self.stream.wait_event(self._applied[self.cur])
            _capi.check(L.gt_storage_select_store(self._h, self.cur), "gt_storage_select_store")
            if n_kmers_upper is not None:  # equal-length reads: the caller knows the k-mer count, so the round budget can be tight
                _capi.check(L.gt_storage_hint_kmers(self._h, int(n_kmers_upper)), "gt_storage_hint_kmers")
            return int(_capi.check(L.gt_insert_sequences_dev(self._h, shifter_kind, K, d_bases_ptr, d_offsets_ptr,
                                                             n_reads, n_bases, _capi.MODE_BLIND),
                                   "gt_insert_sequences_dev"))

    def bucket_sequences_dev_async(self, shifter_kind, K, d_bases_ptr, d_offsets_ptr, n_reads, n_bases, d_kmer_total_ptr=None,
                                   n_kmers_upper=None):
        """bucket_sequences_dev without the host wait (the k-mer count is added to a device uint64)."""
        L = _capi.lib()
        with self.torch.cuda.stream(self.stream):
            if self._applied[self.cur] is not None:
                self.stream.wait_event(self._applied[self.cur])
            _capi.check(L.gt_storage_select_store(self._h, self.cur), "gt_storage_select_store")
            if n_kmers_upper is not None:  # equal-length reads: the caller knows the k-mer count, so the round budget can be tight
                _capi.check(L.gt_storage_hint_kmers(self._h, int(n_kmers_upper)), "gt_storage_hint_kmers")
            _capi.check(L.gt_insert_sequences_dev_async(self._h, shifter_kind, K, d_bases_ptr, d_offsets_ptr, n_reads, n_bases,
                                                        _capi.MODE_BLIND, d_kmer_total_ptr), "gt_insert_sequences_dev_async")

    def bucket_packed_dev_async(self, shifter_kind, K, d_words_ptr, n_words_alloc, d_offsets_ptr, d_flags_ptr, n_reads, n_bases,
                                d_kmer_total_ptr=None, n_kmers_upper=None):
        """Step 1 of a round for a batch that is already 2-bit packed in HBM (the host packed it, so only 0.25 B/base
        crossed PCIe): hash + bucket, no pack kernel, no host wait."""
        L = _capi.lib()
        with self.torch.cuda.stream(self.stream):
            if self._applied[self.cur] is not None:
                self.stream.wait_event(self._applied[self.cur])
            _capi.check(L.gt_storage_select_store(self._h, self.cur), "gt_storage_select_store")
            if n_kmers_upper is not None:  # equal-length reads: the caller knows the k-mer count, so the round budget can be tight
                _capi.check(L.gt_storage_hint_kmers(self._h, int(n_kmers_upper)), "gt_storage_hint_kmers")
            _capi.check(L.gt_insert_packed_dev_async(self._h, shifter_kind, K, d_words_ptr, n_words_alloc, d_offsets_ptr, d_flags_ptr,
                                                     n_reads, n_bases, _capi.MODE_BLIND, d_kmer_total_ptr),
                        "gt_insert_packed_dev_async")

    def exchange_and_apply(self):
        """Steps 2 and 3 of a round (collective: every rank calls it once per round): exchange of
        the current set, k_apply of this rank's slices on the apply stream, switch sets.  (ce transport: the round's
        staging areas go to the copy engines now; its fill exchange and apply are queued by the next call.)"""
        torch, w = self.torch, self.cur
        if self.transport == "ce":
            return self._ship_and_finish_previous()
        if self.transport == "p2p":
            # the fill exchange is stream-ordered after k_bucket on every rank, so its completion is also the signal
            # that all peer stores have landed
            self._finish_fills_and_apply(w)
        else:
            with torch.cuda.stream(self.stream):
                self.sets[w].exchange()
                ev = torch.cuda.Event()
                ev.record(self.stream)
            self.apply_stream.wait_event(ev)
            _capi.check(_capi.lib().gt_storage_apply_store(self._h, w), "gt_storage_apply_store")
            done = torch.cuda.Event()
            done.record(self.apply_stream)
            self._applied[w] = done
        self.cur = w ^ 1

    # ---- ce transport: copy engines ship round r while the SMs finish round r-1 and hash round r+1 -------------------
    def _finish_fills_and_apply(self, w):
        """fill-count all-to-all of set w (after this rank's copies of that set have landed) + apply (p2p and ce)."""
        torch, dist, plan = self.torch, self.dist, self.plan
        with torch.cuda.stream(self.stream):
            # Peers start writing into my OTHER set as soon as they are past this exchange, so it must not complete
            # before that set's k_apply has finished here.
            if self._applied[w ^ 1] is not None:
                self.stream.wait_event(self._applied[w ^ 1])
            torch.index_select(self.fill_send[w], 0, self._perm, out=self._fx[w])
            # software NVLink counter: entries this rank produced for peers' slices this round (cursors of foreign
            # buckets, 16-byte run padding included); nvidia-smi's link counters are not available on every box.
            # (Plain elementwise ops only: nothing here may make the host wait for the stream.)
            if getattr(self, "_foreign", None) is None:
                own = torch.as_tensor(np.asarray(plan.owner) != self.rank, device=self.device)
                self._foreign = own.to(torch.int64)
                self._peer_entries = torch.zeros(1, dtype=torch.int64, device=self.device)
            self._peer_entries += (self.fill_send[w][:plan.nb].to(torch.int64) * self._foreign).sum()
            dist.all_to_all_single(self.fill_recv[w], self._fx[w], self._fill_out, self._fill_in, group=self.group)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.apply_stream.wait_event(ev)
        _capi.check(_capi.lib().gt_storage_apply_store(self._h, w), "gt_storage_apply_store")
        done = torch.cuda.Event()
        done.record(self.apply_stream)
        self._applied[w] = done

    def _finish_pending(self):
        if self._pending is None:
            return
        w, landed = self._pending
        self._pending = None
        for ev in landed:
            self.stream.wait_event(ev)  # stream order: my copies of that round are in the peers' inboxes before the all-to-all
        self._finish_fills_and_apply(w)

    def _ship_and_finish_previous(self):
        """ce transport, round r (set w): finish round r-1 (its copies have had k_bucket of round r to land), then hand
        round r's staging areas to the copy engines.  The copies start after the all-to-all of round r-1 has completed
        here -- every peer has then entered it, i.e. is past its apply of round r-2, the last reader of the inbox set
        these copies write -- and after k_bucket of round r (both by stream order: one event on the compute stream).
        (Letting the copies start earlier, on a separate "every rank has applied" all-reduce issued from a side stream,
        was tried at 8 ranks and measured 172 vs 226 G k-mers/s.  A collective that is not stream-ordered between two
        kernels has to wait for a CTA slot beside the SM-filling kernels on every rank; the measured build also made the
        host wait once per round (a boolean-mask index in the byte counter), so the comparison is not clean -- the
        variant is not in the tree, the validated order is.)"""
        torch, L, w, me, W = self.torch, _capi.lib(), self.cur, self.rank, self.world
        self._finish_pending()
        go = torch.cuda.Event()
        go.record(self.stream)
        landed = []
        for cs in self.copy_streams:
            cs.wait_event(go)
        prof = self._copy_prof is not None
        if prof:  # GT_SHARD_PROFILE_COPIES=1: time every round's copies (first copy stream started -> last copy landed)
            t0 = torch.cuda.Event(enable_timing=True)
            t0.record(self.copy_streams[0])
        n_bytes = 0
        for i in range(1, W):
            q = (me + i) % W  # every rank starts with a different destination
            cs = self.copy_streams[(i - 1) % len(self.copy_streams)]
            src = self._stage[w][q].data_ptr()
            dst = self._inbox_ptrs[w][q]
            reg = self._R[q] * 4
            _capi.check(L.gt_peer_copy_async(dst + me * reg, src, reg, cs.cuda_stream), "gt_peer_copy_async")
            _capi.check(L.gt_peer_copy_async(dst + self._ovf_off[q] + me * self._ovf_bytes[q], src + (reg + 15) // 16 * 16,
                                             self._ovf_bytes[q], cs.cuda_stream), "gt_peer_copy_async")
            n_bytes += reg + self._ovf_bytes[q]
        self.shipped_bytes += n_bytes
        for cs in self.copy_streams:
            ev = torch.cuda.Event(enable_timing=prof)
            ev.record(cs)
            landed.append(ev)
        if prof:
            self._copy_prof.append((t0, landed, n_bytes))
        self._pending = (w, landed)
        self.cur = w ^ 1

    def copy_profile(self, reset=True):
        """ce transport with GT_SHARD_PROFILE_COPIES=1: (rounds, mean ms per round of copies, GB/s while copying) since the
        last reset.  Synchronises (collective)."""
        if not self._copy_prof:
            return None
        self.synchronize()
        ms = [max(t0.elapsed_time(e) for e in landed) for t0, landed, _ in self._copy_prof]
        nb = sum(b for _, _, b in self._copy_prof)
        out = {"rounds": len(ms), "ms_per_round": sum(ms) / len(ms), "GBps_while_copying": nb / (sum(ms) / 1e3) / 1e9 if sum(ms) > 0 else None}
        if reset:
            self._copy_prof = []
        return out

    def peer_store_bytes(self, reset=False):
        """Bytes this rank's k_bucket stored straight into peers' HBM over NVLink since the last reset (p2p transport;
        counted from the bucket cursors, so the 16-byte run padding is included).  Synchronises."""
        if getattr(self, "_peer_entries", None) is None:
            return 0
        self.synchronize()
        v = int(self._peer_entries.item()) * 4
        if reset:
            self._peer_entries.zero_()
        return v

    def join(self):
        """Order the compute stream after every k_apply queued so far (no host wait): an event
        recorded on ``stream`` after this covers the whole round, both streams."""
        self._finish_pending()  # ce transport: collective
        for ev in self._applied:
            if ev is not None:
                self.stream.wait_event(ev)

    def synchronize(self):
        self._finish_pending()  # ce transport: collective
        self.stream.synchronize()
        self.apply_stream.synchronize()

    # ---- owner-routed queries and tracked inserts ---------------------------------------------
    def _router(self):
        if getattr(self, "_rt", None) is None:
            self._rt = ShardRouter(self.torch, self.dist, self.world, self.group)
        return self._rt

    def _requests(self, values):
        """values: int64 device tensor of hash values -> (requests int64 [m * T], owners int64 [m * T])"""
        torch = self.torch
        T, m = int(self._sizes.size), int(values.numel())
        req = torch.empty(max(m * T, 1), dtype=torch.int64, device=self.device)
        owner = torch.empty(max(m * T, 1), dtype=torch.int32, device=self.device)
        _capi.check(_capi.lib().gt_shard_route_hashes_dev(self._h, values.data_ptr(), m, req.data_ptr(), owner.data_ptr()),
                    "gt_shard_route_hashes_dev")
        return req[:m * T], owner[:m * T].to(torch.int64)

    def query_hashes_dev(self, values):
        """Storage::query for a device tensor of hash values (collective; every rank passes its own, differently sized
        vector): each (hash, table) request travels to the rank that holds the slot (8 B), the owner answers with the
        slot's value (1 B), and the source takes the AND / min over the tables (bitstorage.cc:87-100, bytestorage.cc:
        116-139, nibblestorage.cc:112-130).  -> int16 device tensor."""
        torch = self.torch
        T, m = int(self._sizes.size), int(values.numel())
        self.join()
        with torch.cuda.stream(self.stream):
            req, owner = self._requests(values)
            handle, got = self._router().route(req, owner)
            ans = torch.empty(max(got.numel(), 1), dtype=torch.uint8, device=self.device)
            _capi.check(_capi.lib().gt_shard_answer_dev(self._h, got.data_ptr(), got.numel(), ans.data_ptr()), "gt_shard_answer_dev")
            back = self._router().back(handle, ans[:got.numel()])
            out = back.view(m, T).amin(dim=1).to(torch.int16) if m else torch.zeros(0, dtype=torch.int16, device=self.device)
        self.stream.synchronize()
        return out

    def query_hashes(self, hashes):
        h = np.ascontiguousarray(hashes, dtype=np.uint64)
        v = self.torch.from_numpy(h.view(np.int64)).to(self.device)
        return self.query_hashes_dev(v).cpu().numpy()

    def _hash_values_dev(self, shifter_kind, K, bases, offsets):
        """host reads -> int64 device tensor of hash values (value() of every k-mer, reads back to back)"""
        torch = self.torch
        bases, offsets = _capi.as_reads(bases, offsets)
        n_reads, n_bases = offsets.size - 1, int(offsets[-1] - offsets[0]) if offsets.size > 1 else 0
        vals = torch.empty(max(n_bases, 1), dtype=torch.int64, device=self.device)
        nk = 0
        if n_reads and n_bases:
            with torch.cuda.stream(self.stream):
                pad = np.zeros(n_bases + 16, dtype=np.uint8)
                pad[:n_bases] = bases[int(offsets[0]):int(offsets[-1])]
                d_b = torch.from_numpy(pad).to(self.device)
                d_o = torch.from_numpy((offsets - offsets[0]).view(np.int64)).to(self.device)
                self.stream.synchronize()
                nk = int(_capi.check(_capi.lib().gt_hash_values_dev(shifter_kind, K, d_b.data_ptr(), d_o.data_ptr(), n_reads, n_bases,
                                                                    vals.data_ptr()), "gt_hash_values_dev"))
        return vals[:nk]

    def query_sequences(self, shifter_kind, K, bases, offsets):
        """dBG::query_sequence over a host batch against the sharded tables (collective): hashed on the device, then
        owner-routed."""
        return self.query_hashes_dev(self._hash_values_dev(shifter_kind, K, bases, offsets)).cpu().numpy()

    def insert_hashes_tracked_dev(self, values):
        """Storage::insert for a device tensor of hash values with the reference's return value (collective): is_new[i]
        = the hash found a slot that was empty and no earlier hash of this call (ranks in order, each rank's hashes
        in order) had touched it -- the serial first-toucher rule.  Requests travel to the owners
        with their serial ordinals, the owners answer with one flag per request (the n_unique reverse route of
        section 8e), the source ORs a hash's flags.  -> uint8 device tensor; n_unique_kmers() is updated."""
        torch, dist = self.torch, self.dist
        T, m = int(self._sizes.size), int(values.numel())
        self.join()
        with torch.cuda.stream(self.stream):
            ns = torch.empty(self.world, dtype=torch.int64, device=self.device)
            dist.all_gather_into_tensor(ns, torch.tensor([m], dtype=torch.int64, device=self.device), group=self.group)
            ns_l = [int(x) for x in ns.tolist()]
            if sum(ns_l) >= 2**32:
                raise _capi.GoetiaB200Error("insert_hashes_tracked: more than 2^32-1 hashes in one call")
            base = sum(ns_l[:self.rank])
            req, owner = self._requests(values)
            ords = (base + torch.arange(m, dtype=torch.int64, device=self.device)).repeat_interleave(T).to(torch.int32)
            handle, got = self._router().route(req, owner)
            got_ord = self._router().ship(handle, ords)
            first = torch.zeros(max(got.numel(), 1), dtype=torch.uint8, device=self.device)
            _capi.check(_capi.lib().gt_shard_insert_requests_dev(self._h, got.data_ptr(), got_ord.data_ptr(), got.numel(),
                                                                 first.data_ptr()), "gt_shard_insert_requests_dev")
            back = self._router().back(handle, first[:got.numel()])
            is_new = back.view(m, T).amax(dim=1) if m else torch.zeros(0, dtype=torch.uint8, device=self.device)
            tot = is_new.sum(dtype=torch.int64).reshape(1)
            dist.all_reduce(tot, group=self.group)
            self._n_unique = getattr(self, "_n_unique", 0) + int(tot.item())
        self.stream.synchronize()
        return is_new

    def insert_sequences_tracked(self, shifter_kind, K, bases, offsets):
        """dBG::insert_sequence(sequence, n_new) over a host batch on the sharded tables (collective): -> (k-mers
        consumed, n_new per read) under the serial rule, ranks in order."""
        torch = self.torch
        bases, offsets = _capi.as_reads(bases, offsets)
        vals = self._hash_values_dev(shifter_kind, K, bases, offsets)
        is_new = self.insert_hashes_tracked_dev(vals)
        n = offsets.size - 1
        lens = (offsets[1:] - offsets[:-1]).astype(np.int64)
        ok = np.ones(n, dtype=bool)
        if n:
            valid = np.zeros(256, dtype=bool)
            valid[np.frombuffer(b"ACGTacgt", dtype=np.uint8)] = True
            bad = np.concatenate([[0], np.cumsum(~valid[bases[int(offsets[0]):int(offsets[-1])]])])
            rel = (offsets - offsets[0]).astype(np.int64)
            ok = (bad[rel[1:]] - bad[rel[:-1]]) == 0
        nk = np.where(ok, np.maximum(lens - K + 1, 0), 0)
        flags = is_new.cpu().numpy().astype(np.int64)
        ends = np.cumsum(nk)
        csum = np.concatenate([[0], np.cumsum(flags)])
        n_new = csum[ends] - csum[ends - nk]
        return int(nk.sum()), n_new.astype(np.uint64)

    def n_unique_kmers(self):
        """Running sum of the tracked inserts' is_new over all ranks (blind inserts do not maintain it)."""
        return int(getattr(self, "_n_unique", 0))

    # ---- OXLI v4 files of the WHOLE storage (bitstorage.cc:151-307, bytestorage.cc:480-536, nibblestorage.cc:132-278) ---
    def _file_layout(self):
        kind = self.kind
        head = 6 + (1 if kind == _capi.STORAGE_BYTE else 0) + 13
        starts, pos = [], head
        for size in self._sizes.tolist():
            size = int(size)
            starts.append(pos + 8)
            pos += 8 + (size // 8 + 1 if kind == 0 else size if kind == 1 else size // 2 + 1)
        return head, starts, pos

    def save(self, filename, ksize):
        """Collective: one OXLI file holding the whole storage, byte-identical to what a single-GPU storage (and the
        reference) writes for the same contents: rank 0 writes the header, every rank writes its slot ranges in place."""
        import struct
        kind, n_occ = self.kind, self.n_occupied()
        head, starts, total = self._file_layout()
        parts = self.local_tables()
        if self.rank == 0:
            with open(filename, "wb") as f:
                f.write(b"OXLI")
                f.write(struct.pack("<BB", 4, {0: 2, 1: 1, 2: 7}[kind]))
                if kind == _capi.STORAGE_BYTE:
                    f.write(struct.pack("<B", 0))
                f.write(struct.pack("<IBQ", int(ksize), int(self._sizes.size), n_occ))
                f.truncate(total + (8 if kind == _capi.STORAGE_BYTE else 0))
        self.dist.barrier(group=self.group)
        spb = {0: 8, 1: 1, 2: 2}[kind]
        with open(filename, "r+b") as f:
            for i, size in enumerate(self._sizes.tolist()):
                lo, _ = self.local_range(i)
                if self.rank == 0:
                    f.seek(starts[i] - 8)
                    f.write(struct.pack("<Q", int(size)))
                if parts[i].size:
                    f.seek(starts[i] + lo // spb)
                    f.write(parts[i].tobytes())
            if self.rank == 0 and kind == _capi.STORAGE_BYTE:
                f.seek(total)
                f.write(struct.pack("<Q", 0))  # n_bigcounts
        self.dist.barrier(group=self.group)

    def load(self, filename):
        """Collective: every rank reads its slot ranges of an OXLI file (written by any goetia storage of the same kind and
        table sizes) into its part.  Returns the saved ksize."""
        import struct
        kind = self.kind
        head, starts, total = self._file_layout()
        spb = {0: 8, 1: 1, 2: 2}[kind]
        L = _capi.lib()
        self.synchronize()
        with open(filename, "rb") as f:
            hdr = f.read(head)
            if len(hdr) < head or hdr[:4] != b"OXLI":
                raise _capi.GoetiaB200Error("Does not start with signature for a oxli binary file")
            if hdr[4] != 4 or hdr[5] != {0: 2, 1: 1, 2: 7}[kind]:
                raise _capi.GoetiaB200Error("Incorrect file format version / type")
            ksize, n_tables, _occ = struct.unpack_from("<IBQ", hdr, head - 13)
            if n_tables != self._sizes.size:
                raise _capi.GoetiaB200Error("table count differs from this storage's")
            for i, size in enumerate(self._sizes.tolist()):
                f.seek(starts[i] - 8)
                (fsize,) = struct.unpack("<Q", f.read(8))
                if fsize != int(size):
                    raise _capi.GoetiaB200Error("table sizes differ from this storage's")
                lo, _ = self.local_range(i)
                nb = int(L.gt_storage_table_bytes(self._h, i))
                f.seek(starts[i] + lo // spb)
                buf = np.frombuffer(f.read(nb), dtype=np.uint8)
                if buf.size != nb:
                    raise _capi.GoetiaB200Error("truncated table")
                buf = np.ascontiguousarray(buf)
                _capi.check(L.gt_storage_upload_table(self._h, i, buf.ctypes.data), "gt_storage_upload_table")
        self._n_unique = 0
        self.dist.barrier(group=self.group)
        return int(ksize)

    def local_range(self, i):
        lo, hi = np.zeros(1, dtype=np.uint64), np.zeros(1, dtype=np.uint64)
        _capi.check(_capi.lib().gt_storage_local_range(self._h, i, lo.ctypes.data_as(_capi.u64p),
                                                       hi.ctypes.data_as(_capi.u64p)), "gt_storage_local_range")
        return int(lo[0]), int(hi[0])

    def local_tables(self):
        """Host copies of this rank's parts; concatenated in rank order they are the reference's tables."""
        L = _capi.lib()
        self.synchronize()
        out = []
        for i in range(self._sizes.size):
            buf = np.empty(int(L.gt_storage_table_bytes(self._h, i)), dtype=np.uint8)
            _capi.check(L.gt_storage_download_table(self._h, i, buf.ctypes.data), "gt_storage_download_table")
            out.append(buf)
        return out

    def checksum(self, i):
        """Checksum of the WHOLE table i: the ranks' part checksums add up (gt_storage_checksum is linear)."""
        out = C.c_uint64(0)
        self.synchronize()
        _capi.check(_capi.lib().gt_storage_checksum(self._h, int(i), C.byref(out)), "gt_storage_checksum")
        v = int(out.value)
        t = self.torch.tensor([v & 0xFFFFFFFF, v >> 32], dtype=self.torch.int64, device=self.device)
        self.dist.all_reduce(t, group=self.group)  # 32-bit halves: NCCL has no modular uint64 sum we can rely on
        lo, hi = int(t[0].item()), int(t[1].item())
        return (lo + (hi << 32)) & 0xFFFFFFFFFFFFFFFF

    def n_occupied_local(self):
        a = np.zeros(2, dtype=np.uint64)
        self.synchronize()
        _capi.check(_capi.lib().gt_storage_stats(self._h, a[0:].ctypes.data_as(_capi.u64p),
                                                 a[1:].ctypes.data_as(_capi.u64p)), "gt_storage_stats")
        return int(a[1])

    def n_occupied(self):
        t = self.torch.tensor([self.n_occupied_local()], dtype=self.torch.int64, device="cuda")
        self.dist.all_reduce(t, group=self.group)
        return int(t.item())

    def pending_info(self):
        a = np.zeros(8, dtype=np.uint64)
        _capi.check(_capi.lib().gt_storage_pending_info(self._h, a.ctypes.data), "gt_storage_pending_info")
        keys = ("built", "n_buckets", "slice_shift", "budget_kmers", "entries", "pending_kmers", "n_direct",
                "apply_grid")
        return dict(zip(keys, (int(v) for v in a)))

    def reset(self):
        self.synchronize()
        _capi.check(_capi.lib().gt_storage_reset(self._h), "gt_storage_reset")
        self._n_unique = 0

    def close(self, collective=True):
        """Collective when the transport is p2p or ce (barriers keep a peer from storing into, or
        holding a mapping of, memory that is going away)."""
        if getattr(self, "_h", None):
            L = _capi.load()
            L.gt_synchronize()
            if self.transport in ("p2p", "ce") and collective:
                # nobody may still be storing into a mapping that is about to go away
                try:
                    self.dist.barrier(group=self.group)
                except Exception:
                    pass
            L.gt_set_compute_stream(None)
            L.gt_set_apply_stream(None)
            L.gt_storage_destroy(self._h)
            for p in self._peer_ptrs:
                L.gt_peer_close(p)
            if self.transport in ("p2p", "ce") and collective:
                try:
                    self.dist.barrier(group=self.group)  # peers have unmapped before the memory is freed
                except Exception:
                    pass
            for p in self._own_inbox:
                L.gt_peer_free(p)
            self._peer_ptrs, self._own_inbox = [], []
            self._h = None

    def __del__(self):
        try:
            self.close(collective=False)
        except Exception:
            pass
