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
import numpy as np

from . import _capi

MAX_BUCKETS = 1024


class ShardPlan:
    """gt_shard_plan: slices, owners, capacities.  Identical on every rank."""

    def __init__(self, kind, sizes, world, budget_kmers, slice_log2_bytes=0):
        L = _capi.load()
        self.kind, self.world, self.budget_kmers = int(kind), int(world), int(budget_kmers)
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


class ShardedStorage:
    """This rank's part of a BitStorage / ByteStorage / NibbleStorage sharded over the process group."""

    def __init__(self, kind, sizes, budget_kmers, group=None, slice_log2_bytes=0):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        L = _capi.lib()
        self.kind = int(kind)
        self.plan = ShardPlan(kind, sizes, self.world, budget_kmers, slice_log2_bytes)
        self._sizes = self.plan.sizes
        self._h = L.gt_storage_create_sharded(self.kind, self._sizes.ctypes.data_as(_capi.u64p), self._sizes.size,
                                              self.rank, self.world, int(budget_kmers), int(slice_log2_bytes))
        if not self._h:
            raise _capi.GoetiaB200Error("gt_storage_create_sharded: " + _capi.last_error())
        self.stream = torch.cuda.Stream()
        _capi.check(L.gt_set_compute_stream(self.stream.cuda_stream), "gt_set_compute_stream")
        self.x = ShardExchange(self.plan, self.rank, torch, torch.device("cuda", torch.cuda.current_device()), group)
        _capi.check(L.gt_storage_attach_exchange(self._h, 0, self.x.outbox.data_ptr(), self.x.inbox.data_ptr(),
                                                 self.x.fill_send.data_ptr(), self.x.fill_recv.data_ptr()),
                    "gt_storage_attach_exchange")

    @property
    def handle(self):
        return self._h

    def bucket_sequences_dev(self, shifter_kind, K, d_bases_ptr, d_offsets_ptr, n_reads, n_bases):
        """Step 1 of a round: pack + hash + bucket this rank's (device-resident) reads."""
        with self.torch.cuda.stream(self.stream):
            return int(_capi.check(_capi.lib().gt_insert_sequences_dev(self._h, shifter_kind, K, d_bases_ptr, d_offsets_ptr,
                                                                       n_reads, n_bases, _capi.MODE_BLIND),
                                   "gt_insert_sequences_dev"))

    def exchange_and_apply(self):
        """Steps 2 and 3: all-to-all of counts and bucket regions, then apply this rank's slices."""
        with self.torch.cuda.stream(self.stream):
            self.x.exchange()
            _capi.check(_capi.lib().gt_storage_apply(self._h), "gt_storage_apply")

    def synchronize(self):
        self.stream.synchronize()

    def local_range(self, i):
        lo, hi = np.zeros(1, dtype=np.uint64), np.zeros(1, dtype=np.uint64)
        _capi.check(_capi.lib().gt_storage_local_range(self._h, i, lo.ctypes.data_as(_capi.u64p),
                                                       hi.ctypes.data_as(_capi.u64p)), "gt_storage_local_range")
        return int(lo[0]), int(hi[0])

    def local_tables(self):
        """Host copies of this rank's parts; concatenated in rank order they are the reference's tables."""
        L = _capi.lib()
        out = []
        for i in range(self._sizes.size):
            buf = np.empty(int(L.gt_storage_table_bytes(self._h, i)), dtype=np.uint8)
            _capi.check(L.gt_storage_download_table(self._h, i, buf.ctypes.data), "gt_storage_download_table")
            out.append(buf)
        return out

    def n_occupied_local(self):
        a = np.zeros(2, dtype=np.uint64)
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
        _capi.check(_capi.lib().gt_storage_reset(self._h), "gt_storage_reset")

    def close(self):
        if getattr(self, "_h", None):
            L = _capi.load()
            L.gt_set_compute_stream(None)
            L.gt_storage_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
