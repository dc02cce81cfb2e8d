#!/usr/bin/env python
"""bench.py -- k-mers inserted per second on the BASELINE.json workload (one JSON line).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c3|c1|c2|c2q|c4|c5|storage]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (2-bit pack -> canonical rolling hash -> insert into the
probabilistic tables) over the whole synthetic read set of the workload.  The default workload
is BASELINE.json configs[2] -- the one the metric's target ("BitStorage K=31 insert") and its
1/2/4/8-GPU scaling are quoted on, and it fits one B200: dBG<BitStorage, CanLemireShifter>,
K=31, 4 tables x ~8e9 bits (get_n_primes_near_x(4, 8e9)), 50 M synthetic 150 bp reads
(6.0e9 k-mers per step) from a counter-based generator, so the CPU reference, one GPU and every
rank of N see the same read set.

  value : device-timed (CUDA events on the library's compute stream), reads resident in HBM as
          ASCII + offsets when the timed region starts.
  e2e   : the same step through the host-buffer C-ABI call on the parser's product, pinned 2-bit
          packed reads (gt_insert_sequences_packed; N > 1: ShardedStorage), H2D copies inside the
          timed region, k-mer count read back, wall clock; e2e_ascii: gt_insert_sequences on
          pinned ASCII.
  check : the tables' checksums and occupancy against those of the compiled reference over the
          same reads (tests/golden/fullsize_golden.json), after the timed and after the e2e steps.
  roofline      : algorithmic 256 B/k-mer (4 x (32 B sector read + 32 B write-back), SURVEY.md
                  section 8d) over the whole timed region; per-kernel event times, measured DRAM
                  traffic (profiles/traffic.json), dram_frac and the random-sector rate probed in
                  the run ride along.
  cpu_baseline  : the unmodified reference (oracle/_ref, compiled from /root/reference) -- or
                  the plain-C port when that .so is absent -- on the box's host cores over a
                  bounded sample of the same reads.
  N > 1 : reads sharded over the ranks, tables partitioned by slot range, foreign buckets staged
          locally and shipped by the copy engines over NVLink (goetia_b200/shard.py, transport
          "ce"; GT_SHARD_TRANSPORT=p2p|nccl select the others).

--impl reference times that CPU implementation as its own arm (rank 0 only).
The oracle is only ever the checker / the CPU baseline here, never the measured product path.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

ALGO_BYTES_PER_KMER_PER_TABLE = 64  # 32 B sector read + 32 B write-back

WORKLOADS = {
    # name: (storage kind, K, table x, n_tables, total reads, read length, seed, description)
    "c3": (0, 31, int(8e9), 4, 50_000_000, 150, 44,
           "C3: dBG<BitStorage,CanLemireShifter> K=31, 4 tables x 8e9 bits, 50M x 150bp synthetic reads"),
    "c1": (0, 31, int(1e9), 4, 100_000, 150, 42,
           "C1: dBG<BitStorage,CanLemireShifter> K=31, BitStorage(1e9,4), 100k x 150bp synthetic reads"),
    "c2": (1, 21, int(4e9), 4, 20_000_000, 150, 43,
           "C2: dBG<ByteStorage,CanLemireShifter> K=21, ByteStorage(4e9,4), 20M x 150bp synthetic reads (insert)"),
    "c5": (2, 25, int(8e9), 4, 1_000_000, 10_000, 46,
           "C5: dBG<NibbleStorage,CanLemireShifter> K=25, NibbleStorage(8e9,4), 1M x 10kb synthetic reads"),
    # the query half of C2: insert once (untimed), then time the per-read median-count query over the same reads
    "c2q": (1, 21, int(4e9), 4, 20_000_000, 150, 43,
            "C2 query: dBG<ByteStorage,CanLemireShifter> K=21, ByteStorage(4e9,4), per-read median-count query "
            "(DiginormFilter::median_count_at_least, cutoff 1) over 20M x 150bp synthetic reads after inserting them"),
    # sketch workload (kind -1): not the headline metric; `--workload c4` reports k-mers sketched per second
    "c4": (-1, 31, 1000, 0, 100_000_000, 150, 45,
           "C4: SourmashSketch K=31 scaled=1000 streaming sketch of 100M x 150bp synthetic reads"),
}
SUB_BATCH_BASES = int(os.environ.get("GT_BENCH_SUB_BASES", 900_000_000))  # reads are fed to the library in sub-batches of about this many bases


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS) + ["storage"])
    ap.add_argument("--reads", type=int, default=0, help="override the workload's read count (smoke runs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 20 ms; only the samples taken between begin() and
    stop() -- the timed region -- are reported."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines, self.t0 = index, None, [], None

    def start(self):
        """Launch the sampler and wait (<= 3 s) until it delivers, so that a short timed region is covered."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            t_end = time.time() + 3.0
            while not self.lines and time.time() < t_end:
                time.sleep(0.01)
        except Exception:
            self.proc = None
        self.begin()

    def begin(self):
        self.t0 = time.time()

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self):
        if not self.proc:
            return None
        t1 = time.time()
        time.sleep(0.03)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        inside = [ln for (t, ln) in self.lines if self.t0 <= t <= t1 + 0.03]
        if not inside:  # region shorter than one sampling period: the nearest sample
            inside = [ln for (_, ln) in self.lines[-1:]]
        sm, mx, reasons = [], [], set()
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def synth_reads_device(torch, L, first_read, n_reads, read_len, seed, device):
    """Reads [first_read, first_read + n_reads) of the counter-based synthetic stream `seed` (goetia_b200/synth.py is the
    numpy twin; the CPU reference and every rank see the same bytes), written on the device by gt_synth_bases_dev."""
    from goetia_b200 import _capi
    out = torch.empty(max(n_reads, 1) * read_len + 16, dtype=torch.uint8, device=device)
    torch.cuda.synchronize()
    _capi.check(L.gt_synth_bases_dev(out.data_ptr(), n_reads * read_len, seed, first_read * read_len), "gt_synth_bases_dev")
    L.gt_synchronize()
    return out


def nvlink_counters(index):
    """Sum of this GPU's NVLink data counters (nvidia-smi nvlink -gt d): {"tx_bytes", "rx_bytes"} or None."""
    import re
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(index)], capture_output=True, text=True, timeout=20).stdout
        tx = [int(x) for x in re.findall(r"Data Tx:\s*(\d+)\s*KiB", out)]
        rx = [int(x) for x in re.findall(r"Data Rx:\s*(\d+)\s*KiB", out)]
        if not tx and not rx:
            return None
        return {"tx_bytes": sum(tx) * 1024, "rx_bytes": sum(rx) * 1024, "links": len(tx)}
    except Exception:
        return None


def golden_entry(workload, reads, total_default):
    """Reference-pinned table checksums of the full-size workload (tests/golden/fullsize_golden.json, made by
    tests/golden/make_fullsize_golden.py from the compiled, unmodified reference) -- only for the unmodified read set."""
    try:
        with open(os.path.join(ROOT, "tests", "golden", "fullsize_golden.json")) as f:
            g = json.load(f)
    except Exception:
        return None
    if reads == total_default:
        return g.get(workload)
    for key, e in g.items():  # a pinned prefix of the read set (--reads N): e.g. c5_first_125k
        if key.startswith(workload + "_first_") and int(e.get("reads", -1)) == int(reads) and int(e.get("first_read", 0)) == 0:
            return e
    return None


def checksum_check(gold, sums, n_occ):
    """check block: the tables' position-weighted checksums (computed in HBM, all ranks summed) against the reference's."""
    if gold is None:
        return {"tables_checksum_equal_reference": None, "note": "no golden entry for this read count / workload"}
    ok = [int(a) == int(b) for a, b in zip(sums, gold["checksums"])]
    return {"tables_checksum_equal_reference": bool(all(ok)) and int(n_occ) == int(gold["n_occupied"]),
            "n_occupied_reference": int(gold["n_occupied"]), "checksums": [int(x) for x in sums],
            "reference": "tests/golden/fullsize_golden.json (oracle/_ref, the unmodified reference, %d thread(s), same counter-based reads)"
                         % gold.get("threads", 1)}


def measured_r_rand(L, what, footprint_bytes):
    """The random-sector roofline, measured in this run on this device (gt_probe_random)."""
    v = L.gt_probe_random(what, int(footprint_bytes), 1 << 28)
    return float(v) if v > 0 else None


def cpu_reference_modes(kind, K, sizes, threads):
    """(impl, kind name, [thread counts to try]).  The unmodified reference when oracle/_ref was
    built; its BitStorage is exact under threads (atomic OR), so both of BASELINE.md's modes are
    tried: A = one thread (the reference's own execution model), B = one dBG copy per host core
    over one shared storage."""
    from oracle import binding
    if binding.have_ref():
        return binding.Ref(kind, 1, K, sizes), "reference", ([1, threads] if (kind == 0 and threads > 1) else [1])
    return binding.Port(kind, 1, K, sizes), "port", [1]


def cpu_time_reads(impl, kname, bases, offsets, r0, r1, n_threads):
    b = bases[int(offsets[r0]):int(offsets[r1])]
    o = offsets[r0:r1 + 1] - offsets[r0]
    if kname == "reference":
        return impl.insert_reads(b, o, n_threads=n_threads)
    return impl.insert_reads(b, o)


def cpu_reference_rate(kind, K, sizes, bases, offsets, budget_s, threads):
    """Time the CPU implementation on bounded, never-before-inserted slices of the reads (so
    k-mers are new, as in a real first pass).  Returns dict(value, cores, kind, sample)."""
    n_total = offsets.size - 1
    impl, kname, modes = cpu_reference_modes(kind, K, sizes, threads)
    probe = min(n_total // 4, 100_000)
    nk, secs = cpu_time_reads(impl, kname, bases, offsets, 0, probe, modes[-1])  # faults the tables' pages in
    rate = nk / max(secs, 1e-9)
    kpr = max(1.0, nk / max(1, probe))
    best, notes, r0 = None, [], probe
    for t in modes:
        n = int(min((n_total - r0) // (len(modes) - modes.index(t)), max(20_000, rate * budget_s / len(modes) / kpr)))
        if n <= 0:
            break
        nk, secs = cpu_time_reads(impl, kname, bases, offsets, r0, r0 + n, t)
        r0 += n
        v = nk / secs
        notes.append("%d thread(s): %d reads, %d k-mers, %.1f s -> %.3g k-mers/s" % (t, n, nk, secs, v))
        if best is None or v > best[0]:
            best = (v, t)
    impl.close()
    return {"value": best[0], "unit": "k-mers/s", "cores": best[1], "kind": kname,
            "sample": "fresh slices of the same synthetic read set, dBG::insert_sequence per read into pre-faulted "
                      "tables of the full size; " + "; ".join(notes) + "; value = the faster mode"}


def main():
    args = parse_args()
    if args.workload == "storage":
        return storage_arm(args)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    kind, K, x, n_tables, total_reads, read_len, seed, desc = WORKLOADS[args.workload]
    if args.reads:
        total_reads = args.reads
    kpr = read_len - K + 1

    if args.impl == "reference":
        return reference_arm(args, rank, world, kind, K, x, n_tables, total_reads, read_len, seed, desc)

    import torch
    import goetia_b200 as gb
    from goetia_b200 import _capi

    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    gb.init(local_rank)
    L = _capi.lib()
    sizes = gb.get_n_primes_near_x(n_tables, x) if kind >= 0 else []

    if kind < 0:
        return sketch_arm(args, rank, world, local_rank, torch, gb, _capi)
    if world > 1:
        return multi_gpu_arm(args, rank, world, local_rank, torch, gb, _capi, sizes, total_reads)

    # ---- inputs: resident ASCII sub-batches ------------------------------------------------------
    total_default = WORKLOADS[args.workload][4]
    gold = golden_entry("c2" if args.workload == "c2q" else args.workload, total_reads, total_default)
    reads_per_sub = max(1, min(total_reads, SUB_BATCH_BASES // read_len))
    reads_per_sub -= reads_per_sub % 16 if reads_per_sub > 16 else 0
    subs = []
    r = 0
    while r < total_reads:
        n = min(reads_per_sub, total_reads - r)
        subs.append((synth_reads_device(torch, L, r, n, read_len, seed, dev), n))
        r += n
    offs = torch.arange(reads_per_sub + 1, dtype=torch.int64, device=dev) * read_len
    torch.cuda.synchronize()
    storage = [gb.BitStorage, gb.ByteStorage, gb.NibbleStorage][kind](sizes)
    graph = gb.dBG[type(storage), gb.CanLemireShifter].build(storage, K)
    if args.workload == "c2q":
        return query_arm(args, torch, gb, _capi, L, dev, local_rank, storage, graph, subs, offs, sizes, gold, K, read_len,
                         total_reads, seed, desc)
    # counting storages start every step from empty tables (a saturated table would hide the first-pass cost);
    # a Bloom table is idempotent, so C3 / C1 keep accumulating
    reset_each_step = kind != 0

    d_total = torch.zeros(1, dtype=torch.int64, device=dev)  # k-mers consumed, accumulated on the device
    torch.cuda.synchronize()

    def step_resident():
        # nothing in here waits for the GPU until the flush at the end of the step
        if reset_each_step:
            storage.reset()
        for b, n in subs:
            # equal-length reads: the exact k-mer count rides along, so three sub-batches (not two) fit a pending store
            graph.insert_sequences_dev_async(b.data_ptr(), offs.data_ptr(), n, n * read_len, mode=gb.MODE_BLIND,
                                             d_kmer_total_ptr=d_total.data_ptr(), n_kmers_upper=n * kpr)
        storage.flush()

    def step_checked():
        d_total.zero_()
        torch.cuda.synchronize()
        step_resident()
        return int(d_total.item())

    kmers_per_step = total_reads * kpr
    for _ in range(args.warmup):
        assert step_checked() == kmers_per_step
    L.gt_synchronize()
    _capi.check(L.gt_profile_enable(1), "gt_profile_enable")
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = L.gt_launch_count()
    _capi.check(L.gt_timer_record(0), "gt_timer_record")
    for _ in range(args.steps):
        step_resident()
    _capi.check(L.gt_timer_record(1), "gt_timer_record")
    ms = L.gt_timer_elapsed_ms(0, 1)
    L.gt_synchronize()
    launches = int(L.gt_launch_count() - launches0)
    clocks = sampler.stop()
    prof_ms = np.zeros(5, dtype=np.float64)
    prof_n = np.zeros(5, dtype=np.uint64)
    _capi.check(L.gt_profile_get_detail(prof_ms.ctypes.data, prof_n.ctypes.data, 5), "gt_profile_get_detail")
    L.gt_profile_enable(0)
    value = kmers_per_step * args.steps / (ms / 1e3)

    # ---- parity at full size: table checksums against the compiled reference's (same reads) -----------------
    info = storage.pending_info()
    n_occ = storage.n_occupied()
    sums = [storage.checksum(i) for i in range(n_tables)]
    check = checksum_check(gold, sums, n_occ)
    sample_b = subs[0][0][:2000 * read_len].cpu().numpy()
    sample_o = np.arange(2001, dtype=np.uint64) * np.uint64(read_len)
    q = graph.query_sequences(sample_b, sample_o)
    check["sample_reads_all_present"] = bool((q >= 1).all())
    check["n_occupied"] = n_occ

    # ---- e2e: host buffers through the C ABI ------------------------------------------------------
    # The parsing-to-device pipeline's product is a pinned 2-bit packed batch (the FASTX front end's parser threads pack;
    # north_star: "FastxParser -> pinned 2-bit buffers on CUDA streams"), so the headline e2e step is
    # gt_insert_sequences_packed on packed pinned host buffers: 0.25 B/base + offsets + flags cross PCIe inside the timed
    # region.  The same step from pinned ASCII (gt_insert_sequences, 1 B/base, packed on the device) rides along as e2e_ascii.
    e2e = None
    host_b = None
    if not args.no_e2e:
        from goetia_b200.batch import pack_reads_host
        host_b = torch.empty(total_reads * read_len, dtype=torch.uint8, pin_memory=True)
        p = 0
        for b, n in subs:
            host_b[p:p + n * read_len].copy_(b[:n * read_len])
            p += n * read_len
        host_o = torch.empty(total_reads + 1, dtype=torch.int64, pin_memory=True)
        host_o.copy_(torch.arange(total_reads + 1, dtype=torch.int64) * read_len)
        host_w = torch.zeros((total_reads * read_len + 31) // 32 + 1, dtype=torch.int64, pin_memory=True)
        host_f = torch.zeros(total_reads, dtype=torch.uint8, pin_memory=True)
        torch.cuda.synchronize()
        hb, ho = host_b.numpy(), host_o.numpy().view(np.uint64)
        hw, hf = host_w.numpy().view(np.uint64), host_f.numpy()
        t0 = time.perf_counter()
        pack_reads_host(hb, ho, 0, hw, hf)  # what the parser threads do; outside the timed region, reported below
        pack_s = time.perf_counter() - t0
        os.environ.setdefault("GT_CHUNK_BASES", str(256 << 20))

        def step_packed():
            if reset_each_step:
                storage.reset()
            nk = graph.insert_sequences_packed(hw, ho, hf, mode=gb.MODE_BLIND)  # the k-mer count is read back from the device
            storage.flush()
            return nk

        def step_ascii():
            if reset_each_step:
                storage.reset()
            nk = graph.insert_sequences(hb, ho, mode=gb.MODE_BLIND)
            storage.flush()
            return nk

        def timed(step, steps):
            for _ in range(min(args.warmup, 2)):
                assert step() == kmers_per_step
            L.gt_synchronize()
            t0 = time.perf_counter()
            for _ in range(steps):
                step()
            L.gt_synchronize()
            return (time.perf_counter() - t0) / steps

        dt = timed(step_packed, args.steps)
        dta = timed(step_ascii, max(1, args.steps // 2))
        e2e = {"value": kmers_per_step / dt, "unit": "k-mers/s",
               "h2d_bytes_per_step": int(hw.nbytes + ho.nbytes + hf.nbytes), "d2h_bytes_per_step": 8, "ms_per_step": dt * 1e3,
               "api": "gt_insert_sequences_packed(pinned 2-bit words, offsets, flags) + gt_storage_flush",
               "host_pack": {"seconds": pack_s, "bases_per_s": total_reads * read_len / pack_s, "threads": os.cpu_count(),
                             "note": "gt_pack_reads_host, once, outside the timed region (the parser threads' share of the pipeline)"},
               "e2e_ascii": {"value": kmers_per_step / dta, "ms_per_step": dta * 1e3, "h2d_bytes_per_step": int(hb.nbytes + ho.nbytes),
                             "api": "gt_insert_sequences(pinned ASCII, offsets) + gt_storage_flush (packed on the device)"}}
        check["e2e_tables_checksum_equal_reference"] = (None if gold is None else
                                                        [storage.checksum(i) for i in range(n_tables)] == [int(x) for x in gold["checksums"]])

    # ---- roofline -----------------------------------------------------------------------------------
    peak, peak_src = measured_peak()
    table_bytes = sum(storage.table_bytes(i) for i in range(n_tables))
    r_rand = measured_r_rand(L, 0, table_bytes)
    roofline = insert_roofline(prof_ms, prof_n, ms, kmers_per_step * args.steps, n_tables, peak, peak_src, args.workload, r_rand,
                               table_bytes)

    # ---- CPU baseline ---------------------------------------------------------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        from goetia_b200.synth import synth_reads
        nb = min(total_reads, 4_000_000)
        cb, co = synth_reads(seed, 0, nb, read_len)  # the same bytes the GPU path was fed
        del storage, graph
        cpu = cpu_reference_rate(kind, K, sizes, cb, co, budget_s=15.0, threads=os.cpu_count() or 1)

    out = {
        "metric": "k-mers inserted/sec (device-timed)", "value": value, "unit": "k-mers/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": desc if not args.reads else desc + " [reads overridden: %d]" % total_reads,
                   "K": K, "tablesizes": sizes, "reads": total_reads, "read_len": read_len,
                   "kmers_per_step": kmers_per_step, "mode": "GT_MODE_BLIND (write-combined)",
                   "sub_batches": len(subs), "slice_shift": info["slice_shift"], "n_buckets": info["n_buckets"],
                   "pending_entries": info["entries"], "bucket_overflow_updates": info["n_direct"],
                   "seed": seed, "generator": "counter-based splitmix64 (gt_synth_bases_dev / goetia_b200/synth.py)",
                   "l2": "inputs larger than L2 (%.1f GB reads + %.1f GB tables + %.1f GB update store per step)"
                         % (total_reads * read_len / 1e9, table_bytes / 1e9, info["entries"] * 4 / 1e9),
                   "tables": "reset at the start of every step (counting storage: every step is a first pass)" if reset_each_step
                             else "accumulate across steps (a Bloom table is idempotent; the work per step is the same)"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
        "check": check,
    }
    print(json.dumps(out))
    return 0


def insert_roofline(prof_ms, prof_n, step_ms_total, kmers, n_tables, peak, peak_src, workload, r_rand=None, table_bytes=0):
    """HBM roofline of the insert (64 B algorithmic per (k-mer, table)).

    On the write-combined path the insert is three kernels: k_bucket (hash + partition by 32 MB table slice) ->
    k_rebucket2 (partition a slice's entries by 128 KB window) -> k_apply_win (build the window in shared memory,
    merge it into the table with one coalesced read-modify-write).  The apply of one entry store is queued on a
    second stream beside k_bucket filling the other store; all three kernels are persistent / SM-filling, so their
    CUDA-event times overlap and inflate each other.  `achieved` is therefore the conservative figure: algorithmic
    bytes over the WHOLE timed region (= value x 256 B); the per-kernel event times are listed under `kernels`.
    `traffic` is the DRAM bytes ncu measured for one apply cycle (profiles/traffic.json: two k_bucket launches + one
    k_rebucket2 + one k_apply_win = 103 GB per 1.33e9 k-mers = 77 B/k-mer) -- far below the 256 algorithmic bytes, which
    is the point of write-combining and why `frac` can exceed 1; `dram_frac` is what the path really draws from HBM
    (measured bytes per k-mer x k-mers/s over the copy peak), see DESIGN.md section 3."""
    names = ("k_bucket", "k_apply (k_rebucket + k_apply_win)", "k_walk", "k_rebucket", "k_apply_win")
    algo_bytes = ALGO_BYTES_PER_KMER_PER_TABLE * n_tables
    combined = bool(prof_n[1])
    achieved = kmers * algo_bytes / (step_ms_total / 1e3) / 1e9 if step_ms_total > 0 else 0.0
    traffic = detail = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            detail = json.load(f).get(workload)
        traffic = float(detail["pair"]) if combined else float(detail.get("k_walk")) if detail.get("k_walk") else None
    except Exception:
        pass
    out = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
           "traffic": traffic, "peak_source": peak_src,
           "kernel": "k_bucket + k_rebucket + k_apply_win (write-combined insert, two streams)" if combined else "k_walk",
           "algorithmic_bytes_per_kmer": algo_bytes,
           "kernels": {name: {"launches": int(prof_n[i]), "ms_total": float(prof_ms[i]),
                              "ms_per_launch": float(prof_ms[i] / max(1, int(prof_n[i]))),
                              "share_of_step": float(prof_ms[i] / step_ms_total) if step_ms_total > 0 else 0.0,
                              "achieved_alone": kmers * algo_bytes / (float(prof_ms[i]) / 1e3) / 1e9 if prof_ms[i] > 0 else None}
                       for i, name in enumerate(names) if prof_n[i]}}
    if detail and combined and detail.get("kmers_per_launch"):
        # measured DRAM bytes per k-mer x k-mers/s over the HBM peak: how much of the HBM the path really uses
        per_kmer = float(detail["pair"]) / float(detail["kmers_per_launch"])
        out["traffic_detail"] = detail
        out["dram_bytes_per_kmer_measured"] = per_kmer
        out["dram_frac"] = per_kmer * kmers / (step_ms_total / 1e3) / 1e9 / peak if step_ms_total > 0 else None
    # the random-sector roofline north_star's ">= 50 %" refers to: independent random 32-bit RED.OR over a footprint
    # equal to the tables', MEASURED IN THIS RUN on this device (gt_probe_random; scripts/microbench.cu is the long form)
    if r_rand:
        out["random_sector_roofline"] = {"r_rand_atomics_per_s": r_rand, "measured": "in this run, footprint = the tables' %.1f GB" % (table_bytes / 1e9),
                                         "achieved_updates_per_s": kmers * n_tables / (step_ms_total / 1e3) if step_ms_total > 0 else 0.0,
                                         "frac": kmers * n_tables / (step_ms_total / 1e3) / r_rand if step_ms_total > 0 else 0.0}
    out["traffic_source"] = "ncu capture by the builder, committed under profiles/ (see profiles/traffic.json)"
    return out


def query_arm(args, torch, gb, _capi, L, dev, local_rank, storage, graph, subs, offs, sizes, gold, K, read_len, total_reads, seed, desc):
    """`--workload c2q`: the query half of BASELINE.json configs[1] -- DiginormFilter::median_count_at_least per read
    (diginorm.hh:35-68 over dBG::query_sequence, dbg.hh:349-362) on the C2 graph.  Reads are inserted once (untimed,
    tables checked against the reference's checksums); a step = one pass of the median-count query over all reads.
    Algorithmic bytes: n_tables x 32 B sector reads = 128 B per k-mer."""
    n_tables = len(sizes)
    kpr = read_len - K + 1
    kmers_per_step = total_reads * kpr
    for b, n in subs:
        graph.insert_sequences_dev_async(b.data_ptr(), offs.data_ptr(), n, n * read_len, mode=gb.MODE_BLIND)
    storage.flush()
    n_occ = storage.n_occupied()
    check = checksum_check(gold, [storage.checksum(i) for i in range(n_tables)], n_occ)
    d_total = torch.zeros(1, dtype=torch.int64, device=dev)
    d_pass = [torch.zeros(n, dtype=torch.uint8, device=dev) for _, n in subs]
    torch.cuda.synchronize()
    cutoff = 1

    def step():
        for (b, n), dp in zip(subs, d_pass):
            _capi.check(L.gt_median_count_at_least_dev(storage.handle, _capi.SHIFTER_CAN, K, b.data_ptr(), offs.data_ptr(), n,
                                                       n * read_len, cutoff, dp.data_ptr(), d_total.data_ptr()),
                        "gt_median_count_at_least_dev")

    for _ in range(args.warmup):
        d_total.zero_()
        torch.cuda.synchronize()
        step()
        L.gt_synchronize()
        assert int(d_total.item()) == kmers_per_step
    _capi.check(L.gt_profile_enable(1), "gt_profile_enable")
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = L.gt_launch_count()
    _capi.check(L.gt_timer_record(0), "gt_timer_record")
    for _ in range(args.steps):
        step()
    _capi.check(L.gt_timer_record(1), "gt_timer_record")
    ms = L.gt_timer_elapsed_ms(0, 1)
    L.gt_synchronize()
    launches = int(L.gt_launch_count() - launches0)
    clocks = sampler.stop()
    prof_ms, prof_n = np.zeros(5, dtype=np.float64), np.zeros(5, dtype=np.uint64)
    _capi.check(L.gt_profile_get_detail(prof_ms.ctypes.data, prof_n.ctypes.data, 5), "gt_profile_get_detail")
    L.gt_profile_enable(0)
    value = kmers_per_step * args.steps / (ms / 1e3)
    check["all_reads_pass_cutoff_1"] = bool(all(int(dp.sum().item()) == n for dp, (_, n) in zip(d_pass, subs)))

    e2e = None
    if not args.no_e2e:
        host_b = torch.empty(total_reads * read_len, dtype=torch.uint8, pin_memory=True)
        p = 0
        for b, n in subs:
            host_b[p:p + n * read_len].copy_(b[:n * read_len])
            p += n * read_len
        host_o = torch.empty(total_reads + 1, dtype=torch.int64, pin_memory=True)
        host_o.copy_(torch.arange(total_reads + 1, dtype=torch.int64) * read_len)
        host_p = torch.empty(total_reads, dtype=torch.uint8, pin_memory=True)
        torch.cuda.synchronize()
        hb, ho, hp = host_b.numpy(), host_o.numpy().view(np.uint64), host_p.numpy()

        def step_host():
            nk = L.gt_median_count_at_least(storage.handle, _capi.SHIFTER_CAN, K, hb.ctypes.data, ho.ctypes.data, total_reads, cutoff,
                                            hp.ctypes.data, None)
            return _capi.check(nk, "gt_median_count_at_least")

        assert step_host() == kmers_per_step
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_host()
        L.gt_synchronize()
        dt = time.perf_counter() - t0
        e2e = {"value": kmers_per_step * args.steps / dt, "unit": "k-mers/s", "h2d_bytes_per_step": int(hb.nbytes + ho.nbytes),
               "d2h_bytes_per_step": int(hp.nbytes) + 8, "ms_per_step": dt * 1e3 / args.steps,
               "api": "gt_median_count_at_least(host ASCII, host offsets) -> pass[read] in pinned host memory"}
        check["e2e_all_reads_pass_cutoff_1"] = bool(hp.all())

    peak, peak_src = measured_peak()
    table_bytes = sum(storage.table_bytes(i) for i in range(n_tables))
    r_load = measured_r_rand(L, 1, table_bytes)
    algo = 32 * n_tables
    achieved = kmers_per_step * args.steps * algo / (ms / 1e3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "peak_source": peak_src, "kernel": "k_walk<OP_MEDIAN> (direct random 32 B sector loads, 4 per k-mer)",
                "algorithmic_bytes_per_kmer": algo,
                "kernels": {"k_walk": {"launches": int(prof_n[2]), "ms_total": float(prof_ms[2]),
                                       "ms_per_launch": float(prof_ms[2] / max(1, int(prof_n[2]))),
                                       "share_of_step": float(prof_ms[2] / ms) if ms > 0 else 0.0}},
                "random_sector_roofline": None if not r_load else {
                    "r_rand_loads_per_s": r_load, "measured": "in this run, random 32 B loads over %.1f GB" % (table_bytes / 1e9),
                    "achieved_loads_per_s": kmers_per_step * args.steps * n_tables / (ms / 1e3),
                    "frac": kmers_per_step * args.steps * n_tables / (ms / 1e3) / r_load}}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f).get("c2q")
        if t:
            roofline["traffic"] = float(t["k_walk"])
            roofline["traffic_detail"] = t
            # what the random lookups really draw from HBM (the access granule is larger than the 32 B sector)
            roofline["dram_bytes_per_kmer_measured"] = float(t["k_walk"]) / float(t["kmers_per_launch"])
            roofline["dram_frac"] = roofline["dram_bytes_per_kmer_measured"] * value / 1e9 / peak
    except Exception:
        pass

    cpu = None
    if not args.no_cpu_baseline:
        from goetia_b200.synth import synth_reads
        from oracle import binding
        del storage, graph
        impl = binding.Ref(1, 1, K, sizes) if binding.have_ref() else binding.Port(1, 1, K, sizes)
        kname = "reference" if binding.have_ref() else "port"
        nb = 200_000
        cb, co = synth_reads(seed, 0, nb, read_len)
        impl.insert_reads(cb, co)
        t0, nq, budget = time.perf_counter(), 0, 15.0
        for r_ in range(nb):
            impl.median_count_at_least(cb[r_ * read_len:(r_ + 1) * read_len].tobytes(), cutoff)
            nq += 1
            if (nq & 1023) == 0 and time.perf_counter() - t0 > budget:
                break
        secs = time.perf_counter() - t0
        impl.close()
        cpu = {"value": nq * kpr / secs, "unit": "k-mers/s", "cores": 1, "kind": kname,
               "sample": "DiginormFilter::median_count_at_least per read over %d reads (%d k-mers) of the same set on full-size "
                         "tables holding the first %d reads, %.1f s, one thread" % (nq, nq * kpr, nb, secs)}

    out = {"metric": "k-mers queried/sec (device-timed)", "value": value, "unit": "k-mers/s", "n_gpus": 1, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "u64", "data": "synthetic",
           "config": {"workload": desc if not args.reads else desc + " [reads overridden: %d]" % total_reads, "K": K,
                      "tablesizes": sizes, "reads": total_reads, "read_len": read_len, "kmers_per_step": kmers_per_step,
                      "cutoff": cutoff, "seed": seed, "generator": "counter-based splitmix64 (gt_synth_bases_dev / goetia_b200/synth.py)",
                      "l2": "inputs larger than L2 (%.1f GB reads, %.1f GB tables)" % (total_reads * read_len / 1e9, table_bytes / 1e9)},
           "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "check": check}
    print(json.dumps(out))
    return 0


def nvlink_report(nv0, nv1, kmers_rank_total, n_tables, world, ms, peer_bytes=0, shipped_bytes=0):
    """NVLink traffic of rank 0's GPU over the timed region (nvidia-smi counters) beside the algorithmic figure of
    SURVEY.md section 8d: n_tables x 4 B x (G-1)/G egress per k-mer hashed on this rank."""
    algo = kmers_rank_total * n_tables * 4 * (world - 1) / world
    out = {"algorithmic_egress_bytes_rank0": algo, "algorithmic_egress_bytes_per_kmer": n_tables * 4 * (world - 1) / world,
           "algorithmic_egress_GBps": algo / (ms / 1e3) / 1e9 if ms > 0 else None, "link_peak_GBps_per_direction": 900.0}
    if nv0 and nv1:
        tx, rx = nv1["tx_bytes"] - nv0["tx_bytes"], nv1["rx_bytes"] - nv0["rx_bytes"]
        out.update({"measured_tx_bytes_rank0": tx, "measured_rx_bytes_rank0": rx, "links": nv1.get("links"),
                    "measured_tx_GBps": tx / (ms / 1e3) / 1e9 if ms > 0 else None,
                    "measured_tx_bytes_per_kmer": tx / kmers_rank_total if kmers_rank_total else None,
                    "source": "nvidia-smi nvlink -gt d on rank 0's GPU, before / after the timed region"})
    else:
        out["hardware_counters"] = "nvidia-smi nvlink -gt d reports N/A on this box"
    if peer_bytes:
        out.update({"peer_store_bytes_rank0": peer_bytes, "peer_store_bytes_per_kmer": peer_bytes / kmers_rank_total if kmers_rank_total else None,
                    "peer_store_GBps": peer_bytes / (ms / 1e3) / 1e9 if ms > 0 else None,
                    "peer_store_frac_of_link_peak": peer_bytes / (ms / 1e3) / 1e9 / 900.0 if ms > 0 else None,
                    "peer_store_source": "software counter: cursors of the buckets rank 0's k_bucket filled in peers' HBM "
                                         "(entries x 4 B, 16-byte run padding included), summed over the timed region"})
    if shipped_bytes:
        out.update({"copy_engine_bytes_rank0": shipped_bytes, "copy_engine_bytes_per_kmer": shipped_bytes / kmers_rank_total if kmers_rank_total else None,
                    "copy_engine_GBps": shipped_bytes / (ms / 1e3) / 1e9 if ms > 0 else None,
                    "copy_engine_frac_of_link_peak": shipped_bytes / (ms / 1e3) / 1e9 / 900.0 if ms > 0 else None,
                    "copy_engine_source": "ce transport: bytes rank 0 handed to the copy engines over the timed region (whole bucket "
                                          "regions at their capacity + the overflow lists; peer_store_bytes is the part that was filled)"})
    return out


def multi_gpu_arm(args, rank, world, local_rank, torch, gb, _capi, sizes, total_reads):
    """N > 1: reads sharded over the ranks, every table partitioned by slot range, one bucket
    exchange (NCCL all-to-all) per round.  Strong scaling: the workload's read set is fixed."""
    import torch.distributed as dist
    from goetia_b200.shard import ShardedStorage
    kind, K, x, n_tables, _, read_len, seed, desc = WORKLOADS[args.workload]
    L = _capi.lib()
    dev = torch.device("cuda", local_rank)
    kpr = read_len - K + 1
    reads_rank = total_reads // world + (1 if rank < total_reads % world else 0)
    max_rank_reads = total_reads // world + (1 if total_reads % world else 0)
    # rounds are smaller than the single-GPU sub-batches: k_apply of round i overlaps k_bucket of round i+1,
    # so only the last round's apply is exposed
    # rounds per step (measured, ce transport): 3 at N = 2 (75.1 vs 73.9 G k-mers/s with 5), 4 at N = 8 (225.9 vs 211.4
    # with 2: the copy engines stay busier when the rounds are finer); GT_BENCH_ROUNDS / GT_BENCH_ROUND_BASES override
    if "GT_BENCH_ROUND_BASES" in os.environ:
        rounds = max(2, -(-max_rank_reads * read_len // int(os.environ["GT_BENCH_ROUND_BASES"])))
    else:
        rounds = max(2, int(os.environ.get("GT_BENCH_ROUNDS", 3 if world <= 4 else 4)))
    rounds = max(2, min(rounds, max(2, max_rank_reads)))
    while -(-max_rank_reads // rounds) * kpr > (1 << 31):  # a round's k-mers must fit the 32-bit bucket cursors' budget
        rounds += 1
    per_round = -(-max_rank_reads // rounds)
    # equal-length reads: the k-mer count of every round is known, so the round budget (which sizes the exchange buffers
    # and, with the ce transport, the bytes the copy engines ship) is given in k-mers and every batch carries its count
    st = ShardedStorage(kind, sizes, per_round * kpr)
    # rank r holds a contiguous range of the SAME global read set the single-GPU run (and the CPU reference) sees
    first_read = rank * (total_reads // world) + min(rank, total_reads % world)
    gold = golden_entry(args.workload, total_reads, WORKLOADS[args.workload][4])
    reset_each_step = kind != 0
    subs, r = [], 0
    for i in range(rounds):  # every rank runs the same number of rounds (the exchange is collective)
        n = max(0, min(per_round, reads_rank - r))
        subs.append((synth_reads_device(torch, L, first_read + r, n, read_len, seed, dev), n))
        r += n
    offs = torch.arange(per_round + 1, dtype=torch.int64, device=dev) * read_len
    torch.cuda.synchronize()

    d_total = torch.zeros(1, dtype=torch.int64, device=dev)  # k-mers consumed, accumulated on the device
    torch.cuda.synchronize()

    def step():
        # no host wait anywhere in a step: rounds are queued back to back on the two streams
        if reset_each_step:
            st.reset()
        for b, n in subs:
            if n:
                st.bucket_sequences_dev_async(_capi.SHIFTER_CAN, K, b.data_ptr(), offs.data_ptr(), n, n * read_len,
                                              d_total.data_ptr(), n_kmers_upper=n * kpr)
            st.exchange_and_apply()

    kmers_rank = reads_rank * kpr
    for _ in range(args.warmup):
        d_total.zero_()
        torch.cuda.synchronize()
        step()
        st.synchronize()
        assert int(d_total.item()) == kmers_rank
    st.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()  # before the barrier: its start-up must not skew the ranks
    dist.barrier()
    torch.cuda.synchronize()
    _capi.check(L.gt_profile_enable(1), "gt_profile_enable")
    st.peer_store_bytes(reset=True)
    st.shipped_bytes = 0
    if getattr(st, "_copy_prof", None) is not None:
        st._copy_prof = []
    nv0 = nvlink_counters(local_rank) if rank == 0 else None  # before the barrier: not inside the timed region
    dist.barrier()
    torch.cuda.synchronize()
    sampler.begin()
    launches0 = L.gt_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(st.stream)
    for _ in range(args.steps):
        step()
    st.join()  # the compute stream waits for the apply stream: ev1 covers both
    ev1.record(st.stream)
    st.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    nv1 = nvlink_counters(local_rank) if rank == 0 else None
    peer_bytes = st.peer_store_bytes()
    shipped_bytes = st.shipped_bytes
    copy_prof = st.copy_profile() if hasattr(st, "copy_profile") else None
    dist.barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    launches = int(L.gt_launch_count() - launches0)
    clocks = sampler.stop()
    prof_ms, prof_n = np.zeros(5, dtype=np.float64), np.zeros(5, dtype=np.uint64)
    _capi.check(L.gt_profile_get_detail(prof_ms.ctypes.data, prof_n.ctypes.data, 5), "gt_profile_get_detail")
    L.gt_profile_enable(0)
    kmers_per_step = total_reads * kpr
    value = kmers_per_step * args.steps / (ms / 1e3)
    info = st.pending_info()  # raises if any update was dropped
    n_occ = st.n_occupied()
    check = checksum_check(gold, [st.checksum(i) for i in range(n_tables)], n_occ)  # collective: the shards' sums add up
    check["n_occupied_all_ranks"] = n_occ

    # e2e: pinned 2-bit packed host reads (the parser stage's product) -> H2D inside the timed region -> hash + bucket ->
    # exchange -> apply; wall clock, max over ranks
    e2e = None
    if not args.no_e2e:
        from goetia_b200.batch import pack_reads_host
        wpr = (per_round * read_len + 31) // 32 + 8  # words per round (+ slack the kernels may read)
        hosts = []
        for b, n in subs:
            hw = torch.zeros(wpr, dtype=torch.int64, pin_memory=True)
            hf = torch.zeros(max(per_round, 1), dtype=torch.uint8, pin_memory=True)
            if n:
                ascii_np = b[:n * read_len].cpu().numpy()
                pack_reads_host(ascii_np, np.arange(n + 1, dtype=np.uint64) * np.uint64(read_len), 0,
                                hw.numpy().view(np.uint64), hf.numpy())
            hosts.append((hw, hf))
        host_offs = torch.empty(per_round + 1, dtype=torch.int64, pin_memory=True)
        host_offs.copy_(offs)
        dwords = [torch.zeros(wpr, dtype=torch.int64, device=dev) for _ in range(2)]
        dflags = [torch.zeros(max(per_round, 1), dtype=torch.uint8, device=dev) for _ in range(2)]
        doff = [torch.empty(per_round + 1, dtype=torch.int64, device=dev) for _ in range(2)]
        copy_stream = torch.cuda.Stream()
        torch.cuda.synchronize()

        pieces = max(1, int(os.environ.get("GT_BENCH_E2E_PIECES", "4")))

        def step_host():
            # Every round travels in `pieces` H2D copies (copy stream); hash + bucket of a piece (compute stream) starts as
            # soon as it has landed, so only the first piece of a step is exposed.  The exchange and the apply (apply
            # stream) stay per round.  One result read-back (the k-mer count) ends the step.
            consumed = [None, None]
            if reset_each_step:
                st.reset()
            with torch.cuda.stream(st.stream):
                d_total.zero_()
            for i, ((hw, hf), (b, n)) in enumerate(zip(hosts, subs)):
                dw, df, do = dwords[i & 1], dflags[i & 1], doff[i & 1]
                with torch.cuda.stream(copy_stream):
                    if consumed[i & 1] is not None:
                        copy_stream.wait_event(consumed[i & 1])
                    do.copy_(host_offs, non_blocking=True)
                    df.copy_(hf, non_blocking=True)
                # piece boundaries on multiples of 64 reads: a piece then starts on a 16-byte boundary of the packed stream whatever the read length
                per_piece = (-(-max(n, 1) // pieces) + 63) // 64 * 64
                for r0 in range(0, max(n, 1), per_piece):
                    r1 = min(max(n, 1), r0 + per_piece)
                    w0, w1 = r0 * read_len // 32, (r1 * read_len + 31) // 32
                    with torch.cuda.stream(copy_stream):
                        dw[w0:w1 + 1].copy_(hw[w0:w1 + 1], non_blocking=True)
                        ready = torch.cuda.Event()
                        ready.record(copy_stream)
                    st.stream.wait_event(ready)
                    if n:
                        # equal-length reads: the first r1-r0+1 offsets describe any piece
                        st.bucket_packed_dev_async(_capi.SHIFTER_CAN, K, dw.data_ptr() + w0 * 8, w1 - w0 + 1, do.data_ptr(),
                                                   df.data_ptr() + r0, r1 - r0, (r1 - r0) * read_len, d_total.data_ptr(),
                                                   n_kmers_upper=(r1 - r0) * kpr)
                free = torch.cuda.Event()
                free.record(st.stream)
                consumed[i & 1] = free
                st.exchange_and_apply()
            st.synchronize()
            return int(d_total.item())  # the step's result read back from the device

        for _ in range(min(args.warmup, 2)):
            assert step_host() == kmers_rank
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_host()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dts = float(dt.item())
        n_occ2 = st.n_occupied()
        e2e_ok = None if gold is None else ([st.checksum(i) for i in range(n_tables)] == [int(x) for x in gold["checksums"]]
                                            and n_occ2 == int(gold["n_occupied"]))
        check["e2e_tables_checksum_equal_reference"] = e2e_ok
        e2e = {"value": kmers_per_step * args.steps / dts, "unit": "k-mers/s",
               "h2d_bytes_per_step": int(sum(hw.numel() * 8 + hf.numel() for hw, hf in hosts) + rounds * host_offs.numel() * 8) * world,
               "d2h_bytes_per_step": 8 * world, "ms_per_step": dts * 1e3 / args.steps,
               "api": "ShardedStorage: pinned 2-bit packed host reads (gt_pack_reads_host, outside the timed region) -> H2D in %d "
                      "pieces per round -> gt_insert_packed_dev_async (hash + bucket, transport %s) -> exchange -> apply (per rank)"
                      % (pieces, st.transport)}

    if rank == 0:
        peak, peak_src = measured_peak()
        ins_ms = float(prof_ms[0] + prof_ms[1])
        algo_bytes = ALGO_BYTES_PER_KMER_PER_TABLE * n_tables
        achieved = kmers_rank * args.steps * algo_bytes / (ins_ms / 1e3) / 1e9 if ins_ms > 0 else 0.0
        out = {
            "metric": "k-mers inserted/sec (device-timed)", "value": value, "unit": "k-mers/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": desc if not args.reads else desc + " [reads overridden: %d]" % total_reads,
                       "K": K, "tablesizes": sizes, "reads": total_reads, "read_len": read_len,
                       "kmers_per_step": kmers_per_step, "mode": "GT_MODE_BLIND (write-combined)",
                       "parallelism": "reads sharded over %d ranks; tables partitioned by slot range; " % world + (
                           "k_bucket stores foreign buckets straight into the owner's HBM over NVLink (CUDA IPC peer "
                           "memory), one small NCCL all-to-all of fill counts per round; k_apply of round i overlaps "
                           "k_bucket of round i+1" if st.transport == "p2p" else
                           "k_bucket writes foreign buckets to local staging areas, the copy engines ship them into the "
                           "owners' HBM over NVLink (CUDA IPC peer memory) while the SMs apply round i-1 and hash round i+1; "
                           "one small NCCL all-to-all of fill counts per round" if st.transport == "ce" else
                           "one NCCL all-to-all of bucket regions per round; k_apply of round i overlaps k_bucket of "
                           "round i+1"),
                       "transport": st.transport,
                       "rounds_per_step": rounds, "slice_shift": info["slice_shift"], "n_buckets": info["n_buckets"],
                       "bucket_overflow_updates": info["n_direct"], "seed": seed,
                       "generator": "counter-based splitmix64 (gt_synth_bases_dev / goetia_b200/synth.py); rank r holds a contiguous range of the one global read set",
                       "l2": "inputs larger than L2 (per rank: %.1f GB reads, %.1f GB exchange buffers)"
                             % (reads_rank * read_len / 1e9, info["entries"] * 4 / 1e9),
                       "tables": "reset at the start of every step" if reset_each_step else "accumulate across steps (Bloom tables are idempotent)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src,
                         "kernel": "k_bucket + k_apply on rank 0 (its share of the reads / of the slices)",
                         "algorithmic_bytes_per_kmer": algo_bytes,
                         "kernels": {name: {"launches": int(prof_n[i]), "ms_total": float(prof_ms[i]),
                                            "share_of_step": float(prof_ms[i] / ms) if ms > 0 else 0.0}
                                     for i, name in enumerate(("k_bucket", "k_apply", "k_walk", "k_rebucket", "k_apply_win")) if prof_n[i]}},
            "nvlink": dict(nvlink_report(nv0, nv1, kmers_rank * args.steps, n_tables, world, ms, peer_bytes, shipped_bytes),
                           **({"copy_engine_profile_rank0": copy_prof} if copy_prof else {})),
            "cpu_baseline": None,
            "check": check,
        }
        print(json.dumps(out))
    st.close()
    dist.destroy_process_group()
    return 0


def sketch_arm(args, rank, world, local_rank, torch, gb, _capi):
    """`--workload c4`: every rank sketches its shard of the reads (k_sketch), then the hash sets are
    united (all-gather + add_hashes) -- exact because a sketch is a set.  Reported in k-mers sketched/s."""
    _, K, scaled, _, total_reads, read_len, seed, desc = WORKLOADS["c4"]
    if args.reads:
        total_reads = args.reads
    L = _capi.lib()
    dev = torch.device("cuda", local_rank)
    kpr = read_len - K + 1
    reads_rank = total_reads // world + (1 if rank < total_reads % world else 0)
    reads_per_sub = max(1, min(reads_rank, SUB_BATCH_BASES // read_len))
    first_read = rank * (total_reads // world) + min(rank, total_reads % world)
    subs, r = [], 0
    while r < reads_rank:
        n = min(reads_per_sub, reads_rank - r)
        subs.append((synth_reads_device(torch, L, first_read + r, n, read_len, seed, dev), n))
        r += n
    offs = torch.arange(reads_per_sub + 1, dtype=torch.int64, device=dev) * read_len
    torch.cuda.synchronize()
    sk = gb.SourmashSketch.Sketch(0, K, False, False, False, 42, scaled)

    def step():
        nk = 0
        for b, n in subs:
            nk += sk.insert_sequences_dev(b.data_ptr(), offs.data_ptr(), n, n * read_len)
        if world > 1:
            sk.allgather_merge()
        return nk

    if world > 1:
        import torch.distributed as dist
    for _ in range(args.warmup):
        assert step() == reads_rank * kpr
    L.gt_synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = L.gt_launch_count()
    t0 = time.perf_counter()
    _capi.check(L.gt_timer_record(0), "gt_timer_record")
    for _ in range(args.steps):
        step()
    _capi.check(L.gt_timer_record(1), "gt_timer_record")
    ms_dev = L.gt_timer_elapsed_ms(0, 1)
    L.gt_synchronize()
    ms_wall = (time.perf_counter() - t0) * 1e3
    ms = max(ms_dev, ms_wall) if world > 1 else ms_dev  # the union's collectives are not on the library stream
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    launches = int(L.gt_launch_count() - launches0)
    clocks = sampler.stop()
    n_mins = sk.size()
    kmers_per_step = total_reads * kpr
    value = kmers_per_step * args.steps / (ms / 1e3)
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle.binding import PortSketch
        from goetia_b200.synth import synth_reads
        nb = min(reads_rank, 150_000)
        cb, co = synth_reads(seed, 0, nb, read_len)
        o = PortSketch(0, K, 42, scaled=scaled)
        nk, secs = o.add_reads(cb, co)
        cpu = {"value": nk / secs, "unit": "k-mers/s", "cores": 1, "kind": "port",
               "sample": "%d reads (%d k-mers) of the same set through the plain-C restatement of "
                         "KmerMinHash::add_sequence (libsourmash is absent: parity unpinned), %.1f s" % (nb, nk, secs)}
        sub = gb.SourmashSketch.Sketch(0, K, False, False, False, 42, scaled)
        sub.insert_sequences(cb, co)
        check = {"prefix_hash_set_equals_oracle": bool(np.array_equal(sub.mins(), o.mins())), "n_mins": n_mins}
    else:
        check = {"n_mins": n_mins}
    if rank == 0:
        peak, peak_src = measured_peak()
        in_bytes = total_reads * read_len * args.steps  # ASCII read once by k_pack; packed words re-read by k_sketch
        out = {"metric": "k-mers sketched/sec (device-timed)", "value": value, "unit": "k-mers/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
               "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
               "config": {"workload": desc if not args.reads else desc + " [reads overridden: %d]" % total_reads, "K": K,
                          "scaled": scaled, "reads": total_reads, "read_len": read_len, "kmers_per_step": kmers_per_step,
                          "seed": seed, "l2": "inputs larger than L2 (%.1f GB of reads per rank)" % (reads_rank * read_len / 1e9),
                          "parallelism": "reads sharded over %d rank(s); hash sets united by all-gather" % world},
               "clocks": clocks, "e2e": None, "gpu_launches": launches,
               "roofline": {"bound": "hbm", "achieved": in_bytes * 1.25 / (ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                            "frac": in_bytes * 1.25 / (ms / 1e3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                            "kernel": "k_sketch",
                            "note": "integer-ALU bound (MurmurHash3_x64_128 per window), not HBM bound: 1.25 B/base "
                                    "algorithmic (ASCII in + 2-bit words out + words in) is far below the HBM roof"},
               "cpu_baseline": cpu, "check": check}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


def storage_arm(args):
    """`--workload storage`: the reference's own storage micro-benchmark shape (src/goetia/benchmarks/
    bench_storage.cc:17-61: n uniform 64-bit hashes, Storage(n/4, 4)) through the hash-vector entry points
    gt_insert_hashes / gt_query_hashes with HOST buffers -- the only shape BASELINE.md holds published numbers
    for (notebooks/Benchmarks.ipynb:336-342, hardware unknown).  Secondary benchmark, one GPU."""
    import goetia_b200 as gb
    from goetia_b200 import _capi
    gb.init(int(os.environ.get("LOCAL_RANK", "0")))
    L = _capi.lib()
    n = args.reads or 100_000_000
    published = {0: ("BitStorage", 1e8 / 13.42, 1e8 / 8.09), 1: ("ByteStorage", 1e8 / 17.75, 1e8 / 12.06),
                 2: ("NibbleStorage", 1e8 / 35.79, 1e8 / 12.05)}
    rng = np.random.default_rng(5)
    hashes = rng.integers(0, 2**63, n, dtype=np.int64).view(np.uint64) * np.uint64(2) + rng.integers(0, 2, n).astype(np.uint64)
    counts = np.zeros(n, dtype=np.int16)
    rows = {}
    for kind, (name, pub_ins, pub_q) in published.items():
        st = [gb.BitStorage, gb.ByteStorage, gb.NibbleStorage][kind](gb.get_n_primes_near_x(4, n // 4))
        t_ins, t_q = [], []
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            _capi.check(L.gt_insert_hashes(st.handle, hashes.ctypes.data, n, gb.MODE_BLIND, None), "gt_insert_hashes")
            L.gt_synchronize()
            t1 = time.perf_counter()
            _capi.check(L.gt_query_hashes(st.handle, hashes.ctypes.data, n, counts.ctypes.data), "gt_query_hashes")
            t2 = time.perf_counter()
            if i >= args.warmup:
                t_ins.append(t1 - t0)
                t_q.append(t2 - t1)
        assert int(counts.min()) >= 1
        rows[name] = {"insert_hashes_per_s": n / float(np.median(t_ins)), "query_hashes_per_s": n / float(np.median(t_q)),
                      "published_insert_hashes_per_s": pub_ins, "published_query_hashes_per_s": pub_q,
                      "insert_vs_published": n / float(np.median(t_ins)) / pub_ins,
                      "query_vs_published": n / float(np.median(t_q)) / pub_q}
        st.close()
    v = rows["BitStorage"]["insert_hashes_per_s"]
    print(json.dumps({"metric": "hashes inserted/sec (host buffers, wall clock)", "value": v, "unit": "hashes/s", "n_gpus": 1,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": n / v * 1e3, "higher_is_better": True,
                      "scaling": "strong", "vs_baseline": rows["BitStorage"]["insert_vs_published"], "dtype": "u64",
                      "data": "synthetic", "config": {"workload": "bench_storage.cc shape: %d uniform u64 hashes, Storage(n/4, 4), "
                                                                  "gt_insert_hashes / gt_query_hashes from pageable host memory" % n,
                                                      "published": "notebooks/Benchmarks.ipynb:336-342 (hardware unknown, 1 thread)"},
                      "storages": rows}))
    return 0


def reference_arm(args, rank, world, kind, K, x, n_tables, total_reads, read_len, seed, desc):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    if rank != 0:
        return 0
    if kind < 0:
        print(json.dumps({"impl": "reference", "unavailable": "the sketch arithmetic lives in libsourmash (absent); "
                                                             "see cpu_baseline of --workload c4"}))
        return 0
    from oracle import binding
    threads = os.cpu_count() or 1
    sizes = binding.Port.primes_near(n_tables, x)
    impl, kname, modes = cpu_reference_modes(kind, K, sizes, threads)
    from goetia_b200.synth import synth_reads
    kpr = read_len - K + 1
    cursor = [0]

    def make(n):  # the next n reads of the workload's own read stream: every call inserts reads never seen before
        b, o = synth_reads(seed, cursor[0], n, read_len)
        cursor[0] += n
        return b, o

    def run(b, o, t):
        return cpu_time_reads(impl, kname, b, o, 0, o.size - 1, t)

    # calibrate both modes on fresh reads (this also faults the tables in), keep the faster one
    rates = {}
    for t in modes:
        b, o = make(100_000)
        nk, secs = run(b, o, t)
        rates[t] = nk / max(secs, 1e-9)
    use_threads = max(rates, key=rates.get)
    n_steps = args.steps + args.warmup
    per_step_s = min(12.0, 120.0 / max(1, n_steps))
    n = int(max(20_000, min(total_reads, rates[use_threads] * per_step_s / kpr)))
    tot_k, tot_s = 0, 0.0
    for i in range(n_steps):
        b, o = make(n)  # every step inserts reads never seen before
        nk, secs = run(b, o, use_threads)
        if i >= args.warmup:
            tot_k += nk
            tot_s += secs
    value = tot_k / tot_s
    sample = ("%d fresh synthetic reads (%d k-mers) per step of the workload's shape into the full-size tables; "
              "calibration: %s" % (n, n * kpr, ", ".join("%d thread(s) %.3g k-mers/s" % (t, r) for t, r in rates.items())))
    out = {"impl": "reference", "metric": "k-mers inserted/sec (device-timed)", "value": value, "unit": "k-mers/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_s * 1e3 / args.steps,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
           "config": {"workload": desc, "K": K, "tablesizes": sizes, "reads_per_step": n, "read_len": read_len,
                      "note": "CPU implementation on the host cores; each step is a bounded sample of the workload"},
           "cpu_baseline": {"value": value, "unit": "k-mers/s", "cores": use_threads, "kind": kname, "sample": sample},
           "e2e": {"value": value, "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
