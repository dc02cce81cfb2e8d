#!/usr/bin/env python
"""NVLink ceiling of this box for the sharded storage's exchange pattern: every GPU copies one buffer to every
other GPU at the same time (copy engines, peer access), ONE process driving all GPUs.  Prints one JSON line per
configuration: GB/s leaving each GPU.  (python scripts/nvlink_probe.py [MB per copy] [streams per GPU ...])"""
import json
import sys
import time

import torch

n = torch.cuda.device_count()
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 512
stream_counts = [int(x) for x in sys.argv[2:]] or [1, 2, 4, 7]
src = [torch.empty(mb << 20, dtype=torch.uint8, device="cuda:%d" % g).fill_(g) for g in range(n)]
dst = [[torch.empty(mb << 20, dtype=torch.uint8, device="cuda:%d" % q) if q != g else None for q in range(n)] for g in range(n)]
for g in range(n):
    for q in range(n):
        if q != g and not torch.cuda.can_device_access_peer(g, q):
            print(json.dumps({"error": "no peer access %d -> %d" % (g, q)}))
            sys.exit(0)


def run(n_streams, senders, reps=4):
    streams = {g: [torch.cuda.Stream(device=g) for _ in range(n_streams)] for g in senders}

    def once():
        for g in senders:
            i = 0
            for k in range(1, n):
                q = (g + k) % n
                with torch.cuda.device(g), torch.cuda.stream(streams[g][i % n_streams]):
                    dst[g][q].copy_(src[g], non_blocking=True)
                i += 1
    once()
    for g in range(n):
        torch.cuda.synchronize(g)
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    for g in range(n):
        torch.cuda.synchronize(g)
    dt = time.perf_counter() - t0
    return reps * (n - 1) * (mb << 20) / dt / 1e9


for ns in stream_counts:
    one = run(ns, [0])
    allg = run(ns, list(range(n)))
    print(json.dumps({"gpus": n, "MB_per_copy": mb, "streams_per_gpu": ns, "one_sender_GBps": round(one, 1),
                      "all_senders_GBps_per_gpu": round(allg, 1)}))
