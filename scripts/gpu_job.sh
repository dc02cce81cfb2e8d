set -x
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 python -m pytest tests/test_shard.py -m gpu -x -q -k "ce or overflow" > gpurun_out/k_pytest_n2_ce.log 2>&1; echo "rc=$?"
tail -5 gpurun_out/k_pytest_n2_ce.log
GT_SHARD_TRANSPORT=ce timeout 400 $R --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/k_n2_ce.json 2> gpurun_out/k_n2_ce.err; echo "rc=$?"
tail -c 600 gpurun_out/k_n2_ce.err
GT_SHARD_TRANSPORT=ce GT_BENCH_ROUND_BASES=1500000000 timeout 400 $R --master-port 29542 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e > gpurun_out/k_n2_ce_r3.json 2> gpurun_out/k_n2_ce_r3.err; echo "rc=$?"
GT_SHARD_TRANSPORT=p2p timeout 400 $R --master-port 29543 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e > gpurun_out/k_n2_p2p.json 2> gpurun_out/k_n2_p2p.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("k_n2_ce","k_n2_ce_r3","k_n2_p2p"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(d["value"]/1e9,2), d["ms_per_step"], d["config"]["rounds_per_step"], (d.get("e2e") or {}).get("value"), d["check"].get("tables_checksum_equal_reference"), d["check"].get("e2e_tables_checksum_equal_reference"))
        print("   ", {k:(v["ms_total"],v["launches"]) for k,v in d["roofline"]["kernels"].items()})
        print("   ", {k:v for k,v in d["nvlink"].items() if "GBps" in k})
    except Exception as e: print(f,"ERR",e)
PY
