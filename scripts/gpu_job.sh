set -x
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 python -m pytest tests/test_shard.py -m gpu -x -q -k "ce or p2p" > gpurun_out/w_pytest_n2.log 2>&1; echo "rc=$?"
tail -6 gpurun_out/w_pytest_n2.log
GT_SHARD_PROFILE_COPIES=1 timeout 300 $R --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/w_n2.json 2> gpurun_out/w_n2.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("w_n2",):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(d["value"]/1e9,2), d["ms_per_step"], d["config"]["rounds_per_step"], (d.get("e2e") or {}).get("value"), d["check"].get("tables_checksum_equal_reference"), d["check"].get("e2e_tables_checksum_equal_reference"), d["nvlink"].get("copy_engine_profile_rank0"))
        print("   ", {k:(round(v["ms_total"],1),v["launches"]) for k,v in d["roofline"]["kernels"].items()})
        print("   ", {k:v for k,v in d["nvlink"].items() if "GBps" in k or "per_kmer" in k})
    except Exception as e: print(f,"ERR",e)
PY
grep -v "^\*\|OMP\|^$" gpurun_out/w_n2.err | tail -5
