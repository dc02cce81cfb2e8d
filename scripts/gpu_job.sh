set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/p_pytest_all.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/p_pytest_all.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/p_c3.json 2> gpurun_out/p_c3.err; echo "rc=$?"
timeout 300 python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/p_c2.json 2> gpurun_out/p_c2.err; echo "rc=$?"
timeout 300 python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/p_c5.json 2> gpurun_out/p_c5.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("p_c3","p_c2","p_c5"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(d["value"]/1e9,2), round(d["ms_per_step"],2), (d.get("e2e") or {}).get("value"), d["check"].get("tables_checksum_equal_reference"), d["roofline"].get("dram_frac"))
        print("   ", {k:(v["launches"],round(v["ms_per_launch"],2)) for k,v in d["roofline"]["kernels"].items()})
    except Exception as e: print(f,"ERR",e)
PY
tail -3 gpurun_out/p_c3.err
