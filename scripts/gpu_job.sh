set -x
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/y_pytest_all.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/y_pytest_all.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/y_smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/y_smoke.log
timeout 300 python bench.py > gpurun_out/y_bench_default.json 2> gpurun_out/y_bench_default.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/y_bench_default.json").read().strip().splitlines()[-1])
print(round(d["value"]/1e9,2), round(d["ms_per_step"],2), d["steps"], d["warmup"], (d.get("e2e") or {}).get("value"), d["check"], d["cpu_baseline"]["value"], d["roofline"]["frac"], d["roofline"].get("dram_frac"), d["gpu_launches"])
PY
