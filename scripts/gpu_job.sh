set -x
timeout 900 python -m pytest tests/test_gpu_bucket.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/j_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/j_pytest.log
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 $B --workload c2 > gpurun_out/j_c2.json 2> gpurun_out/j_err3.err; echo "rc=$?"
timeout 300 $B --workload c5 > gpurun_out/j_c5.json 2> gpurun_out/j_err4.err; echo "rc=$?"
GT_BUCKET_OVERLAP=0 timeout 300 $B --workload c2 > gpurun_out/j_c2_noovl.json 2> gpurun_out/j_err5.err; echo "rc=$?"
