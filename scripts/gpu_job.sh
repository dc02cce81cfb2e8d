set -x
timeout 900 python -m pytest tests/test_gpu_bucket.py -x -q -m gpu > gpurun_out/j_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/j_pytest.log
B="python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
GT_BUCKET_OVERLAP=0 timeout 300 $B > gpurun_out/j_c3_4cta_noovl.json 2> gpurun_out/j_err1.err; echo "rc=$?"
GT_BUCKET_OVERLAP=0 GT_BUCKET_CTAS=3 timeout 300 $B > gpurun_out/j_c3_3cta_noovl.json 2> gpurun_out/j_err2.err; echo "rc=$?"
timeout 300 $B > gpurun_out/j_c3_4cta.json 2> gpurun_out/j_err3.err; echo "rc=$?"
GT_BUCKET_CTAS=3 timeout 300 $B > gpurun_out/j_c3_3cta.json 2> gpurun_out/j_err4.err; echo "rc=$?"
