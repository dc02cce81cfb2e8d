set -x
timeout 900 python -m pytest tests/test_fastx.py tests/test_packed.py tests/test_gpu_processors.py -x -q -m gpu > gpurun_out/j_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/j_pytest.log
timeout 900 python scripts/bench_fastx.py --reads 16000000 --out gpurun_out/j_fastx2.json > gpurun_out/j_fastx2.log 2>&1; echo "rc=$?"
GT_FASTX_THREADS=16 timeout 900 python scripts/bench_fastx.py --reads 16000000 --out gpurun_out/j_fastx2_t16.json > gpurun_out/j_fastx2_t16.log 2>&1; echo "rc=$?"
GT_FASTX_THREADS=12 timeout 900 python scripts/bench_fastx.py --reads 16000000 --out gpurun_out/j_fastx2_t12.json > gpurun_out/j_fastx2_t12.log 2>&1; echo "rc=$?"
