set -x
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/j_pytest_all.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/j_pytest_all.log
