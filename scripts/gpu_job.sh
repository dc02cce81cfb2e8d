# What a verification call on a GPU box runs (gpurun -- 'bash scripts/gpu_job.sh'): the whole GPU suite, the smoke
# test and the default bench line.  Multi-GPU: gpurun --gpus N -- 'python -m torch.distributed.run --nnodes=1
# --nproc-per-node N --master-addr 127.0.0.1 bench.py --gpus N'.
set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
timeout 300 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?"
