set -x
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 600 $R --master-port 29541 bench.py --gpus 4 --steps 3 --warmup 3 --no-e2e > gpurun_out/j_n4_win.json 2> gpurun_out/j_n4_win.err; echo "rc=$?"
GT_APPLY_WINDOWS=0 timeout 600 $R --master-port 29542 bench.py --gpus 4 --steps 3 --warmup 3 --no-e2e > gpurun_out/j_n4_atomics.json 2> gpurun_out/j_n4_atomics.err; echo "rc=$?"
