set -x
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 $B --workload c3 > gpurun_out/j_c3.json 2> gpurun_out/j_err1.err; echo "rc=$?"
GT_BUCKET_OVERLAP=0 timeout 300 $B --workload c3 > gpurun_out/j_c3_noovl.json 2> gpurun_out/j_err2.err; echo "rc=$?"
timeout 300 $B --workload c2 > gpurun_out/j_c2.json 2> gpurun_out/j_err3.err; echo "rc=$?"
timeout 300 $B --workload c5 > gpurun_out/j_c5.json 2> gpurun_out/j_err4.err; echo "rc=$?"
timeout 300 $B --workload c1 > gpurun_out/j_c1.json 2> gpurun_out/j_err5.err; echo "rc=$?"
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
