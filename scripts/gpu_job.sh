set -x
for g in 0 32 64 128; do
GT_L2_FETCH_BYTES=$g timeout 300 python bench.py --workload c2q --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/j_c2q_l2_$g.json 2> gpurun_out/j_c2q_l2_$g.err
done
GT_L2_FETCH_BYTES=32 timeout 300 python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/j_c3_l2_32.json 2> gpurun_out/j_c3_l2_32.err
python - <<'PY'
import ctypes, torch
rt = ctypes.CDLL("libcudart.so")
v = ctypes.c_size_t(0)
print("default L2 fetch granularity:", rt.cudaDeviceGetLimit(ctypes.byref(v), 0x05), v.value)
PY
