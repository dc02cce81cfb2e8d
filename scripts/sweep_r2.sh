run() { # name workload env...
  name=$1; wl=$2; shift 2
  env "$@" timeout 300 python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/sw_$name.json 2> gpurun_out/sw_$name.err || echo "FAIL $name"
}
run c3_s25_w15 c3 GT_SLICE_LOG2_BYTES=25 GT_WINDOW_LOG2_BYTES=15
run c3_s25_w17 c3 GT_SLICE_LOG2_BYTES=25 GT_WINDOW_LOG2_BYTES=17
run c3_s24_w16 c3 GT_SLICE_LOG2_BYTES=24 GT_WINDOW_LOG2_BYTES=16
run c3_s26_w17 c3 GT_SLICE_LOG2_BYTES=26 GT_WINDOW_LOG2_BYTES=17
run c2_s25_w16 c2 GT_SLICE_LOG2_BYTES=25 GT_WINDOW_LOG2_BYTES=16
run c2_s26_w17 c2 GT_SLICE_LOG2_BYTES=26 GT_WINDOW_LOG2_BYTES=17
run c2_s25_w17 c2 GT_SLICE_LOG2_BYTES=25 GT_WINDOW_LOG2_BYTES=17
run c5_s25_w16 c5 GT_SLICE_LOG2_BYTES=25 GT_WINDOW_LOG2_BYTES=16
run c5_s26_w17 c5 GT_SLICE_LOG2_BYTES=26 GT_WINDOW_LOG2_BYTES=17
run c3_noovl c3 GT_BUCKET_OVERLAP=0
