# refresh of the ncu launch lists after the last planner changes (two lanes for ResNet-152 B=128, fused FP8 hand-over)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_fp8.py tests/test_gpu_modules.py -m gpu -x -q -k "refusals or host_and_uint8 or argument_errors" > gpurun_out/r2_p_t.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_p_t.log
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
run() {
  python tools/ncu_step.py $2 $3 $4 > gpurun_out/ncu_plain_$1.log 2>&1 &&
  ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r2_$1.csv python tools/ncu_step.py $2 $3 $4 > gpurun_out/ncu_$1.log 2>&1
  echo "$1 rc=$?"
}
run r50fp8 resnet50 256 fp8
run r152bf16 resnet152 128 bf16
