# ncu launch lists of one step per config (metrics pass), then one --set full capture of the FP8 pair kernel and the
# BF16 layer4 3x3 with the tail split. All ncu runs of this call count as one tool.
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
run() {  # name arch batch dtype
  python tools/ncu_step.py $2 $3 $4 > gpurun_out/ncu_plain_$1.log 2>&1 &&
  ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r2_$1.csv python tools/ncu_step.py $2 $3 $4 > gpurun_out/ncu_$1.log 2>&1
  echo "$1 rc=$?"
}
export RNB_AUTOTUNE=0
run r50bf16 resnet50 256 bf16
run r50fp8 resnet50 256 fp8
run r152bf16 resnet152 128 bf16
run r18tf32 resnet18 256 tf32
python tools/ncu_step.py resnet50 256 fp8 > gpurun_out/ncu_plain_full.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_igemm2_kernel -s 12 -c 3 -o gpurun_out/prof_r2_fp8_pair python tools/ncu_step.py resnet50 256 fp8 > gpurun_out/ncu_full_fp8.log 2>&1
echo "full fp8 rc=$?"
python tools/ncu_step.py resnet50 256 bf16 > gpurun_out/ncu_plain_full2.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_igemm2_kernel -s 18 -c 2 -o gpurun_out/prof_r2_bf16_l4 python tools/ncu_step.py resnet50 256 bf16 > gpurun_out/ncu_full_bf16.log 2>&1
echo "full bf16 rc=$?"
ls -la gpurun_out/*.ncu-rep
