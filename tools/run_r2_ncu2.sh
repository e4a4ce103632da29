mkdir -p gpurun_out
export RNB_AUTOTUNE=0
python tools/ncu_step.py resnet50 256 fp8 > gpurun_out/ncu_plain_c3.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_igemm_kernel -s 1 -c 2 -o gpurun_out/prof_r2_fp8_c3res python tools/ncu_step.py resnet50 256 fp8 > gpurun_out/ncu_full_c3.log 2>&1
echo "rc=$?"; ls -la gpurun_out/prof_r2_fp8_c3res.ncu-rep
