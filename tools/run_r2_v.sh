mkdir -p gpurun_out
python tools/ncu_step.py resnet50 256 bf16 > gpurun_out/ncu_plain_v.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:stem_tc_t_kernel -c 1 -f -o gpurun_out/stem_fused_r2 python tools/ncu_step.py resnet50 256 bf16 > gpurun_out/ncu_v.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_v.log; ls -la gpurun_out/stem_fused_r2.ncu-rep
