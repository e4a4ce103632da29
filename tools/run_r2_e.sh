mkdir -p gpurun_out
python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "scheduling_switches or stated_config or fused_bottleneck" > gpurun_out/r2_e_t.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_e_t.log
python tools/ab.py resnet152 128 "" "RNB_C3N1_HYBRID=0" "RNB_C3N1=1" > gpurun_out/ab4_r152.txt 2>&1; cat gpurun_out/ab4_r152.txt
python tools/ab.py resnet50 256 "" "RNB_C3N1_HYBRID=0" > gpurun_out/ab4_r50.txt 2>&1; cat gpurun_out/ab4_r50.txt
python tools/ab.py resnet50 128 "" "RNB_C3N1_HYBRID=0" > gpurun_out/ab4_r50_b128.txt 2>&1; cat gpurun_out/ab4_r50_b128.txt
