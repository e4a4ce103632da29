mkdir -p gpurun_out
python -m pytest tests/test_gpu_fp8.py -m gpu -x -q -s > gpurun_out/r2_g_t.log 2>&1; echo "pytest rc=$?"; grep -E "rel err|passed|failed|Error" gpurun_out/r2_g_t.log | tail -16
AB_DTYPE=fp8 python tools/ab.py resnet50 256 "" "RNB_FP8_HANDOVER=0" > gpurun_out/ab6_fp8_r50.txt 2>&1; cat gpurun_out/ab6_fp8_r50.txt
AB_DTYPE=fp8 python tools/ab.py resnet152 128 "" "RNB_FP8_HANDOVER=0" > gpurun_out/ab6_fp8_r152.txt 2>&1; cat gpurun_out/ab6_fp8_r152.txt
