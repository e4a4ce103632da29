mkdir -p gpurun_out
python -m pytest tests/test_gpu_fp8.py tests/test_gpu_model.py -m gpu -x -q -k "fp8 or two_lanes" > gpurun_out/r2_l_t.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_l_t.log
RNB_VERBOSE=1 AB_DTYPE=fp8 python tools/ab.py resnet50 256 "" "RNB_LANES=1" "RNB_LANES=2" 2>&1 | grep -E "rnb lanes|resnet50" > gpurun_out/ab8_fp8_r50.txt; cat gpurun_out/ab8_fp8_r50.txt
RNB_VERBOSE=1 AB_DTYPE=fp8 python tools/ab.py resnet152 128 "" "RNB_LANES=1" 2>&1 | grep -E "rnb lanes|resnet152" > gpurun_out/ab8_fp8_r152.txt; cat gpurun_out/ab8_fp8_r152.txt
RNB_VERBOSE=1 AB_DTYPE=tf32 python tools/ab.py resnet18 256 "" "RNB_LANES=1" 2>&1 | grep -E "rnb lanes|resnet18" > gpurun_out/ab8_tf32_r18.txt; cat gpurun_out/ab8_tf32_r18.txt
