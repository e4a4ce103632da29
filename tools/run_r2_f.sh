mkdir -p gpurun_out
python -m pytest tests/test_gpu_model.py tests/test_gpu_conv.py -m gpu -x -q -k "scheduling_switches or sixteen or tail_split" > gpurun_out/r2_f_t.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_f_t.log
python tools/ab.py resnet152 128 "" "RNB_C3N1_AUTO=0" > gpurun_out/ab5_r152.txt 2>&1; cat gpurun_out/ab5_r152.txt
python tools/ab.py resnet50 128 "" "RNB_C3N1_AUTO=0" > gpurun_out/ab5_r50_b128.txt 2>&1; cat gpurun_out/ab5_r50_b128.txt
python tools/ab.py resnet50 256 "" > gpurun_out/ab5_r50.txt 2>&1; cat gpurun_out/ab5_r50.txt
AB_DTYPE=fp8 python tools/ab.py resnet50 256 "" > gpurun_out/ab5_fp8_r50.txt 2>&1; cat gpurun_out/ab5_fp8_r50.txt
RNB_VERBOSE=1 AB_DTYPE=fp8 python tools/ab.py resnet50 256 "" 2>&1 | grep "rnb plan" | awk '{print $NF, $0}' | cut -c1-160 > gpurun_out/plan_fp8_r50.txt; grep -c "single-16w" gpurun_out/plan_fp8_r50.txt; RNB_VERBOSE=1 python tools/ab.py resnet50 256 "" 2>&1 | grep "rnb plan" | cut -c1-160 > gpurun_out/plan_bf16_r50.txt; grep -c "single-16w" gpurun_out/plan_bf16_r50.txt
