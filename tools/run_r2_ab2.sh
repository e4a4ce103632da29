mkdir -p gpurun_out
python -m pytest tests/test_gpu_conv.py tests/test_gpu_fp8.py -m gpu -x -q > gpurun_out/r2_ab2_t.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_ab2_t.log
python tools/ab.py resnet50 256 "" > gpurun_out/ab2_bf16.txt 2>&1; cat gpurun_out/ab2_bf16.txt
for c in 0 1 2; do RNB_FP8_CFG=$c AB_DTYPE=fp8 python tools/ab.py resnet50 256 "" > gpurun_out/ab2_fp8_cfg$c.txt 2>&1; cat gpurun_out/ab2_fp8_cfg$c.txt; done
