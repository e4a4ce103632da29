mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fp8.py -m gpu -x -q > gpurun_out/r2_t_t.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_t_t.log
RNB_VERBOSE=1 AB_DTYPE=fp8 python tools/ab.py resnet50 256 "" 2>&1 | grep -E "deepest|resnet50 B" | cut -c1-170 > gpurun_out/ab11_fp8.txt; grep -c deepest gpurun_out/ab11_fp8.txt; tail -1 gpurun_out/ab11_fp8.txt
