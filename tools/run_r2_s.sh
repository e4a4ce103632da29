mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -x -q -k "resident_weight" > gpurun_out/r2_s_t1.log 2>&1; echo "pytest(first-in-process) rc=$?"; tail -2 gpurun_out/r2_s_t1.log
timeout 1200 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py -m gpu -x -q > gpurun_out/r2_s_t2.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_s_t2.log
