mkdir -p gpurun_out
python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "scheduling_switches or fused_bottleneck or stated_config or launch_accounting" > gpurun_out/r2_m_t.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_m_t.log
RNB_VERBOSE=1 python tools/batch_sweep.py resnet50 bf16 96 97 128 193 194 222 256 2>&1 | grep -E "conv3 \+ next|resnet50 bf16" | tee gpurun_out/sweep2_r50.txt
RNB_VERBOSE=1 python tools/batch_sweep.py resnet152 bf16 128 2>&1 | grep -E "conv3 \+ next|rnb lanes|resnet152 bf16" | tee gpurun_out/sweep2_r152.txt
