mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -x -q -k "tile_families or resident_weight" > gpurun_out/r2_r_t.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_r_t.log
RNB_VERBOSE=2 RNB_LANES=1 python tools/ncu_step.py resnet50 256 bf16 2>&1 | grep -E "autotune.*(k3|1024->256|2048->512|512->128|256->128|1024->512)" | sort | uniq > gpurun_out/autotune_r50.txt; cat gpurun_out/autotune_r50.txt | cut -c1-160
python tools/ab.py resnet50 256 "" > gpurun_out/ab10_r50.txt 2>&1; cat gpurun_out/ab10_r50.txt
