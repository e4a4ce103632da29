mkdir -p gpurun_out
python tools/ab.py resnet50 256 "" "RNB_NO_SPLIT=1" > gpurun_out/ab_split_r50.txt 2>&1
python tools/ab.py resnet152 128 "" "RNB_NO_SPLIT=1" "RNB_C3N1=1" "RNB_C3N1=1 RNB_NO_SPLIT=1" > gpurun_out/ab_split_r152.txt 2>&1
python -m pytest tests/test_gpu_group.py -m gpu -x -q > gpurun_out/r2_group_1gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_group_1gpu.log
cat gpurun_out/ab_split_r50.txt gpurun_out/ab_split_r152.txt; tail -5 gpurun_out/r2_group_1gpu.log
