mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_group.py tests/test_gpu_modules.py tests/test_preprocess.py -m gpu -q > gpurun_out/r2_c_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_c_tests.log
tail -25 gpurun_out/r2_c_tests.log
W=$(python -c "from resnet_c_b200 import weights; print(weights.cached_weights_dir('resnet50', 0))")
for mode in direct copy; do
  RNB_GROUP_GATHER=$mode ./build/resnet_infer_mgpu resnet50 bf16 2 256 20 $W > gpurun_out/mgpu2_$mode.txt 2> gpurun_out/mgpu2_$mode.err; echo "mgpu $mode rc=$?"
  cat gpurun_out/mgpu2_$mode.txt; tail -3 gpurun_out/mgpu2_$mode.err
done
./build/resnet_infer_mgpu resnet50 bf16 1 256 20 $W > gpurun_out/mgpu1.txt 2>&1; cat gpurun_out/mgpu1.txt
nvidia-smi topo -m > gpurun_out/topo2.txt 2>&1
