set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_model.py -m gpu -x -q -s -k "stated_config or every_fusion" > gpurun_out/r2_t1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t1.log
python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/launches_r2a.json > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
python bench.py --arch resnet152 --batch 128 --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/launches_r2a_r152.json > gpurun_out/r2_bench1_r152.json 2> gpurun_out/r2_bench1_r152.err
python tools/time_binaries.py > gpurun_out/time_binaries.json 2> gpurun_out/time_binaries.err
tail -3 gpurun_out/r2_t1.log; cat gpurun_out/r2_bench1.json | cut -c1-600
