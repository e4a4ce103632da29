mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_stem.py -m gpu -x -q > gpurun_out/r2_y_stem.log 2>&1; echo "stem pytest rc=$?"; tail -3 gpurun_out/r2_y_stem.log
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "tf32 or launch_accounting or resnet18" > gpurun_out/r2_y_model.log 2>&1; echo "model pytest rc=$?"; tail -3 gpurun_out/r2_y_model.log
timeout 300 python bench.py --arch resnet18 --dtype tf32 --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/r2_y_prof_r18.json > gpurun_out/r2_y_bench_r18.json 2> gpurun_out/r2_y_bench_r18.err; echo "bench r18 tf32 rc=$?"; cut -c1-250 gpurun_out/r2_y_bench_r18.json
