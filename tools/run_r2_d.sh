mkdir -p gpurun_out
python -m pytest tests/test_gpu_fp8.py -m gpu -x -q > gpurun_out/r2_d_t.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_d_t.log
AB_DTYPE=fp8 python tools/ab.py resnet50 256 "" > gpurun_out/ab3_fp8_r50.txt 2>&1; cat gpurun_out/ab3_fp8_r50.txt
AB_DTYPE=fp8 python tools/ab.py resnet152 128 "" > gpurun_out/ab3_fp8_r152.txt 2>&1; cat gpurun_out/ab3_fp8_r152.txt
AB_DTYPE=fp8 python tools/ab.py resnet18 256 "" > gpurun_out/ab3_fp8_r18.txt 2>&1; cat gpurun_out/ab3_fp8_r18.txt
AB_DTYPE=tf32 python tools/ab.py resnet18 256 "" > gpurun_out/ab3_tf32_r18.txt 2>&1; cat gpurun_out/ab3_tf32_r18.txt
timeout 600 python bench.py --arch resnet18 --dtype tf32 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_r18tf32.json 2> gpurun_out/r2_bench_r18tf32.err; cut -c1-300 gpurun_out/r2_bench_r18tf32.json
