mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fp8.py -m gpu -q -s -x > gpurun_out/r2_fp8_t.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_fp8_t.log
tail -40 gpurun_out/r2_fp8_t.log
timeout 600 python bench.py --dtype fp8 --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/launches_r2_fp8.json > gpurun_out/r2_bench_fp8.json 2> gpurun_out/r2_bench_fp8.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/r2_bench_fp8.json; tail -5 gpurun_out/r2_bench_fp8.err
AB_DTYPE=fp8 python tools/ab.py resnet50 256 "" "RNB_FP8_FROM=0" > gpurun_out/ab_fp8_from.txt 2>&1; cat gpurun_out/ab_fp8_from.txt
python tools/ab.py resnet50 256 "" > gpurun_out/ab_bf16_ref.txt 2>&1; cat gpurun_out/ab_bf16_ref.txt
