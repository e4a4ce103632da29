# round-2 validation: all GPU tests, the headline bench with per-launch profile, ResNet-152 B=128, binary timings
mkdir -p gpurun_out
t0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t2.log 2>&1; echo "pytest rc=$? secs=$(( $(date +%s)-t0 ))" >> gpurun_out/r2_t2.log
t0=$(date +%s)
timeout 600 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/launches_r2a.json > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$? secs=$(( $(date +%s)-t0 ))"
timeout 600 python bench.py --arch resnet152 --batch 128 --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/launches_r2a_r152.json > gpurun_out/r2_bench1_r152.json 2> gpurun_out/r2_bench1_r152.err
timeout 600 python tools/time_binaries.py > gpurun_out/time_binaries.json 2> gpurun_out/time_binaries.err
tail -5 gpurun_out/r2_t2.log; cut -c1-1500 gpurun_out/r2_bench1.json; cut -c1-400 gpurun_out/r2_bench1_r152.json; cat gpurun_out/time_binaries.json
