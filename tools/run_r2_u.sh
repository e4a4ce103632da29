mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_stem.py -m gpu -x -q > gpurun_out/r2_u_stem.log 2>&1; echo "stem pytest rc=$?"; tail -15 gpurun_out/r2_u_stem.log
timeout 300 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/r2_u_prof.json > gpurun_out/r2_u_bench.json 2> gpurun_out/r2_u_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r2_u_bench.json; tail -3 gpurun_out/r2_u_bench.err
RNB_STEM_FUSED=0 timeout 300 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/r2_u_prof_unfused.json > gpurun_out/r2_u_bench_unfused.json 2> gpurun_out/r2_u_bench_unfused.err; echo "bench unfused rc=$?"; cut -c1-400 gpurun_out/r2_u_bench_unfused.json
