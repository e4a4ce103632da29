set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_t2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t2.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
tail -5 gpurun_out/r2_t2.log; cut -c1-300 gpurun_out/r2_bench2.json
