mkdir -p gpurun_out
RNB_NO_GRAPH=1 python tools/stem_epi_ab.py 256 0 > gpurun_out/stem_w.txt 2>&1; tail -2 gpurun_out/stem_w.txt
t0=$(date +%s)
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_w_gpu.log 2>&1; echo "pytest rc=$? secs=$(( $(date +%s)-t0 ))" >> gpurun_out/r2_w_gpu.log
tail -6 gpurun_out/r2_w_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_w_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_w_smoke.log
timeout 300 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/r2_w_prof.json > gpurun_out/r2_w_bench.json 2> gpurun_out/r2_w_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2_w_bench.json
