mkdir -p gpurun_out
t0=$(date +%s)
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_full_gpu3.log 2>&1; echo "pytest rc=$? secs=$(( $(date +%s)-t0 ))" >> gpurun_out/r2_full_gpu3.log
tail -6 gpurun_out/r2_full_gpu3.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke3.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_smoke3.log
