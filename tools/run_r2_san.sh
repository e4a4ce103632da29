mkdir -p gpurun_out
python tools/sanitize_small.py > gpurun_out/san_plain.log 2>&1 && timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/san_memcheck.log 2>&1
echo "rc=$?"; tail -15 gpurun_out/san_plain.log; grep -E "ERROR SUMMARY|Invalid|error" gpurun_out/san_memcheck.log | head -20; tail -5 gpurun_out/san_memcheck.log
