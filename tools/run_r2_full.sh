mkdir -p gpurun_out
t0=$(date +%s)
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_full_gpu.log 2>&1; echo "pytest rc=$? secs=$(( $(date +%s)-t0 ))" >> gpurun_out/r2_full_gpu.log
tail -6 gpurun_out/r2_full_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r2_bench_ref.json
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_final.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), "frac", d["roofline"]["frac"], "sus", d["sustained"]["ms_per_step"], d["parity"]["ok"], d["cpu_baseline"])
PY
