mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_stem.py -m gpu -x -q > gpurun_out/r2_aa_stem.log 2>&1; echo "stem pytest rc=$?"; tail -5 gpurun_out/r2_aa_stem.log
t0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_aa_gpu.log 2>&1; echo "pytest rc=$? secs=$(( $(date +%s)-t0 ))" >> gpurun_out/r2_aa_gpu.log
tail -5 gpurun_out/r2_aa_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/r2_aa_prof_r50.json > gpurun_out/r2_aa_bench_r50.json 2> gpurun_out/r2_aa_bench_r50.err; echo "bench r50 rc=$?"
timeout 300 python bench.py --arch resnet18 --dtype tf32 --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/r2_aa_prof_r18.json > gpurun_out/r2_aa_bench_r18.json 2> gpurun_out/r2_aa_bench_r18.err; echo "bench r18 rc=$?"
for f in r50 r18; do python - <<PY
import json
d=json.loads(open("gpurun_out/r2_aa_bench_$f.json").read().strip().splitlines()[-1])
p=json.load(open("gpurun_out/r2_aa_prof_$f.json"))
print("$f", round(d["value"]), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],4), "sus", round(d["sustained"]["ms_per_step"],4), d["parity"]["ok"], d["parity"]["rel_err"], "launches", d["gpu_launches"], "stem us", [round(r["us"],1) for r in p["launches"][:2]])
PY
done
