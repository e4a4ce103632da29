mkdir -p gpurun_out
t0=$(date +%s)
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_full_gpu2.log 2>&1; echo "pytest rc=$? secs=$(( $(date +%s)-t0 ))" >> gpurun_out/r2_full_gpu2.log
tail -6 gpurun_out/r2_full_gpu2.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke2.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_smoke2.log
timeout 600 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/launches_r2_final.json > gpurun_out/r2_bench_final2.json 2> gpurun_out/r2_bench_final2.err; echo "bench rc=$?"
timeout 600 python bench.py --arch resnet152 --batch 128 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_r152_final.json 2> gpurun_out/r2_bench_r152_final.err; echo "bench rc=$?"
timeout 600 python bench.py --dtype fp8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_fp8_final.json 2> gpurun_out/r2_bench_fp8_final.err; echo "bench rc=$?"
for f in gpurun_out/r2_bench_final2.json gpurun_out/r2_bench_r152_final.json gpurun_out/r2_bench_fp8_final.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1])
print(round(d['value']), d['ms_per_step'], 'frac', round(d['roofline']['frac'],4), d['parity']['ok'], round(d['parity']['rel_err'],4), 'e2e', round(d['e2e']['value']), d['gpu_launches'], 'sus', d['sustained']['ms_per_step'], d['clocks'])"; done
