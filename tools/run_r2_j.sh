mkdir -p gpurun_out
timeout 600 python bench.py --arch resnet152 --batch 128 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_r152_lanes.json 2> gpurun_out/r2_bench_r152_lanes.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_r50_b.json 2> gpurun_out/r2_bench_r50_b.err; echo "bench rc=$?"
for f in gpurun_out/r2_bench_r152_lanes.json gpurun_out/r2_bench_r50_b.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1])
print(round(d['value']), d['ms_per_step'], d['roofline']['frac'], d['parity']['ok'], round(d['e2e']['value']), d['gpu_launches'], d['sustained']['ms_per_step'], d['clocks'])"; done
