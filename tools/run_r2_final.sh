# final validation of round 2 (after the stem rebuild): GPU suite, smoke, bench lines of four configs, the reference arm,
# ncu launch lists of the headline and of ResNet-18 TF32
mkdir -p gpurun_out
t0=$(date +%s)
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_gpu.log 2>&1; echo "pytest rc=$? secs=$(( $(date +%s)-t0 ))" >> gpurun_out/r2f_gpu.log
tail -4 gpurun_out/r2f_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2f_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/r2f_prof_r50.json > gpurun_out/r2f_bench_r50.json 2> gpurun_out/r2f_bench_r50.err; echo "bench r50 rc=$?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/r2f_bench_ref.json
timeout 600 python bench.py --arch resnet152 --batch 128 --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/r2f_prof_r152.json > gpurun_out/r2f_bench_r152.json 2> gpurun_out/r2f_bench_r152.err; echo "bench r152 rc=$?"
timeout 600 python bench.py --dtype fp8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_bench_fp8.json 2> gpurun_out/r2f_bench_fp8.err; echo "bench fp8 rc=$?"
timeout 600 python bench.py --arch resnet18 --dtype tf32 --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/r2f_prof_r18.json > gpurun_out/r2f_bench_r18.json 2> gpurun_out/r2f_bench_r18.err; echo "bench r18 rc=$?"
for f in r50 r152 fp8 r18; do python - <<PY
import json
d=json.loads(open("gpurun_out/r2f_bench_$f.json").read().strip().splitlines()[-1])
print("$f", round(d["value"]), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "u8", round(d.get("e2e_u8",{}).get("value",0)), "frac", round(d["roofline"]["frac"],4), "sus", round(d["sustained"]["ms_per_step"],4), d["parity"]["ok"], d["parity"]["rel_err"], "launches", d["gpu_launches"], d["clocks"]["reasons"])
PY
done
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
run() {  # name arch batch dtype
  python tools/ncu_step.py $2 $3 $4 > gpurun_out/ncu_plain_$1.log 2>&1 &&
  ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r2f_$1.csv python tools/ncu_step.py $2 $3 $4 > gpurun_out/ncu_$1.log 2>&1
  echo "$1 rc=$?"
}
export RNB_AUTOTUNE=0
run r50bf16 resnet50 256 bf16
run r18tf32 resnet18 256 tf32
