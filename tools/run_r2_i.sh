mkdir -p gpurun_out
python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "two_lanes or forwards_on_different or fresh_output or full_batch" > gpurun_out/r2_i_t.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_i_t.log
RNB_VERBOSE=1 python tools/ab.py resnet152 128 "" "RNB_LANES=1" "RNB_LANES=2" 2>&1 | grep -E "rnb lanes|resnet152" > gpurun_out/ab7_r152.txt; cat gpurun_out/ab7_r152.txt
RNB_VERBOSE=1 python tools/ab.py resnet50 256 "" "RNB_LANES=2" 2>&1 | grep -E "rnb lanes|resnet50" > gpurun_out/ab7_r50.txt; cat gpurun_out/ab7_r50.txt
RNB_VERBOSE=1 python tools/ab.py resnet50 128 "" "RNB_LANES=1" 2>&1 | grep -E "rnb lanes|resnet50" > gpurun_out/ab7_r50_b128.txt; cat gpurun_out/ab7_r50_b128.txt
timeout 600 python bench.py --arch resnet152 --batch 128 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_r152_lanes.json 2> gpurun_out/r2_bench_r152_lanes.err; echo "bench rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_r152_lanes.json').read().strip().splitlines()[-1])
print(round(d['value']), d['ms_per_step'], d['roofline']['frac'], d['parity']['ok'], d['e2e']['value'], d['gpu_launches'], d['sustained']['ms_per_step'])"
