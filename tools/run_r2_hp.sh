#!/bin/bash
# host-pack bring-up: new GPU tests, then the headline bench (stderr keeps the RNB_VERBOSE decision line)
mkdir -p gpurun_out
nproc > gpurun_out/hp_nproc.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/hp_nproc.txt
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_stem.py -x -q -m gpu -k "host_pack" 2>&1 | tail -15 > gpurun_out/hp_tests.txt
cat gpurun_out/hp_tests.txt
RNB_VERBOSE=1 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_hp.json 2> gpurun_out/bench_hp.err
tail -3 gpurun_out/bench_hp.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_hp.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], json.dumps(d['e2e']))
PY
