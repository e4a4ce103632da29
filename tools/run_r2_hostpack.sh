#!/bin/bash
# host packing: its GPU tests, then the interleaved A/B of the serving loop (profiles/e2e_hostpack_ab_r2.txt)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "host_pack or uint8_input" 2>&1 | tail -5 > gpurun_out/hp_tests.txt
cat gpurun_out/hp_tests.txt
RNB_VERBOSE=1 timeout 600 python tools/e2e_ab.py resnet50 256 bf16 40 2> gpurun_out/e2e_ab.err | tee gpurun_out/e2e_ab_r50.txt
grep "host pack" gpurun_out/e2e_ab.err
