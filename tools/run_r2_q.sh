mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -x -q -k "resident_weight or tail_split or bottleneck_layer_shapes" > gpurun_out/r2_q_t.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_q_t.log
RNB_VERBOSE=1 python tools/ab.py resnet50 256 "" 2>&1 | grep -E "resident|resnet50 B" | cut -c1-200 > gpurun_out/ab9_r50.txt; cat gpurun_out/ab9_r50.txt
python - <<PY
import os, sys, time, torch
sys.path.insert(0, ".")
from resnet_c_b200 import engine
g = torch.Generator().manual_seed(0)
x = torch.randn(256, 128, 28, 28, generator=g).cuda(); w = (torch.randn(128,128,3,3,generator=g)*0.04).cuda()
for tile in ("1128", "31128", "1128", "31128"):
    os.environ["RNB_FORCE_TILE"] = tile
    for _ in range(3): engine.conv_bn_act_forward(x, w, None, None, True, 1, 1, "bf16")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): engine.conv_bn_act_forward(x, w, None, None, True, 1, 1, "bf16")
    e1.record(); torch.cuda.synchronize()
    print("tile", tile, "per call incl. layout conversions + packing: %.1f us" % (e0.elapsed_time(e1) * 100))
PY
